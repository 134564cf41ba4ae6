// Tile-union WF forward for 16-bit types (sm_100a): out[b,i,ic,c] = sum_j w[b,i,j,ic] f[b,idx[b,i,j],c], IC = 4.
//
// ncu on the token-by-token kernel of clusten_wf2.cu (profiles/r1_wf2_small_s0_bf16_v8_ncu_keys.txt) showed it bound by L1 data-pipe
// wavefronts (81 %): every token pulls its own M rows through L1 in 4-rows-per-instruction fragments, hit or miss.  Here a
// tile = 16 consecutive tokens of the plan's order (spatial neighbours, wf2.cuh) and the rows of the UNION of their
// octets are staged ONCE per tile: cp.async (coalesced 16-byte chunks) into shared memory, ldmatrix.trans into B
// fragments.  The tile's weights are staged transposed ([ic][token][j]) so that one ldmatrix per ic yields the whole A
// fragment: each lane points its row at (token, slot of the octet in that token's neighbourhood) or at a zero chunk when
// the token does not reference the octet.  A warp owns 8*NT channels of a tile and all four ic (B fragments are re-used
// four times); NW = C / (8*NT) warps share a tile's weights.  Tokens outside the tile structure (plan.timp_list) are
// left to wf2's token-by-token kernel.
#include "t2.cuh"
#include "wf2.cuh"

namespace clusten {
namespace wf3 {

struct Args {
    const void *W, *F;
    void *out;
    int B, Nq, C, M;
    int64_t f_sb;
    int f_sn, rsw, tile_sm, ysm;         // rsw: bytes per staged weight row; tile_sm / ysm: bytes per tile / per warp
};

__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], uint32_t s) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(s));
}
__device__ __forceinline__ void sts32(uint32_t s, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(s), "r"(v)); }

constexpr int NSTG = 4;                  // cp.async ring depth of the feature stages (3 stages of 16 rows in flight per warp)

template <typename T, int NT>
__global__ void __launch_bounds__(256)
fwd_tile_kernel(const Args a, const WfPlanView pv, int NW) {
    extern __shared__ __align__(16) unsigned char dyn3[];
    constexpr int ROWB = NT * 16 + 16;                   // staged feature row: NT 16-byte chunks + pad (odd number of chunks)
    constexpr int STGB = 16 * ROWB;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int nwarps = blockDim.x >> 5, TPC = nwarps / NW;
    const int tic = warp / NW, cw = warp - tic * NW;
    const int64_t tb = min((int64_t)blockIdx.x * TPC + tic, (int64_t)a.B * pv.T - 1);   // surplus warps redo the last tile
    const int b = (int)(tb / pv.T), tile = (int)(tb - (int64_t)b * pv.T);
    const int p0 = tile * 16, nrow = min(16, a.Nq - p0);
    const uint32_t sm = (uint32_t)__cvta_generic_to_shared(dyn3);
    const uint32_t smW = sm + tic * a.tile_sm;                       // [4 ic][16 tokens][rsw] + one zero chunk
    const uint32_t smZ = smW + 64 * a.rsw;
    const uint32_t smY = sm + TPC * a.tile_sm + warp * a.ysm;        // NSTG stages x 16 rows x ROWB
    const int M = a.M, C = a.C;
    const int my_tok = pv.perm[(int64_t)b * a.Nq + p0 + min(lane & 15, nrow - 1)];     // lane r (and r + 16) holds token r
    const int U = pv.tile_u[tb];
    const int *octp = pv.tile_oct + tb * WF_UMAX;
    const int oc0 = octp[lane], oc1 = octp[32 + lane];
    auto octet = [&](int u) { u = min(u, max(U - 1, 0)); return __shfl_sync(FULL, u < 32 ? oc0 : oc1, u & 31); };
    const int cb = cw * 8 * NT;
    const T *Fb = t2::opaque(reinterpret_cast<const T *>(a.F) + b * a.f_sb + cb);
    const int P = (U + 1) >> 1;
    auto stage = [&](int p) {                                        // always commits a group (empty beyond P)
        if (p < P) {
            const int o0 = octet(2 * p) * 8, o1 = octet(2 * p + 1) * 8;
            const uint32_t dst = smY + (p % NSTG) * STGB;
#pragma unroll
            for (int x = 0; x < NT / 2; ++x) {
                const int ch = lane + 32 * x, row = ch / NT, blk = ch - row * NT;
                const int src = ((row < 8 ? o0 : o1) + (row & 7)) * a.f_sn + 8 * blk;
                t2::cp16(dst + row * ROWB + blk * 16, t2::at(Fb, src));
            }
        }
        t2::cp_commit();
    };
    // the tile's slot table (WF_UMAX rows of 16 bytes) rides in the first commit group, into this warp's own copy
    const uint32_t smS = smY + NSTG * STGB;
    {
        const int8_t *st = pv.slot_t + tb * WF_UMAX * 16;
        if (lane < U) t2::cp16(smS + 16 * lane, st + 16 * lane);
        if (lane + 32 < U) t2::cp16(smS + 16 * (lane + 32), st + 16 * (lane + 32));
    }
    // feature stages first: they do not depend on the weights, so their latency overlaps the weight transposition below
#pragma unroll
    for (int p = 0; p < NSTG - 1; ++p) stage(p);
    // ---- weights of the tile, transposed to [ic][token][j] (the NW warps of the tile share the work)
    {
        const T *Wb = reinterpret_cast<const T *>(a.W) + (int64_t)b * a.Nq * M * 4;
        const int hp = M >> 1, items = nrow * hp;
        for (int i0 = 0; i0 < items; i0 += NW * 32 * 4) {
            uint4 v[4];
            int rr[4], jj[4];
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                const int item = min(i0 + (x * NW + cw) * 32 + lane, items - 1);
                rr[x] = item / hp;
                jj[x] = item - rr[x] * hp;
                const int tok = __shfl_sync(FULL, my_tok, rr[x]);
                v[x] = t2::ldg16(Wb + ((int64_t)tok * M + 2 * jj[x]) * 4);
            }
#pragma unroll
            for (int x = 0; x < 4; ++x) {                            // (items clamped to the last one are rewritten with equal data)
                const uint32_t dst = smW + rr[x] * a.rsw + jj[x] * 4;
                sts32(dst, __byte_perm(v[x].x, v[x].z, 0x5410));
                sts32(dst + 16 * a.rsw, __byte_perm(v[x].x, v[x].z, 0x7632));
                sts32(dst + 32 * a.rsw, __byte_perm(v[x].y, v[x].w, 0x5410));
                sts32(dst + 48 * a.rsw, __byte_perm(v[x].y, v[x].w, 0x7632));
            }
        }
        if (cw == 0 && lane < 4) sts32(smZ + 4 * lane, 0u);
    }
    __syncthreads();
    float acc[4][NT][4];
#pragma unroll
    for (int ic = 0; ic < 4; ++ic)
#pragma unroll
        for (int n = 0; n < NT; ++n) acc[ic][n][0] = acc[ic][n][1] = acc[ic][n][2] = acc[ic][n][3] = 0.f;
    // ldmatrix row of this lane: A matrices (q & 1) -> token rows 0-7 / 8-15, (q >> 1) -> first / second octet of the pair
    const int q = lane >> 3, am = ((q & 1) << 3) + (lane & 7), awhich = q >> 1;
    const uint32_t slotp = smS + am + awhich * 16;
    const uint32_t arow = smW + am * a.rsw;
    const uint32_t lrow = smY + (((q & 1) << 3) + (lane & 7)) * ROWB + (q >> 1) * 16;
    for (int p = 0; p < P; ++p) {
        stage(p + NSTG - 1);
        t2::cp_wait<NSTG - 1>();
        __syncwarp();
        int sl = -1;
        if (2 * p + awhich < U) asm volatile("ld.shared.s8 %0, [%1];" : "=r"(sl) : "r"(slotp + 32 * p));
        const uint32_t abase = sl >= 0 ? arow + 16 * sl : smZ;
        const uint32_t astep = sl >= 0 ? 16 * a.rsw : 0;
        uint32_t af[4][4];
#pragma unroll
        for (int ic = 0; ic < 4; ++ic) ldsm4(af[ic], abase + ic * astep);
        const uint32_t yst = lrow + (p % NSTG) * STGB;
#pragma unroll
        for (int n = 0; n < NT; n += 2) {
            uint32_t bfr[4];
            t2::ldsm4t(bfr, yst + n * 16);
#pragma unroll
            for (int ic = 0; ic < 4; ++ic) {
                t2::mma16<T>(acc[ic][n], af[ic][0], af[ic][1], af[ic][2], af[ic][3], bfr[0], bfr[1]);
                t2::mma16<T>(acc[ic][n + 1], af[ic][0], af[ic][1], af[ic][2], af[ic][3], bfr[2], bfr[3]);
            }
        }
        __syncwarp();
    }
    t2::cp_wait<0>();
    if ((int64_t)blockIdx.x * TPC + tic != tb) return;               // surplus warp
    // ---- store: rows g / g+8 of the tile; tokens outside the tile structure have slot -1 everywhere (all-zero result) and
    // are overwritten afterwards by the token-by-token kernel, so they are simply written too
    T *Ob = reinterpret_cast<T *>(a.out) + (int64_t)b * a.Nq * 4 * C + cb + 2 * t;
    const int tka = __shfl_sync(FULL, my_tok, g), tkb = __shfl_sync(FULL, my_tok, g + 8);
    const int ta = g < nrow ? tka : -1, tb2 = g + 8 < nrow ? tkb : -1;
#pragma unroll
    for (int ic = 0; ic < 4; ++ic) {
        T *oa = Ob + ((int64_t)max(ta, 0) * 4 + ic) * C, *ob = Ob + ((int64_t)max(tb2, 0) * 4 + ic) * C;
#pragma unroll
        for (int n = 0; n < NT; ++n) {
            t2::st_pair_if<T>(oa + 8 * n, acc[ic][n][0], acc[ic][n][1], ta);
            t2::st_pair_if<T>(ob + 8 * n, acc[ic][n][2], acc[ic][n][3], tb2);
        }
    }
}

static inline bool pick_shape(int C, int &NT, int &NW) {
    static const int cand[5][2] = {{4, 4}, {6, 4}, {8, 4}, {6, 8}, {8, 8}};       // (NT, max NW), in order of preference
    for (const auto &c : cand) {
        if (C % (8 * c[0]) == 0) {
            const int nw = C / (8 * c[0]);
            if (nw <= c[1] && (nw & (nw - 1)) == 0) { NT = c[0]; NW = nw; return true; }
        }
    }
    return false;
}

template <typename T>
static int fwd_t(const T *w, const T *f, T *out, const void *plan, int B, int Nq, int Nk, int C, int M, int64_t f_sb, int64_t f_sn,
                 cudaStream_t st) {
    int NT, NW;
    if (!pick_shape(C, NT, NW)) return 0;
    if ((reinterpret_cast<uintptr_t>(f) | reinterpret_cast<uintptr_t>(w)) % 16 || reinterpret_cast<uintptr_t>(out) % 4 || f_sb % 8 || f_sn % 8)
        return 0;
    if ((int64_t)Nk * f_sn + C >= (1LL << 31)) return 0;
    const WfPlanView pv = wf_plan_view(const_cast<void *>(plan), B, Nq, M, Nk);
    Args a;
    a.W = w; a.F = f; a.out = out; a.B = B; a.Nq = Nq; a.C = C; a.M = M; a.f_sb = f_sb; a.f_sn = (int)f_sn;
    a.rsw = 2 * M + ((2 * M / 16) % 2 == 0 ? 16 : 0);
    a.tile_sm = 64 * a.rsw + 16;
    a.ysm = NSTG * 16 * (NT * 16 + 16) + WF_UMAX * 16;
    const int nwarps = NW > 4 ? NW : 4, TPC = nwarps / NW;
    const size_t smem = (size_t)TPC * a.tile_sm + (size_t)nwarps * a.ysm;
    if (smem > 160 * 1024) return 0;
    const int grid = ceil_div((int64_t)B * pv.T, TPC);
    static bool attr = false;                            // (one flag per T: this function is a template)
    if (!attr) {
        cudaFuncSetAttribute(fwd_tile_kernel<T, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        cudaFuncSetAttribute(fwd_tile_kernel<T, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        cudaFuncSetAttribute(fwd_tile_kernel<T, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        attr = true;
    }
    auto launch = [&](auto kern) { kern<<<grid, nwarps * 32, smem, st>>>(a, pv, NW); };
    if (NT == 4) launch(fwd_tile_kernel<T, 4>);
    else if (NT == 6) launch(fwd_tile_kernel<T, 6>);
    else launch(fwd_tile_kernel<T, 8>);
    note_launches(1);
    return 1;
}

}  // namespace wf3

// tile kernel + token-by-token kernel for the tokens it leaves out; 0 = shape not supported (caller falls back)
int wf3_fwd(const void *w, const void *f, const int64_t *idx, void *out, const void *plan, int B, int Nq, int Nk, int C, int M,
            int IC, int64_t f_sb, int64_t f_sn, int dtype, cudaStream_t st) {
    if (!plan || IC != 4 || M % 8 || M > 256 || (dtype != CLUSTEN_BF16 && dtype != CLUSTEN_F16)) return 0;
    int NT, NW;
    if (!wf3::pick_shape(C, NT, NW)) return 0;
    // the leftover kernel must be able to take the call too (same alignment rules, checked there)
    int ok;
    if (dtype == CLUSTEN_BF16)
        ok = wf3::fwd_t<__nv_bfloat16>((const __nv_bfloat16 *)w, (const __nv_bfloat16 *)f, (__nv_bfloat16 *)out, plan, B, Nq, Nk, C, M, f_sb, f_sn, st);
    else
        ok = wf3::fwd_t<__half>((const __half *)w, (const __half *)f, (__half *)out, plan, B, Nq, Nk, C, M, f_sb, f_sn, st);
    if (!ok) return 0;
    if (!wf2_fwd_listed(w, f, idx, out, plan, B, Nq, Nk, C, M, IC, f_sb, f_sn, dtype, st)) return -1;
    return 1;
}

}  // namespace clusten
