// Fused ClusterAttention core, forward, CTA-cooperative and TMA-staged (sm_100a).  Same arithmetic as clusten_fused.cu
// (mask2former/modeling/backbone/aff.py:114-155: QK logits + relative-position bias + cluster mask + blank token + softmax +
// attention-times-V), different data movement:
//
//   * a CTA owns one tile GROUP (64 consecutive tokens = 4 mma M-tiles, tile.cuh) and one head; its four warps share ONE staged
//     copy of the group's key / value octets instead of re-fetching them per (warp, head) in 64-byte pieces;
//   * an octet (8 consecutive key rows = one balanced cluster of the reference, point_utils.py:282-285) is one TMA box
//     [8 rows x C channels] per operand (cp.async.bulk.tensor.4d over a 4-D tensor map of the strided [B,H,N,C] operand, or ONE
//     box [8 rows x (k|v)] when k and v are the two halves of the kv Linear's output row, aff.py:105-113); rows beyond N are
//     zero-filled by the unit, completion is signalled on mbarriers (one per 4 octets: the warps start on the first octets while
//     the later ones are still in flight);
//   * the boxes land hardware-swizzled (SWIZZLE_32B/64B/128B by row width), so the ldmatrix / ld.shared fragment reads of the
//     mma.sync contraction are bank-conflict free without padding;
//   * a group whose union exceeds the staging capacity is processed in rounds (the capacity is a tuning knob, not a limit).
//
// The contraction stays on mma.sync (m16n8k16 for 16-bit, 3xTF32 m16n8k8 for fp32): per (16-token tile, head) the work is a
// [16 x C] x [C x 8U] product over the tile's own union (U ~ 19 octets), far below a tcgen05 tile (M >= 64 rows over the whole
// CTA union would triple the redundant columns), and the kernel is bound by data movement, not by tensor throughput.
#include <cuda.h>

#include "fused.cuh"

namespace clusten {
namespace tma {

constexpr int OCT_PER_BAR = 4;           // octets per mbarrier
constexpr int NBAR_MAX = 32;             // -> staging capacity <= 128 octets
constexpr int GW = GROUP_TILES;          // warps per CTA (one per 16-token tile of the group)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// (a lost completion must not hang the device: after ~2^22 probes the CTA traps and the launch reports an error)
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin)
        if (spin > (1u << 22)) __trap();
}
__device__ __forceinline__ void tma_box_4d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], uint32_t s) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(s));
}
__device__ __forceinline__ void ldsm2(uint32_t (&r)[2], uint32_t s) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];\n" : "=r"(r[0]), "=r"(r[1]) : "r"(s));
}
__device__ __forceinline__ void ldsm2t(uint32_t (&r)[2], uint32_t s) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];\n" : "=r"(r[0]), "=r"(r[1]) : "r"(s));
}
__device__ __forceinline__ float2 lds_f2(uint32_t s) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(s));
    return v;
}
// x = hi + lo with hi = x truncated to tf32 (one LOP3) and lo = x - hi (exact, one FADD); the tensor cores read the upper 19 bits of
// lo.  Two instructions instead of the seven `cvt.rna.tf32.f32` costs twice over (sm_100 emulates the rounding: FSETP, SEL, LOP3,
// VIADD ...); hi + lo carries 21 mantissa bits, the dropped lo x lo products are below 2^-20.
__device__ __forceinline__ void split3(uint32_t x, uint32_t &hi, uint32_t &lo) {
    hi = x & 0xffffe000u;
    lo = __float_as_uint(__uint_as_float(x) - __uint_as_float(hi));
}
__device__ __forceinline__ void mma3(float (&d)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], uint32_t b0h, uint32_t b1h,
                                     uint32_t b0l, uint32_t b1l) {
    t2::mma_tf32(d, al[0], al[1], al[2], al[3], b0h, b1h);
    t2::mma_tf32(d, ah[0], ah[1], ah[2], ah[3], b0l, b1l);
    t2::mma_tf32(d, ah[0], ah[1], ah[2], ah[3], b0h, b1h);
}

// Geometry of one staged octet = ONE box of 8 rows.  IB = bytes of one operand row (C channels).  PACKED: the box rows are [k | v]
// (2*IB bytes) and one load serves both contractions; else the staging area holds the K boxes during the dot phase and is refilled
// with the V boxes (same octets, same positions) while the softmax runs.  RB = row bytes of a box = its swizzle span.
template <typename T, int C, bool PACKED> struct Geo {
    static constexpr int IB = C * (int)sizeof(T);
    static constexpr int RB = PACKED ? 2 * IB : IB;
    static_assert(RB == 32 || RB == 64 || RB == 128, "box rows are 32, 64 or 128 bytes");
    static constexpr int BOX = 8 * RB;                     // bytes of one box = period of its swizzle pattern
    static constexpr int V_COL = PACKED ? IB : 0;          // byte offset of v inside a box row
    static constexpr int SH = RB == 128 ? 0 : RB == 64 ? 1 : 2;
    // byte offset inside a box of (row r, byte o of the row): 16-byte chunks XOR-ed with the 128-byte line index (CU_TENSOR_MAP_SWIZZLE_*)
    __device__ static constexpr int at(int r, int o) { return r * RB + ((((o >> 4) ^ ((r >> SH) & (RB / 16 - 1))) << 4) | (o & 15)); }
};

struct Launch { int cap, warp_bytes, rcp_qm; };            // staging capacity (octets, multiple of OCT_PER_BAR), per-warp scratch bytes,
                                                           // ceil(2^20 / (M / 4))

constexpr int ZPAD = 8;                                    // floats of zeros in front of every S row: "no slot" (-1) reads them

__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t lds32(uint32_t s) { return t2::lds32(s); }
__device__ __forceinline__ uint2 lds64(uint32_t s) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(s));
    return v;
}
__device__ __forceinline__ void sts_f2_if(uint32_t s, float a, float b, int slot) {
    asm volatile("{ .reg .pred p; setp.ge.s32 p, %3, 0; @p st.shared.v2.f32 [%0], {%1, %2}; }" ::"r"(s), "f"(a), "f"(b), "r"(slot) : "memory");
}
__device__ __forceinline__ int s8(uint32_t w, int byte) { return (int)(int8_t)(w >> (8 * byte)); }

// 16 x 8 logits of the warp's tokens against one staged K octet
template <typename T, int C, bool PACKED>
__device__ __forceinline__ void qk_octet(float (&acc)[4], const uint32_t (&qa)[sizeof(T) == 4 ? C / 8 : C / 16][4],
                                         const uint32_t (&ql)[sizeof(T) == 4 ? C / 8 : 1][4], uint32_t blk, const uint32_t *k_off) {
    constexpr bool F32 = sizeof(T) == 4;
    acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
    if constexpr (F32) {
#pragma unroll
        for (int s = 0; s < C / 8; ++s) {
            uint32_t b0h, b0l, b1h, b1l;
            split3(lds32(blk + k_off[2 * s]), b0h, b0l);
            split3(lds32(blk + k_off[2 * s + 1]), b1h, b1l);
            mma3(acc, qa[s], ql[s], b0h, b1h, b0l, b1l);
        }
    } else if constexpr (C == 32) {
        uint32_t kf[4];
        ldsm4(kf, blk + k_off[0]);
        t2::mma16<T>(acc, qa[0][0], qa[0][1], qa[0][2], qa[0][3], kf[0], kf[1]);
        t2::mma16<T>(acc, qa[1][0], qa[1][1], qa[1][2], qa[1][3], kf[2], kf[3]);
    } else {
        uint32_t kf[2];
        ldsm2(kf, blk + k_off[0]);
        t2::mma16<T>(acc, qa[0][0], qa[0][1], qa[0][2], qa[0][3], kf[0], kf[1]);
    }
}

// out += P[:, octets A|B] . V[octets A|B]  (16-bit: one k16 step = two octets; sA / sB = slot16 entries of the lane's rows, 0xffff = none)
template <typename T, int C, bool PACKED>
__device__ __forceinline__ void pv_pair16(float (&acc)[C / 8][4], uint32_t Sa_u, uint32_t Sb_u, uint32_t sA, uint32_t sB, uint32_t blk,
                                          const uint32_t *v_off) {
    uint32_t af[4];
    { const float2 p = lds_f2(Sa_u + s8(sA, 0) * 32); af[0] = t2::pack_pair<T>(p.x, p.y); }
    { const float2 p = lds_f2(Sb_u + s8(sA, 1) * 32); af[1] = t2::pack_pair<T>(p.x, p.y); }
    { const float2 p = lds_f2(Sa_u + s8(sB, 0) * 32); af[2] = t2::pack_pair<T>(p.x, p.y); }
    { const float2 p = lds_f2(Sb_u + s8(sB, 1) * 32); af[3] = t2::pack_pair<T>(p.x, p.y); }
#pragma unroll
    for (int n2 = 0; n2 < C / 16; ++n2) {
        uint32_t bf[4];
        t2::ldsm4t(bf, blk + v_off[n2]);
        t2::mma16<T>(acc[2 * n2], af[0], af[1], af[2], af[3], bf[0], bf[1]);
        t2::mma16<T>(acc[2 * n2 + 1], af[0], af[1], af[2], af[3], bf[2], bf[3]);
    }
}

// fp32: one k8 step = one octet (k index t <-> key 2t, t + 4 <-> key 2t + 1: see v_off)
template <int C>
__device__ __forceinline__ void pv_octet32(float (&acc)[C / 8][4], uint32_t Sa_u, uint32_t Sb_u, uint32_t s16, uint32_t blk, const uint32_t *v_off) {
    const float2 p0 = lds_f2(Sa_u + s8(s16, 0) * 32), p1 = lds_f2(Sb_u + s8(s16, 1) * 32);
    uint32_t ah[4], al[4];
    split3(__float_as_uint(p0.x), ah[0], al[0]);
    split3(__float_as_uint(p1.x), ah[1], al[1]);
    split3(__float_as_uint(p0.y), ah[2], al[2]);
    split3(__float_as_uint(p1.y), ah[3], al[3]);
#pragma unroll
    for (int n = 0; n < C / 8; ++n) {
        uint32_t b0h, b0l, b1h, b1l;
        split3(lds32(blk + v_off[2 * n]), b0h, b0l);
        split3(lds32(blk + v_off[2 * n + 1]), b1h, b1l);
        mma3(acc[n], ah, al, b0h, b1h, b0l, b1l);
    }
}

template <typename T, int C, bool PACKED, bool PB>
__global__ void __launch_bounds__(GW * 32)
attn_fused_tma_kernel(const __grid_constant__ CUtensorMap mapK, const __grid_constant__ CUtensorMap mapV, const FArgsOf<PB> a,
                      const PackView pk, const GroupView gv, const Launch L) {
    using G = Geo<T, C, PACKED>;
    extern __shared__ unsigned char dyn_raw[];
    const int generic_flag = pk.flags[0];                  // != 0: the pack routes this tensor to the generic kernels (checked below)
    constexpr bool F32 = sizeof(T) == 4;
    constexpr int KS = F32 ? C / 8 : C / 16;               // mma k-steps of the dot phase
    constexpr int NT = C / 8;                              // 8-channel n-tiles of the output
    constexpr int UNR = F32 ? 2 : 4;                       // octets in flight per warp
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int M = a.M, H = a.H, Nq = a.Nq, MP = M + ZPAD + 4;
    const int cap = L.cap;
    const int b = blockIdx.y / H, h = blockIdx.y - b * H;
    const int bg = b * gv.TG + blockIdx.x;
    const int tile = blockIdx.x * GW + warp;
    const bool active = tile < pk.T;
    const int bt = b * pk.T + (active ? tile : pk.T - 1), i0 = tile * TILE_TOK;

    // ---- shared memory: [staged octets][mbarriers][per warp: S, slot table, group positions] ------------------------------------
    const uint32_t base = (smem_u32(dyn_raw) + 1023u) & ~1023u;
    unsigned char *dyn = dyn_raw + (base - smem_u32(dyn_raw));
    const uint32_t stage = base;
    const uint32_t bars = base + (uint32_t)cap * G::BOX;
    unsigned char *wmem = dyn + (size_t)cap * G::BOX + NBAR_MAX * 8 + (size_t)warp * L.warp_bytes;
    float *S = reinterpret_cast<float *>(wmem) + ZPAD;                           // [16][MP]: row = [ZPAD zeros | M logits / e | blank | 1/sum | pad]
    unsigned char *slot_s = wmem + (size_t)16 * MP * 4;                          // [8][SLOT_G_ROW]: slot16 of (row g | row g+8, union entry u)
    unsigned char *spos_s = slot_s + SLOT_G_TILE;                                // [U_MAX]: group position of union entry u
    const int nbar = cap / OCT_PER_BAR;
    // (independent global reads first, so that their latencies overlap: group size, the first round's octet ids, tile context)
    const int GU = gv.grp_u[bg];
    const int *goct = gv.grp_oct + (int64_t)bg * GU_MAX;
    int oct_first[4];                                      // warp 0: octets lane, lane+32, .. of the first load round (list is zero-padded)
#pragma unroll
    for (int x = 0; x < 4; ++x) oct_first[x] = (warp == 0 && lane + 32 * x < min(cap, GU_MAX)) ? __ldg(goct + lane + 32 * x) : 0;
    const int U_ld = pk.tile_u[bt];
    const uint4 imp_ld = __ldg(reinterpret_cast<const uint4 *>(pk.tok_imp + (int64_t)bt * TILE_TOK));
    if (generic_flag) return;
    if (threadIdx.x == 0) {
        for (int k = 0; k < nbar; ++k) mbar_init(bars + 8 * k, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int rounds = (GU + cap - 1) / cap;
    // one load round: octets [r*cap, r*cap + n) of the group's list -> staging positions 0..n-1 (warp 0; every barrier gets exactly
    // one arrival per round, so its parity is the load counter's)
    auto issue_round = [&](int r, const CUtensorMap *map) {
        const int p0 = r * cap, n = min(GU - p0, cap);
        if (lane < nbar) {
            const int cnt = min(max(n - lane * OCT_PER_BAR, 0), OCT_PER_BAR);
            if (cnt > 0) mbar_expect_tx(bars + 8 * lane, (uint32_t)cnt * G::BOX);
            else mbar_arrive(bars + 8 * lane);
        }
        __syncwarp();
        for (int x = lane; x < n; x += 32)
            tma_box_4d(stage + (uint32_t)x * G::BOX, map, bars + 8 * (x / OCT_PER_BAR), 0, __ldg(goct + p0 + x) * 8, h, b);
    };
    if (warp == 0) {                                       // first round of K (or K|V) boxes, octet ids already in registers
        const int n = min(GU, cap);
        if (lane < nbar) {
            const int cnt = min(max(n - lane * OCT_PER_BAR, 0), OCT_PER_BAR);
            if (cnt > 0) mbar_expect_tx(bars + 8 * lane, (uint32_t)cnt * G::BOX);
            else mbar_arrive(bars + 8 * lane);
        }
        __syncwarp();
#pragma unroll
        for (int x = 0; x < 4; ++x)
            if (lane + 32 * x < n) tma_box_4d(stage + (uint32_t)(lane + 32 * x) * G::BOX, &mapK, bars + 8 * ((lane + 32 * x) / OCT_PER_BAR), 0, oct_first[x] * 8, h, b);
    }

    // ---- per-warp tile context -------------------------------------------------------------------------------------------------
    const int U = active ? U_ld : 0;
    const int ra = i0 + g, rb = ra + 8;
    uint32_t impm = 0;
    // phase 2a's operands (bias indices -> bias values, mask bytes) of the first PF quads per lane: fetched now, used after phase 1
    constexpr int PF = PB ? 0 : 6;
    const int nquad = active ? min(TILE_TOK, Nq - i0) * (M >> 2) : 0;
    float4 bias_pf[PF > 0 ? PF : 1];
    uint32_t mask_pf[PF > 0 ? PF : 1];
    if constexpr (PF > 0) {
        const int4 *bi = reinterpret_cast<const int4 *>(a.bias_idx + ((int64_t)b * Nq + i0) * M);
        const uint32_t *mk4 = a.mask ? reinterpret_cast<const uint32_t *>(a.mask + ((int64_t)b * Nq + i0) * M) : nullptr;
        const float *tabh = t2::opaque(a.bias_tab + h);
        int4 bv[PF];
#pragma unroll
        for (int k = 0; k < PF; ++k) {
            const int e = lane + 32 * k;
            bv[k] = e < nquad ? __ldg(bi + e) : make_int4(0, 0, 0, 0);
            mask_pf[k] = (mk4 && e < nquad) ? __ldg(mk4 + e) : 0x01010101u;
        }
#pragma unroll
        for (int k = 0; k < PF; ++k)
            bias_pf[k] = make_float4(__ldg(t2::at(tabh, bv[k].x * H)), __ldg(t2::at(tabh, bv[k].y * H)), __ldg(t2::at(tabh, bv[k].z * H)),
                                     __ldg(t2::at(tabh, bv[k].w * H)));
    }
    if (active) {
        const uint4 v = imp_ld;
        if (v.x | v.y | v.z | v.w) {
            const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int qd = 0; qd < 4; ++qd)
#pragma unroll
                for (int kq = 0; kq < 4; ++kq) impm |= ((w4[qd] >> (8 * kq)) & 1u) << (4 * qd + kq);
        }
        const uint4 *sg = reinterpret_cast<const uint4 *>(gv.slot_g + (int64_t)bt * SLOT_G_TILE);
        for (int x = lane; x < SLOT_G_TILE / 16; x += 32) reinterpret_cast<uint4 *>(slot_s)[x] = __ldg(sg + x);
        if (lane < U_MAX / 4) reinterpret_cast<uint32_t *>(spos_s)[lane] = __ldg(reinterpret_cast<const uint32_t *>(gv.sub_pos + (int64_t)bt * U_MAX) + lane);
        if (lane < 16) *reinterpret_cast<float4 *>(S + lane * MP - ZPAD) = make_float4(0.f, 0.f, 0.f, 0.f), *reinterpret_cast<float4 *>(S + lane * MP - ZPAD + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const uint32_t S_u = smem_u32(S), slot_u = smem_u32(slot_s) + g * SLOT_G_ROW, spos_u = smem_u32(spos_s);
    const uint32_t Sa_u = S_u + (uint32_t)(g * MP + 2 * t) * 4, Sb_u = Sa_u + (uint32_t)(8 * MP) * 4;
    uint32_t ready = 0;                                    // general path: barriers seen complete in the current load round
    int loads = 0;                                         // load rounds issued so far - 1 (CTA-uniform): parity of the barriers
    auto wait_oct = [&](int x) {
        const int k = x / OCT_PER_BAR;
        if (!((ready >> k) & 1u)) { mbar_wait(bars + 8 * k, loads & 1); ready |= 1u << k; }
    };

    // ---- q fragments (straight from global: read once per warp) and the blank logit q . blank_k[h] (aff.py:140) -------------------
    uint32_t qa[KS][4], ql[F32 ? KS : 1][4];
    if (active) {
        const T *Q = t2::opaque(reinterpret_cast<const T *>(a.q) + b * a.q_sb + h * a.q_sh);
        const int oa = min(ra, Nq - 1) * (int)a.q_sn, ob = min(rb, Nq - 1) * (int)a.q_sn;
        const T *bk = reinterpret_cast<const T *>(a.blank_k) + h * C;
        float pa = 0.f, pb = 0.f;
#pragma unroll
        for (int s = 0; s < KS; ++s) {
            if constexpr (F32) {
                const float x0 = __uint_as_float(t2::ldg4(t2::at(Q, oa + 8 * s + t))), x1 = __uint_as_float(t2::ldg4(t2::at(Q, ob + 8 * s + t)));
                const float x2 = __uint_as_float(t2::ldg4(t2::at(Q, oa + 8 * s + t + 4))), x3 = __uint_as_float(t2::ldg4(t2::at(Q, ob + 8 * s + t + 4)));
                const float w0 = __ldg(bk + 8 * s + t), w1 = __ldg(bk + 8 * s + t + 4);
                pa = fmaf(x0, w0, pa); pa = fmaf(x2, w1, pa);
                pb = fmaf(x1, w0, pb); pb = fmaf(x3, w1, pb);
                split3(__float_as_uint(x0), qa[s][0], ql[s][0]);
                split3(__float_as_uint(x1), qa[s][1], ql[s][1]);
                split3(__float_as_uint(x2), qa[s][2], ql[s][2]);
                split3(__float_as_uint(x3), qa[s][3], ql[s][3]);
            } else {
                qa[s][0] = t2::ldg4(t2::at(Q, oa + 16 * s + 2 * t));     qa[s][1] = t2::ldg4(t2::at(Q, ob + 16 * s + 2 * t));
                qa[s][2] = t2::ldg4(t2::at(Q, oa + 16 * s + 8 + 2 * t)); qa[s][3] = t2::ldg4(t2::at(Q, ob + 16 * s + 8 + 2 * t));
                const uint32_t w0 = t2::ldg4(bk + 16 * s + 2 * t), w1 = t2::ldg4(bk + 16 * s + 8 + 2 * t);
                auto lo = [](uint32_t w) { return to_f(*reinterpret_cast<const T *>(&w)); };
                auto hi = [](uint32_t w) { const uint32_t x = w >> 16; return to_f(*reinterpret_cast<const T *>(&x)); };
                pa = fmaf(lo(qa[s][0]), lo(w0), pa); pa = fmaf(hi(qa[s][0]), hi(w0), pa);
                pa = fmaf(lo(qa[s][2]), lo(w1), pa); pa = fmaf(hi(qa[s][2]), hi(w1), pa);
                pb = fmaf(lo(qa[s][1]), lo(w0), pb); pb = fmaf(hi(qa[s][1]), hi(w0), pb);
                pb = fmaf(lo(qa[s][3]), lo(w1), pb); pb = fmaf(hi(qa[s][3]), hi(w1), pb);
            }
        }
        pa += __shfl_xor_sync(FULL, pa, 1); pa += __shfl_xor_sync(FULL, pa, 2);
        pb += __shfl_xor_sync(FULL, pb, 1); pb += __shfl_xor_sync(FULL, pb, 2);
        if (t == 0) { S[g * MP + M] = pa; S[(g + 8) * MP + M] = pb; }
    }
    __syncwarp();
    // lane-constant fragment offsets inside a staged box
    uint32_t k_off[F32 ? 2 * KS : 1];
    if constexpr (F32) {
#pragma unroll
        for (int s = 0; s < KS; ++s) { k_off[2 * s] = G::at(g, (8 * s + t) * 4); k_off[2 * s + 1] = G::at(g, (8 * s + t + 4) * 4); }
    } else {
        k_off[0] = G::at(lane & 7, ((lane >> 3) & (C == 32 ? 3 : 1)) * 16);       // ldmatrix: matrix lane>>3 = 16-byte chunk of the k row
    }
    uint32_t v_off[F32 ? 2 * NT : NT / 2];
    if constexpr (F32) {
        // k index t <-> key 2t, k index t+4 <-> key 2t+1 (any permutation of the summation index is a valid mma): the probabilities
        // of one lane are then adjacent in S and the V words of a warp fall into 32 different banks of the swizzled box
#pragma unroll
        for (int n = 0; n < NT; ++n) {
            v_off[2 * n] = G::at(2 * t, G::V_COL + (8 * n + g) * 4);
            v_off[2 * n + 1] = G::at(2 * t + 1, G::V_COL + (8 * n + g) * 4);
        }
    } else {
        // ldmatrix.trans: matrices (octet A, n), (octet B, n), (octet A, n+1), (octet B, n+1); lane>>3 picks the matrix
#pragma unroll
        for (int n2 = 0; n2 < NT / 2; ++n2) v_off[n2] = G::at(lane & 7, G::V_COL + (2 * n2 + (lane >> 4)) * 16);
    }

    // ---- phase 1: logits of the 16 tokens against every octet of the tile's union (tensor cores), selected 8-column blocks -> S -----
    if (rounds == 1) {
        // fast path (the group's union fits the staging area): UNR octets in flight, barriers taken in list order
        int wmax = -1;
        for (int u0 = 0; u0 < U; u0 += UNR) {
            const uint32_t xw = UNR == 4 ? lds32(spos_u + u0) : (uint32_t)t2::lds16(spos_u + u0);
            uint32_t sw[2];
            if constexpr (UNR == 4) { const uint2 v = lds64(slot_u + 2 * u0); sw[0] = v.x; sw[1] = v.y; }
            else { sw[0] = lds32(slot_u + 2 * u0); sw[1] = 0; }
            int kneed = 0;
#pragma unroll
            for (int j = 0; j < UNR; ++j) kneed = max(kneed, (int)((xw >> (8 * j)) & 0xffu));      // (entries beyond U are 0)
            kneed /= OCT_PER_BAR;
            while (wmax < kneed) { ++wmax; mbar_wait(bars + 8 * wmax, 0); }
            float acc[UNR][4];
#pragma unroll
            for (int j = 0; j < UNR; ++j) qk_octet<T, C, PACKED>(acc[j], qa, ql, stage + ((xw >> (8 * j)) & 0xffu) * G::BOX, k_off);
#pragma unroll
            for (int j = 0; j < UNR; ++j) {
                const int sga = s8(sw[j >> 1], 2 * (j & 1)), sgb = s8(sw[j >> 1], 2 * (j & 1) + 1);
                sts_f2_if(Sa_u + sga * 32, acc[j][0], acc[j][1], sga);
                sts_f2_if(Sb_u + sgb * 32, acc[j][2], acc[j][3], sgb);
            }
        }
    } else {
        for (int r = 0; r < rounds; ++r) {
            if (r > 0) {
                __syncthreads();                           // every warp is done with the previous round's octets
                ++loads; ready = 0;
                if (warp == 0) issue_round(r, &mapK);
            }
            const int rbase = r * cap;
            for (int u = 0; u < U; ++u) {
                const int x = (int)spos_s[u] - rbase;
                if ((unsigned)x >= (unsigned)cap) continue;    // staged in another round
                wait_oct(x);
                float acc[4];
                qk_octet<T, C, PACKED>(acc, qa, ql, stage + (uint32_t)x * G::BOX, k_off);
                const uint32_t s2 = t2::lds16(slot_u + 2 * u);
                const int sga = s8(s2, 0), sgb = s8(s2, 1);
                sts_f2_if(Sa_u + sga * 32, acc[0], acc[1], sga);
                sts_f2_if(Sb_u + sgb * 32, acc[2], acc[3], sgb);
            }
        }
    }
    // unpacked operands: the staging area is refilled with the V boxes of the same octets while the softmax runs
    if (!PACKED && rounds == 1) {
        __syncthreads();
        ++loads; ready = 0;
        if (warp == 0) issue_round(0, &mapV);
    }
    __syncwarp();
    // ---- phase 2a: + bias + mask, four consecutive neighbours per lane (coalesced reads of the tile's bias-index / mask block) ----
    // A masked neighbour (aff.py:137: logit - 100, NOT -inf) is a wildcard of the mask-aware pack: the octet column that stands for
    // it was scored against whatever row the octet holds there (zeros beyond the last key row).  The reference scores it against
    // row idx[b,i,j] (= 0 for the padded tail of the last cluster, point_utils.py:283), and with large logits exp(. - 100) is not
    // nothing: such entries (a few tokens per sample) get their exact logit here and their exact value row before phase 3.
    bool saw_mask = false;
    if (active) {
        const int rows = min(TILE_TOK, Nq - i0), QM = M >> 2;
        const uint8_t *mk = a.mask ? a.mask + ((int64_t)b * Nq + i0) * M : nullptr;
        if constexpr (!PB) {
            const int4 *bi = reinterpret_cast<const int4 *>(a.bias_idx + ((int64_t)b * Nq + i0) * M);
            const float *tabh = t2::opaque(a.bias_tab + h);
            auto apply = [&](int e, float4 gq, uint32_t m4) {
                const int row = (e * L.rcp_qm) >> 20, j = (e - row * QM) * 4;
                float4 x = *reinterpret_cast<float4 *>(S + row * MP + j);
                x.x += gq.x; x.y += gq.y; x.z += gq.z; x.w += gq.w;
                if (m4 != 0x01010101u && ((m4 & 0xffu) == 0 || (m4 & 0xff00u) == 0 || (m4 & 0xff0000u) == 0 || (m4 >> 24) == 0)) {
                    saw_mask = true;                                 // wildcard column: keep the bias only; masked_logit_pass adds the rest
                    if (!(m4 & 0xffu)) x.x = gq.x;
                    if (!(m4 & 0xff00u)) x.y = gq.y;
                    if (!(m4 & 0xff0000u)) x.z = gq.z;
                    if (!(m4 >> 24)) x.w = gq.w;
                }
                *reinterpret_cast<float4 *>(S + row * MP + j) = x;
            };
#pragma unroll
            for (int k = 0; k < PF; ++k)
                if (lane + 32 * k < nquad) apply(lane + 32 * k, bias_pf[k], mask_pf[k]);
            for (int e = lane + 32 * PF; e < nquad; e += 32) {           // (M > 48: the quads beyond the prefetched ones)
                const int4 bv = __ldg(bi + e);
                const float4 gq = make_float4(__ldg(t2::at(tabh, bv.x * H)), __ldg(t2::at(tabh, bv.y * H)), __ldg(t2::at(tabh, bv.z * H)),
                                              __ldg(t2::at(tabh, bv.w * H)));
                apply(e, gq, mk ? __ldg(reinterpret_cast<const uint32_t *>(mk) + e) : 0x01010101u);
            }
        } else {
            const PosBiasW pw = pos_bias_load(a.pe_w, a.pe_b, h);
            const float2 *PQ = reinterpret_cast<const float2 *>(a.pos_q) + (int64_t)b * Nq + i0;
            const float2 *PK = t2::opaque(reinterpret_cast<const float2 *>(a.pos_k) + (int64_t)b * a.Nk);
            const longlong2 *ix = reinterpret_cast<const longlong2 *>(a.idx + ((int64_t)b * Nq + i0) * M);
            const int klast = a.Nk - 1;
            for (int e = lane; e < rows * QM; e += 32) {
                const int row = (e * L.rcp_qm) >> 20, j = (e - row * QM) * 4;
                const longlong2 i01 = __ldg(ix + 2 * e), i23 = __ldg(ix + 2 * e + 1);
                const float2 pq = __ldg(PQ + row);
                const float2 k0 = __ldg(t2::at(PK, min(max((int)i01.x, 0), klast))), k1 = __ldg(t2::at(PK, min(max((int)i01.y, 0), klast)));
                const float2 k2 = __ldg(t2::at(PK, min(max((int)i23.x, 0), klast))), k3 = __ldg(t2::at(PK, min(max((int)i23.y, 0), klast)));
                float4 x = *reinterpret_cast<float4 *>(S + row * MP + j);
                const float g0 = pos_bias(pw, pq, k0), g1 = pos_bias(pw, pq, k1), g2 = pos_bias(pw, pq, k2), g3 = pos_bias(pw, pq, k3);
                x.x += g0; x.y += g1; x.z += g2; x.w += g3;
                if (mk) {
                    const uchar4 m4 = __ldg(reinterpret_cast<const uchar4 *>(mk) + e);
                    if (!(m4.x && m4.y && m4.z && m4.w)) {
                        saw_mask = true;                             // (PB: masked_logit_pass adds the key row's bias too)
                        if (!m4.x) x.x = 0.f;
                        if (!m4.y) x.y = 0.f;
                        if (!m4.z) x.z = 0.f;
                        if (!m4.w) x.w = 0.f;
                    }
                }
                *reinterpret_cast<float4 *>(S + row * MP + j) = x;
            }
        }
    }
    __syncwarp();
    const bool any_mask = __any_sync(FULL, saw_mask);
    if (any_mask) masked_logit_pass<T, PB>(a, b, h, i0, min(TILE_TOK, Nq - i0), impm, S, MP, lane);
    // ---- phase 2b: softmax over M + 1 logits, two lanes per token row; e_j stay unnormalised in S ---------------------------------
    if (active) {
        constexpr float LOG2E = 1.4426950408889634f;
        const int row = lane >> 1, half = lane & 1;
        const int i = i0 + row;
        const bool rvalid = i < Nq && !((impm >> row) & 1u);
        float *Sr = S + row * MP;
        const int Mh = M >> 1, j0 = half * Mh, j1 = j0 + Mh;
        float mx = -INFINITY;
        if (rvalid) {
            for (int j = j0; j < j1; j += 4) {
                const float4 x = *reinterpret_cast<float4 *>(Sr + j);
                mx = fmaxf(fmaxf(mx, fmaxf(x.x, x.y)), fmaxf(x.z, x.w));
            }
            if (half == 0) mx = fmaxf(mx, Sr[M]);
        }
        mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, 1));
        float sum = 0.f;
        if (rvalid) {
            for (int j = j0; j < j1; j += 4) {
                float4 x = *reinterpret_cast<float4 *>(Sr + j);
                x.x = ex2f((x.x - mx) * LOG2E); x.y = ex2f((x.y - mx) * LOG2E); x.z = ex2f((x.z - mx) * LOG2E); x.w = ex2f((x.w - mx) * LOG2E);
                sum += (x.x + x.y) + (x.z + x.w);
                *reinterpret_cast<float4 *>(Sr + j) = x;
            }
            if (half == 0) { const float e = ex2f((Sr[M] - mx) * LOG2E); Sr[M] = e; sum += e; }
        } else {
            for (int j = j0; j < j1; j += 4) *reinterpret_cast<float4 *>(Sr + j) = make_float4(0.f, 0.f, 0.f, 0.f);
            if (half == 0) Sr[M] = 0.f;
        }
        sum += __shfl_xor_sync(FULL, sum, 1);
        if (half == 0) Sr[M + 1] = rvalid ? 1.f / sum : 0.f;
        if (a.lse && half == 0 && rvalid) a.lse[((int64_t)b * H + h) * Nq + i] = mx + logf(sum);
    }
    __syncwarp();
    // ---- phase 3: out = sum_j e_j v_j over the tile's union (tensor cores, V fragments from the staged octets) ---------------------
    float acc[NT][4];
#pragma unroll
    for (int n = 0; n < NT; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;
    if (any_mask) {
        // masked entries: e_j leaves S and enters the accumulators with the row the reference reads, v[idx[b,i,j]] (fused.cuh)
        masked_value_pass<T, PB>(a, b, i0, min(TILE_TOK, Nq - i0), impm, S, MP, lane, [&](int row, int64_t kidx, float e) {
            const T *vr = reinterpret_cast<const T *>(a.v) + b * a.v_sb + h * a.v_sh + kidx * a.v_sn + 2 * t;
            if (g == (row & 7)) {
#pragma unroll
                for (int n = 0; n < NT; ++n) {
                    const float v0 = e * to_f(vr[8 * n]), v1 = e * to_f(vr[8 * n + 1]);
                    if (row < 8) { acc[n][0] += v0; acc[n][1] += v1; }
                    else { acc[n][2] += v0; acc[n][3] += v1; }
                }
            }
        });
    }
    if (rounds == 1) {
        int wmax = PACKED ? NBAR_MAX : -1;                 // (packed: every octet of the union was waited for in phase 1)
        for (int u0 = 0; u0 < U; u0 += 4) {
            const uint32_t xw = lds32(spos_u + u0);
            const uint2 sw = lds64(slot_u + 2 * u0);
            if constexpr (!PACKED) {
                int kneed = max(max((int)(xw & 0xffu), (int)((xw >> 8) & 0xffu)), max((int)((xw >> 16) & 0xffu), (int)(xw >> 24))) / OCT_PER_BAR;
                while (wmax < kneed) { ++wmax; mbar_wait(bars + 8 * wmax, loads & 1); }
            }
            if constexpr (F32) {
#pragma unroll
                for (int j = 0; j < 4; ++j)                // (entries beyond U: no slots -> zero probabilities against staged octet 0)
                    pv_octet32<C>(acc, Sa_u, Sb_u, (j < 2 ? sw.x : sw.y) >> (16 * (j & 1)), stage + ((xw >> (8 * j)) & 0xffu) * G::BOX, v_off);
            } else {
                const uint32_t hiB = (lane >> 3) & 1;      // this lane addresses octet B of the pair in the ldmatrix
                pv_pair16<T, C, PACKED>(acc, Sa_u, Sb_u, sw.x & 0xffffu, sw.x >> 16, stage + ((xw >> (8 * hiB)) & 0xffu) * G::BOX, v_off);
                pv_pair16<T, C, PACKED>(acc, Sa_u, Sb_u, sw.y & 0xffffu, sw.y >> 16, stage + ((xw >> (16 + 8 * hiB)) & 0xffu) * G::BOX, v_off);
            }
        }
    } else {
        for (int r = 0; r < rounds; ++r) {
            __syncthreads();
            ++loads; ready = 0;
            if (warp == 0) issue_round(r, PACKED ? &mapK : &mapV);
            const int rbase = r * cap;
            if constexpr (F32) {
                for (int u = 0; u < U; ++u) {
                    const int x = (int)spos_s[u] - rbase;
                    if ((unsigned)x >= (unsigned)cap) continue;
                    wait_oct(x);
                    pv_octet32<C>(acc, Sa_u, Sb_u, t2::lds16(slot_u + 2 * u), stage + (uint32_t)x * G::BOX, v_off);
                }
            } else {
                int pend_x = -1, pend_u = 0;
                for (int u = 0; u < U; ++u) {
                    const int x = (int)spos_s[u] - rbase;
                    if ((unsigned)x >= (unsigned)cap) continue;
                    wait_oct(x);
                    if (pend_x < 0) { pend_x = x; pend_u = u; continue; }
                    pv_pair16<T, C, PACKED>(acc, Sa_u, Sb_u, t2::lds16(slot_u + 2 * pend_u), t2::lds16(slot_u + 2 * u),
                                            stage + (uint32_t)(((lane >> 3) & 1) ? x : pend_x) * G::BOX, v_off);
                    pend_x = -1;
                }
                // odd count: the B half re-reads octet A with zero probabilities
                if (pend_x >= 0) pv_pair16<T, C, PACKED>(acc, Sa_u, Sb_u, t2::lds16(slot_u + 2 * pend_u), 0xffffu, stage + (uint32_t)pend_x * G::BOX, v_off);
            }
        }
    }
    // ---- epilogue: + e_blank * blank_v, * 1/sum, token-major store -------------------------------------------------------------------
    if (active) {
        const float *Sa = S + g * MP, *Sb = S + (g + 8) * MP;
        const float inva = Sa[M + 1], invb = Sb[M + 1], eba = Sa[M], ebb = Sb[M];
        const T *bv = reinterpret_cast<const T *>(a.blank_v) + h * C;
        T *Ob = t2::opaque(reinterpret_cast<T *>(a.out) + b * a.o_sb + h * a.o_sh);
        T *oa = t2::at(Ob, ra * (int)a.o_sn + 2 * t), *ob = t2::at(Ob, rb * (int)a.o_sn + 2 * t);
        const bool wa = ra < Nq && !((impm >> g) & 1u), wb = rb < Nq && !((impm >> (g + 8)) & 1u);
#pragma unroll
        for (int n = 0; n < NT; ++n) {
            const int ch = 8 * n + 2 * t;
            const float v0 = to_f(bv[ch]), v1 = to_f(bv[ch + 1]);
            if (wa) t2::st_pair<T>(oa + 8 * n, (acc[n][0] + eba * v0) * inva, (acc[n][1] + eba * v1) * inva);
            if (wb) t2::st_pair<T>(ob + 8 * n, (acc[n][2] + ebb * v0) * invb, (acc[n][3] + ebb * v1) * invb);
        }
        if (a.probs) {                                     // normalised probabilities, coalesced (16 rows x (M+1), rows adjacent)
            float *pr = a.probs + (((int64_t)b * H + h) * Nq + i0) * (M + 1);
            const int lim = min(TILE_TOK, Nq - i0) * (M + 1);
            for (int x = lane; x < lim; x += 32) {
                const int row = x / (M + 1), col = x - row * (M + 1);
                if (!((impm >> row) & 1u)) pr[x] = S[row * MP + col] * S[row * MP + M + 1];
            }
        }
    }
    __syncwarp();
    // impure tokens of the whole call, one (token, head) per warp, spread over the grid (t2::slow_items); S row 0 as scratch
    const t2::SlowIter si = t2::slow_items(pk, H);
    for (int it = si.first; it < si.n; it += si.stride) {
        const int gi = pk.imp_list[it / H], hh = it % H;
        const int bb = gi / Nq;
        fused_row_generic<T, PB>(a, bb, hh, gi - bb * Nq, S, lane);
    }
}

// ---- host side --------------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// 4-D map (channels, rows, heads, batch) of a strided [B,H,N,*] operand; box = [row_elems x 8 rows], swizzle span = box row bytes
static bool make_map(CUtensorMap *m, int dtype, const void *ptr, int row_elems, int es, int N, int H, int B, int64_t sn, int64_t sh, int64_t sb) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) return false;
    const CUtensorMapDataType dt = dtype == CLUSTEN_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                   : dtype == CLUSTEN_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    const int rb = row_elems * es;
    const CUtensorMapSwizzle sw = rb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : rb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    const cuuint64_t dims[4] = {(cuuint64_t)row_elems, (cuuint64_t)N, (cuuint64_t)H, (cuuint64_t)B};
    // (a size-1 dimension may carry any stride the caller's view reports: give it a harmless one)
    const cuuint64_t strides[3] = {(cuuint64_t)sn * es, (cuuint64_t)(H > 1 ? sh : sn * N) * es, (cuuint64_t)(B > 1 ? sb : sn * N) * es};
    const cuuint32_t box[4] = {(cuuint32_t)row_elems, 8u, 1u, 1u}, estr[4] = {1u, 1u, 1u, 1u};
    return enc(m, dt, 4, const_cast<void *>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int env_int(const char *name, int dflt) {
    const char *e = getenv(name);
    return e && *e ? atoi(e) : dflt;
}

template <typename T, int C, bool PACKED, bool PB>
static int launch_cfg(const CUtensorMap &mk, const CUtensorMap &mv, const FusedArgsPB &a, const void *pack, cudaStream_t st, bool *taken) {
    using G = Geo<T, C, PACKED>;
    auto kern = attn_fused_tma_kernel<T, C, PACKED, PB>;
    const int warp_bytes = ((16 * (a.M + ZPAD + 4) * 4 + SLOT_G_TILE + 64) + 15) & ~15;
    const size_t fixed = 1024 + NBAR_MAX * 8 + (size_t)GW * warp_bytes;
    // staging capacity: the largest multiple of OCT_PER_BAR that still leaves `want` CTAs per SM, at least the typical union of a
    // 64-token group (~28 octets of 8 rows for M = 48, ~50 for M = 144); larger groups take a second round
    static const int cap_env = env_int("CLUSTEN_TMA_CAP", 0);
    int cap = cap_env;
    if (cap <= 0) {
        const int typical = a.M <= 64 ? 36 : 64;
        cap = 0;
        for (int want = 4; want >= 1 && cap < typical; --want) {
            const int64_t room = (int64_t)(227 * 1024) / want - 1024 - (int64_t)fixed;
            cap = (int)std::min<int64_t>(room / G::BOX, 128);
        }
        cap = std::min(cap, a.M <= 64 ? 40 : 72);
    }
    cap = std::max(OCT_PER_BAR, std::min(cap, OCT_PER_BAR * NBAR_MAX) / OCT_PER_BAR * OCT_PER_BAR);
    const size_t smem = fixed + (size_t)cap * G::BOX;
    if (smem > 227 * 1024) return 0;
    static bool attr_done[64] = {};                        // (per instantiation and device)
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_done[dev & 63]) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) { cudaGetLastError(); return 0; }
        attr_done[dev & 63] = true;
    }
    const PackView pk = pack_view(const_cast<void *>(pack), a.B, a.Nq, a.Nk);
    const GroupView gv = group_view(const_cast<void *>(pack), a.B, a.Nq, a.Nk);
    const dim3 grid(gv.TG, a.B * a.H);
    const Launch L{cap, warp_bytes, ((1 << 20) + a.M / 4 - 1) / (a.M / 4)};
    if constexpr (PB) kern<<<grid, GW * 32, smem, st>>>(mk, mv, a, pk, gv, L);
    else kern<<<grid, GW * 32, smem, st>>>(mk, mv, static_cast<const FusedArgs &>(a), pk, gv, L);
    note_launches(1);
    *taken = true;
    return check_launch("attn_fused_tma");
}

template <typename T, bool PB>
static int launch_t(const FusedArgsPB &a, int dtype, const void *pack, cudaStream_t st, bool *taken) {
    constexpr int es = (int)sizeof(T);
    const int C = a.C;
    if (C != 16 && C != 32) return 0;
    // k and v as the two halves of one row ([B,N,H,2,C] output of the kv Linear): one box per octet when the row fits a swizzle span
    const bool packed = reinterpret_cast<const T *>(a.v) == reinterpret_cast<const T *>(a.k) + C && a.v_sb == a.k_sb && a.v_sh == a.k_sh &&
                        a.v_sn == a.k_sn && 2 * C * es <= 128 && env_int("CLUSTEN_TMA_PACKED", 1) != 0;
    auto ok = [&](const void *p, int64_t sb, int64_t sh, int64_t sn) {
        return (reinterpret_cast<uintptr_t>(p) & 15u) == 0 && (sn * es) % 16 == 0 && (a.H == 1 || (sh * es) % 16 == 0) &&
               (a.B == 1 || (sb * es) % 16 == 0) && sn > 0 && sh >= 0 && sb >= 0;
    };
    if (!ok(a.k, a.k_sb, a.k_sh, a.k_sn) || !ok(a.v, a.v_sb, a.v_sh, a.v_sn)) return 0;
    CUtensorMap mk, mv;
    if (packed) {
        if (!make_map(&mk, dtype, a.k, 2 * C, es, a.Nk, a.H, a.B, a.k_sn, a.k_sh, a.k_sb)) return 0;
        mv = mk;
    } else {
        if (!make_map(&mk, dtype, a.k, C, es, a.Nk, a.H, a.B, a.k_sn, a.k_sh, a.k_sb)) return 0;
        if (!make_map(&mv, dtype, a.v, C, es, a.Nk, a.H, a.B, a.v_sn, a.v_sh, a.v_sb)) return 0;
    }
    if (C == 32) {
        if constexpr (es == 4) return launch_cfg<T, 32, false, PB>(mk, mv, a, pack, st, taken);
        else return packed ? launch_cfg<T, 32, true, PB>(mk, mv, a, pack, st, taken) : launch_cfg<T, 32, false, PB>(mk, mv, a, pack, st, taken);
    }
    return packed ? launch_cfg<T, 16, true, PB>(mk, mv, a, pack, st, taken) : launch_cfg<T, 16, false, PB>(mk, mv, a, pack, st, taken);
}

}  // namespace tma

int fused_tma_launch(const FusedArgsPB &a, bool pos_bias, int dtype, const void *pack, cudaStream_t st, bool *taken) {
    *taken = false;
    // Opt-in (CLUSTEN_TMA_ATTN=1): on the B200 this kernel and the per-warp kernel of clusten_fused.cu finish within +-10 % of each
    // other at every AFF stage shape (profiles/r2_attn_fwd_tma_vs_perwarp.md); end to end the per-warp kernel is ahead by 0.5-2 %,
    // so it stays the default.
    const int enabled = tma::env_int("CLUSTEN_TMA_ATTN", 0);      // (read per call: tests switch it at run time)
    if (!enabled || !pack) return 0;
    const int es = dtype == CLUSTEN_F32 ? 4 : 2;
    const int M = a.M;
    // q / out rows are read and written as 4- or 8-byte pieces; the bias-index / mask blocks of a tile as 16- / 4-byte vectors
    auto al = [&](const void *p, int64_t sb, int64_t sh, int64_t sn) {
        return (reinterpret_cast<uintptr_t>(p) & 7u) == 0 && (sb * es) % 8 == 0 && (sh * es) % 8 == 0 && (sn * es) % 8 == 0;
    };
    if (M % 8 != 0 || M > 256 || a.Nq <= 0 || a.Nk <= 0 || (int64_t)a.B * a.H > 65535) return 0;
    if (!al(a.q, a.q_sb, a.q_sh, a.q_sn) || !al(a.out, a.o_sb, a.o_sh, a.o_sn)) return 0;
    if (!pos_bias && (!a.bias_idx || (reinterpret_cast<uintptr_t>(a.bias_idx) & 15u))) return 0;
    if (a.mask && (reinterpret_cast<uintptr_t>(a.mask) & 3u)) return 0;
    if ((reinterpret_cast<uintptr_t>(a.blank_k) | reinterpret_cast<uintptr_t>(a.blank_v)) & 3u) return 0;
    if (pos_bias && (reinterpret_cast<uintptr_t>(a.idx) & 15u)) return 0;
    switch (dtype) {
        case CLUSTEN_F32: return pos_bias ? tma::launch_t<float, true>(a, dtype, pack, st, taken) : tma::launch_t<float, false>(a, dtype, pack, st, taken);
        case CLUSTEN_F16: return pos_bias ? tma::launch_t<__half, true>(a, dtype, pack, st, taken) : tma::launch_t<__half, false>(a, dtype, pack, st, taken);
        case CLUSTEN_BF16: return pos_bias ? tma::launch_t<__nv_bfloat16, true>(a, dtype, pack, st, taken) : tma::launch_t<__nv_bfloat16, false>(a, dtype, pack, st, taken);
        default: return 0;
    }
}

}  // namespace clusten
