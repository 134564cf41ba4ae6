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
__device__ __forceinline__ void mma3(float (&d)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], uint32_t b0h, uint32_t b1h,
                                     uint32_t b0l, uint32_t b1l) {
    t2::mma_tf32(d, al[0], al[1], al[2], al[3], b0h, b1h);
    t2::mma_tf32(d, ah[0], ah[1], ah[2], ah[3], b0l, b1l);
    t2::mma_tf32(d, ah[0], ah[1], ah[2], ah[3], b0h, b1h);
}

// Geometry of one staged octet.  IB = bytes of one operand row (C channels); PACKED: one box whose rows are [k | v] (2*IB bytes),
// else a K box followed by a V box.  RB = row bytes of a box = its swizzle span (32 / 64 / 128).
template <typename T, int C, bool PACKED> struct Geo {
    static constexpr int IB = C * (int)sizeof(T);
    static constexpr int RB = PACKED ? 2 * IB : IB;
    static_assert(RB == 32 || RB == 64 || RB == 128, "box rows are 32, 64 or 128 bytes");
    static constexpr int BOX = 8 * RB;                     // bytes of one box = period of its swizzle pattern
    static constexpr int BLK = PACKED ? BOX : 2 * BOX;     // bytes of one staged octet (k and v)
    static constexpr int V_BOX = PACKED ? 0 : BOX;         // byte offset of the box holding v
    static constexpr int V_COL = PACKED ? IB : 0;          // byte offset of v inside a box row
    static constexpr int SH = RB == 128 ? 0 : RB == 64 ? 1 : 2;
    // byte offset inside a box of (row r, byte o of the row): 16-byte chunks XOR-ed with the 128-byte line index (CU_TENSOR_MAP_SWIZZLE_*)
    __device__ static constexpr int at(int r, int o) { return r * RB + ((((o >> 4) ^ ((r >> SH) & (RB / 16 - 1))) << 4) | (o & 15)); }
};

struct Launch { int cap, warp_bytes; };                    // staging capacity (octets, multiple of OCT_PER_BAR), per-warp scratch bytes

template <typename T, int C, bool PACKED, bool PB>
__global__ void __launch_bounds__(GW * 32)
attn_fused_tma_kernel(const __grid_constant__ CUtensorMap mapK, const __grid_constant__ CUtensorMap mapV, const FArgsOf<PB> a,
                      const PackView pk, const GroupView gv, const Launch L) {
    using G = Geo<T, C, PACKED>;
    extern __shared__ unsigned char dyn_raw[];
    if (pk.flags[0]) return;                               // the pack routes this tensor to the generic kernels
    constexpr bool F32 = sizeof(T) == 4;
    constexpr int KS = F32 ? C / 8 : C / 16;               // mma k-steps of the dot phase
    constexpr int NT = C / 8;                              // 8-channel n-tiles of the output
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int M = a.M, H = a.H, Nq = a.Nq, MP = M + 4;
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
    const uint32_t bars = base + (uint32_t)cap * G::BLK;
    unsigned char *wmem = dyn + (size_t)cap * G::BLK + NBAR_MAX * 8 + (size_t)warp * L.warp_bytes;
    float *S = reinterpret_cast<float *>(wmem);                                  // [16][MP]: logits / e; [M] blank, [M+1] 1/sum
    unsigned char *slot_s = wmem + (size_t)16 * MP * 4;                          // [U_MAX][16]: slots of rows (g, g+8) adjacent
    unsigned char *spos_s = slot_s + U_MAX * 16;                                 // [U_MAX (64)]: group position of union entry u
    const int nbar = cap / OCT_PER_BAR;
    if (threadIdx.x == 0) {
        for (int k = 0; k < nbar; ++k) mbar_init(bars + 8 * k, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int GU = gv.grp_u[bg];
    const int rounds = (GU + cap - 1) / cap;
    const int *goct = gv.grp_oct + (int64_t)bg * GU_MAX;
    // one load round: octets [r*cap, r*cap + n) of the group's list -> staging positions 0..n-1 (warp 0; every barrier gets exactly
    // one arrival per round, so its parity is the round counter's)
    auto issue_round = [&](int r) {
        const int p0 = r * cap, n = min(GU - p0, cap);
        if (lane < nbar) {
            const int cnt = min(max(n - lane * OCT_PER_BAR, 0), OCT_PER_BAR);
            if (cnt > 0) mbar_expect_tx(bars + 8 * lane, (uint32_t)cnt * G::BLK);
            else mbar_arrive(bars + 8 * lane);
        }
        __syncwarp();
        for (int x = lane; x < n; x += 32) {
            const int row0 = __ldg(goct + p0 + x) * 8;
            const uint32_t bar = bars + 8 * (x / OCT_PER_BAR);
            tma_box_4d(stage + (uint32_t)x * G::BLK, &mapK, bar, 0, row0, h, b);
            if constexpr (!PACKED) tma_box_4d(stage + (uint32_t)x * G::BLK + G::V_BOX, &mapV, bar, 0, row0, h, b);
        }
    };
    if (warp == 0) issue_round(0);

    // ---- per-warp tile context -------------------------------------------------------------------------------------------------
    const int U = active ? pk.tile_u[bt] : 0;
    const int ra = i0 + g, rb = ra + 8;
    uint32_t impm = 0;
    if (active) {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(pk.tok_imp + (int64_t)bt * TILE_TOK));
        if (v.x | v.y | v.z | v.w) {
            const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int qd = 0; qd < 4; ++qd)
#pragma unroll
                for (int kq = 0; kq < 4; ++kq) impm |= ((w4[qd] >> (8 * kq)) & 1u) << (4 * qd + kq);
        }
        // slot table [u][16 rows] -> shared, rows (g, g+8) side by side; group positions of the union entries
        for (int u = lane; u < U_MAX; u += 32) {
            const uint4 s4 = __ldg(reinterpret_cast<const uint4 *>(pk.slot_t + ((int64_t)bt * U_MAX + u) * 16));
            uint4 o;
            o.x = __byte_perm(s4.x, s4.z, 0x5140); o.y = __byte_perm(s4.x, s4.z, 0x7362);
            o.z = __byte_perm(s4.y, s4.w, 0x5140); o.w = __byte_perm(s4.y, s4.w, 0x7362);
            *reinterpret_cast<uint4 *>(slot_s + u * 16) = o;
        }
        if (lane < U_MAX / 4) reinterpret_cast<uint32_t *>(spos_s)[lane] = __ldg(reinterpret_cast<const uint32_t *>(gv.sub_pos + (int64_t)bt * U_MAX) + lane);
    }
    const uint32_t S_u = smem_u32(S), slot_u = smem_u32(slot_s) + 2 * g;
    const T *Q = reinterpret_cast<const T *>(a.q) + b * a.q_sb + h * a.q_sh;
    uint32_t ready = 0;                                    // barriers seen complete in the current round
    int round_no = 0;                                      // load rounds issued so far - 1 (CTA-uniform)
    auto wait_oct = [&](int x) {
        const int k = x / OCT_PER_BAR;
        if (!((ready >> k) & 1u)) { mbar_wait(bars + 8 * k, round_no & 1); ready |= 1u << k; }
    };

    // ---- phase 1: logits of the 16 tokens against every octet of the tile's union (tensor cores), selected blocks -> S ----------
    uint32_t qa[KS][4], ql[F32 ? KS : 1][4];
    if (active) {
        const T *qra = Q + (int64_t)min(ra, Nq - 1) * a.q_sn, *qrb = Q + (int64_t)min(rb, Nq - 1) * a.q_sn;
        const T *bk = reinterpret_cast<const T *>(a.blank_k) + h * C;
        float pa = 0.f, pb = 0.f;                          // blank logit q . blank_k[h] (aff.py:140), this lane's channels
#pragma unroll
        for (int s = 0; s < KS; ++s) {
            if constexpr (F32) {
                const float x0 = __ldg(qra + 8 * s + t), x1 = __ldg(qrb + 8 * s + t), x2 = __ldg(qra + 8 * s + t + 4), x3 = __ldg(qrb + 8 * s + t + 4);
                const float w0 = __ldg(bk + 8 * s + t), w1 = __ldg(bk + 8 * s + t + 4);
                pa = fmaf(x0, w0, pa); pa = fmaf(x2, w1, pa);
                pb = fmaf(x1, w0, pb); pb = fmaf(x3, w1, pb);
                t2::split_tf32(__float_as_uint(x0), qa[s][0], ql[s][0]);
                t2::split_tf32(__float_as_uint(x1), qa[s][1], ql[s][1]);
                t2::split_tf32(__float_as_uint(x2), qa[s][2], ql[s][2]);
                t2::split_tf32(__float_as_uint(x3), qa[s][3], ql[s][3]);
            } else {
                qa[s][0] = t2::ldg4(qra + 16 * s + 2 * t);     qa[s][1] = t2::ldg4(qrb + 16 * s + 2 * t);
                qa[s][2] = t2::ldg4(qra + 16 * s + 8 + 2 * t); qa[s][3] = t2::ldg4(qrb + 16 * s + 8 + 2 * t);
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
    // lane-constant fragment offsets inside a staged octet
    uint32_t k_off[F32 ? 2 * KS : 1];
    if constexpr (F32) {
#pragma unroll
        for (int s = 0; s < KS; ++s) { k_off[2 * s] = G::at(g, (8 * s + t) * 4); k_off[2 * s + 1] = G::at(g, (8 * s + t + 4) * 4); }
    } else {
        k_off[0] = G::at(lane & 7, ((lane >> 3) & (C == 32 ? 3 : 1)) * 16);       // ldmatrix: matrix lane>>3 = 16-byte chunk of the k row
    }
    const uint32_t Sa_u = S_u + (uint32_t)(g * MP + 2 * t) * 4, Sb_u = Sa_u + (uint32_t)(8 * MP) * 4;
    for (int r = 0; r < rounds; ++r) {
        if (r > 0) {
            __syncthreads();                               // every warp is done with the previous round's octets
            ++round_no; ready = 0;
            if (warp == 0) issue_round(r);
        }
        const int rbase = r * cap;
        for (int u = 0; u < U; ++u) {
            const int x = (int)spos_s[u] - rbase;
            if ((unsigned)x >= (unsigned)cap) continue;    // staged in another round
            wait_oct(x);
            const uint32_t blk = stage + (uint32_t)x * G::BLK;
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            if constexpr (F32) {
#pragma unroll
                for (int s = 0; s < KS; ++s) {
                    uint32_t b0h, b0l, b1h, b1l;
                    t2::split_tf32(t2::lds32(blk + k_off[2 * s]), b0h, b0l);
                    t2::split_tf32(t2::lds32(blk + k_off[2 * s + 1]), b1h, b1l);
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
            const uint32_t s2 = t2::lds16(slot_u + u * 16);
            const int sga = (int)(int8_t)(s2 & 0xffu), sgb = (int)(int8_t)(s2 >> 8);
            if (sga >= 0) *reinterpret_cast<float2 *>(S + g * MP + 8 * sga + 2 * t) = make_float2(acc[0], acc[1]);
            if (sgb >= 0) *reinterpret_cast<float2 *>(S + (g + 8) * MP + 8 * sgb + 2 * t) = make_float2(acc[2], acc[3]);
        }
    }
    __syncwarp();
    // ---- phase 2a: + bias + mask, four consecutive neighbours per lane (coalesced reads of the tile's bias-index / mask block) ----
    if (active) {
        const int rows = min(TILE_TOK, Nq - i0), QM = M >> 2;
        if constexpr (!PB) {
            const int32_t *bi = a.bias_idx + ((int64_t)b * Nq + i0) * M;
            const uint8_t *mk = a.mask ? a.mask + ((int64_t)b * Nq + i0) * M : nullptr;
            for (int e = lane; e < rows * QM; e += 32) {
                const int row = e / QM, j = (e - row * QM) * 4;
                const int4 bv = __ldg(reinterpret_cast<const int4 *>(bi) + e);
                float4 x = *reinterpret_cast<float4 *>(S + row * MP + j);
                x.x += __ldg(a.bias_tab + bv.x * H + h);
                x.y += __ldg(a.bias_tab + bv.y * H + h);
                x.z += __ldg(a.bias_tab + bv.z * H + h);
                x.w += __ldg(a.bias_tab + bv.w * H + h);
                if (mk) {
                    const uchar4 m4 = __ldg(reinterpret_cast<const uchar4 *>(mk) + e);
                    if (!m4.x) x.x += -100.f;
                    if (!m4.y) x.y += -100.f;
                    if (!m4.z) x.z += -100.f;
                    if (!m4.w) x.w += -100.f;
                }
                *reinterpret_cast<float4 *>(S + row * MP + j) = x;
            }
        } else {
            const PosBiasW pw = pos_bias_load(a.pe_w, a.pe_b, h);
            const float2 *PQ = reinterpret_cast<const float2 *>(a.pos_q) + (int64_t)b * Nq + i0;
            const float2 *PK = reinterpret_cast<const float2 *>(a.pos_k) + (int64_t)b * a.Nk;
            const int64_t *ix = a.idx + ((int64_t)b * Nq + i0) * M;
            const uint8_t *mk = a.mask ? a.mask + ((int64_t)b * Nq + i0) * M : nullptr;
            const int klast = a.Nk - 1;
            for (int e = lane; e < rows * QM; e += 32) {
                const int row = e / QM, j = (e - row * QM) * 4;
                // (a slot is 8 consecutive key rows: the first index of the quad gives the other three; masked entries are wildcards
                // of the mask-aware pack -- their -100 swamps whatever bias the clamped row yields)
                const longlong2 i01 = __ldg(reinterpret_cast<const longlong2 *>(ix) + 2 * e);
                const longlong2 i23 = __ldg(reinterpret_cast<const longlong2 *>(ix) + 2 * e + 1);
                const float2 pq = __ldg(PQ + row);
                float4 x = *reinterpret_cast<float4 *>(S + row * MP + j);
                x.x += pos_bias(pw, pq, __ldg(PK + min(max((int)i01.x, 0), klast)));
                x.y += pos_bias(pw, pq, __ldg(PK + min(max((int)i01.y, 0), klast)));
                x.z += pos_bias(pw, pq, __ldg(PK + min(max((int)i23.x, 0), klast)));
                x.w += pos_bias(pw, pq, __ldg(PK + min(max((int)i23.y, 0), klast)));
                if (mk) {
                    const uchar4 m4 = __ldg(reinterpret_cast<const uchar4 *>(mk) + e);
                    if (!m4.x) x.x += -100.f;
                    if (!m4.y) x.y += -100.f;
                    if (!m4.z) x.z += -100.f;
                    if (!m4.w) x.w += -100.f;
                }
                *reinterpret_cast<float4 *>(S + row * MP + j) = x;
            }
        }
    }
    __syncwarp();
    // ---- phase 2b: softmax over M + 1 logits, two lanes per token row; e_j stay unnormalised in S ---------------------------------
    if (active) {
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
                x.x = f_exp<T>(x.x - mx); x.y = f_exp<T>(x.y - mx); x.z = f_exp<T>(x.z - mx); x.w = f_exp<T>(x.w - mx);
                sum += (x.x + x.y) + (x.z + x.w);
                *reinterpret_cast<float4 *>(Sr + j) = x;
            }
            if (half == 0) { const float e = f_exp<T>(Sr[M] - mx); Sr[M] = e; sum += e; }
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
    uint32_t v_off[F32 ? 2 * NT : (NT + 1) / 2];
    if constexpr (F32) {
        // k index t <-> key 2t, k index t+4 <-> key 2t+1 (any permutation of the summation index is a valid mma): the probabilities
        // of one lane are then adjacent in S and the V words of a warp fall into 32 different banks of the swizzled box
#pragma unroll
        for (int n = 0; n < NT; ++n) {
            v_off[2 * n] = G::V_BOX + G::at(2 * t, G::V_COL + (8 * n + g) * 4);
            v_off[2 * n + 1] = G::V_BOX + G::at(2 * t + 1, G::V_COL + (8 * n + g) * 4);
        }
    } else {
        // ldmatrix.trans: matrices (octet A, n), (octet B, n), (octet A, n+1), (octet B, n+1); lane>>3 picks the matrix
#pragma unroll
        for (int n2 = 0; n2 < (NT + 1) / 2; ++n2) v_off[n2] = G::V_BOX + G::at(lane & 7, G::V_COL + (2 * n2 + (lane >> 4)) * 16);
    }
    const int reload = rounds > 1;
    for (int r = 0; r < rounds; ++r) {
        if (reload) {
            __syncthreads();
            ++round_no; ready = 0;
            if (warp == 0) issue_round(r);
        }
        const int rbase = r * cap;
        if constexpr (F32) {
            for (int u = 0; u < U; ++u) {
                const int x = (int)spos_s[u] - rbase;
                if ((unsigned)x >= (unsigned)cap) continue;
                wait_oct(x);
                const uint32_t blk = stage + (uint32_t)x * G::BLK;
                const uint32_t s2 = t2::lds16(slot_u + u * 16);
                const int s0 = (int)(int8_t)(s2 & 0xffu), s1 = (int)(int8_t)(s2 >> 8);
                float2 p0 = make_float2(0.f, 0.f), p1 = p0;
                if (s0 >= 0) p0 = lds_f2(Sa_u + (uint32_t)s0 * 32);
                if (s1 >= 0) p1 = lds_f2(Sb_u + (uint32_t)s1 * 32);
                uint32_t ah[4], al[4];
                t2::split_tf32(__float_as_uint(p0.x), ah[0], al[0]);
                t2::split_tf32(__float_as_uint(p1.x), ah[1], al[1]);
                t2::split_tf32(__float_as_uint(p0.y), ah[2], al[2]);
                t2::split_tf32(__float_as_uint(p1.y), ah[3], al[3]);
#pragma unroll
                for (int n = 0; n < NT; ++n) {
                    uint32_t b0h, b0l, b1h, b1l;
                    t2::split_tf32(t2::lds32(blk + v_off[2 * n]), b0h, b0l);
                    t2::split_tf32(t2::lds32(blk + v_off[2 * n + 1]), b1h, b1l);
                    mma3(acc[n], ah, al, b0h, b1h, b0l, b1l);
                }
            }
        } else {
            int pend_x = -1, pend_u = 0;
            auto pair = [&](int uA, int xA, int uB, int xB) {      // octets A (keys 0..7 of the k16 step) and B (8..15; uB < 0: none)
                const uint32_t sA = t2::lds16(slot_u + uA * 16);
                const uint32_t sB = uB >= 0 ? t2::lds16(slot_u + uB * 16) : 0xffffu;
                const int s00 = (int)(int8_t)(sA & 0xffu), s10 = (int)(int8_t)(sA >> 8);
                const int s01 = (int)(int8_t)(sB & 0xffu), s11 = (int)(int8_t)(sB >> 8);
                uint32_t af[4] = {0u, 0u, 0u, 0u};
                if (s00 >= 0) { const float2 p = lds_f2(Sa_u + (uint32_t)s00 * 32); af[0] = t2::pack_pair<T>(p.x, p.y); }
                if (s10 >= 0) { const float2 p = lds_f2(Sb_u + (uint32_t)s10 * 32); af[1] = t2::pack_pair<T>(p.x, p.y); }
                if (s01 >= 0) { const float2 p = lds_f2(Sa_u + (uint32_t)s01 * 32); af[2] = t2::pack_pair<T>(p.x, p.y); }
                if (s11 >= 0) { const float2 p = lds_f2(Sb_u + (uint32_t)s11 * 32); af[3] = t2::pack_pair<T>(p.x, p.y); }
                const uint32_t blk = stage + (uint32_t)(((lane >> 3) & 1) ? xB : xA) * G::BLK;
                if constexpr (NT == 4) {
#pragma unroll
                    for (int n2 = 0; n2 < 2; ++n2) {
                        uint32_t bf[4];
                        t2::ldsm4t(bf, blk + v_off[n2]);
                        t2::mma16<T>(acc[2 * n2], af[0], af[1], af[2], af[3], bf[0], bf[1]);
                        t2::mma16<T>(acc[2 * n2 + 1], af[0], af[1], af[2], af[3], bf[2], bf[3]);
                    }
                } else {
                    uint32_t bf[4];
                    t2::ldsm4t(bf, blk + v_off[0]);
                    t2::mma16<T>(acc[0], af[0], af[1], af[2], af[3], bf[0], bf[1]);
                    t2::mma16<T>(acc[1], af[0], af[1], af[2], af[3], bf[2], bf[3]);
                }
            };
            for (int u = 0; u < U; ++u) {
                const int x = (int)spos_s[u] - rbase;
                if ((unsigned)x >= (unsigned)cap) continue;
                wait_oct(x);
                if (pend_x < 0) { pend_x = x; pend_u = u; continue; }
                pair(pend_u, pend_x, u, x);
                pend_x = -1;
            }
            if (pend_x >= 0) pair(pend_u, pend_x, -1, pend_x);     // odd count: the B half re-reads octet A with zero probabilities
        }
    }
    // ---- epilogue: + e_blank * blank_v, * 1/sum, token-major store -------------------------------------------------------------------
    if (active) {
        const float *Sa = S + g * MP, *Sb = S + (g + 8) * MP;
        const float inva = Sa[M + 1], invb = Sb[M + 1], eba = Sa[M], ebb = Sb[M];
        const T *bv = reinterpret_cast<const T *>(a.blank_v) + h * C;
        T *Ob = reinterpret_cast<T *>(a.out) + b * a.o_sb + h * a.o_sh;
        T *oa = Ob + (int64_t)ra * a.o_sn + 2 * t, *ob = Ob + (int64_t)rb * a.o_sn + 2 * t;
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
    const int warp_bytes = ((16 * (a.M + 4) * 4 + U_MAX * 16 + 64) + 15) & ~15;
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
            cap = (int)std::min<int64_t>(room / G::BLK, 128);
        }
        cap = std::min(cap, a.M <= 64 ? 40 : 72);
    }
    cap = std::max(OCT_PER_BAR, std::min(cap, OCT_PER_BAR * NBAR_MAX) / OCT_PER_BAR * OCT_PER_BAR);
    const size_t smem = fixed + (size_t)cap * G::BLK;
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
    const Launch L{cap, warp_bytes};
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
    static const int enabled = tma::env_int("CLUSTEN_TMA_ATTN", 1);
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
