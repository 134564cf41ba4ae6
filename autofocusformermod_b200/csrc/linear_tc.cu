// fp32 Linear layer  Y[R,N] = epilogue( [LayerNorm] X[R,K] . W[N,K]^T + bias[N] )  on the 5th-generation tensor cores (tcgen05.mma,
// accumulators in tensor memory), fp32-level accuracy from a hi / lo split of both operands:
//     x = xh + xl,  w = wh + wl   (xh, wh = the value rounded to an 11-bit significand, xl, wl = the rounded remainder)
//     x.w ~ xl.wh + xh.wl + xh.wh                                   (the dropped xl.wl term is ~2^-22 of the product)
// in two forms: TF32 halves (kind::tf32, K = 8 per MMA; any input range) or fp16 halves (kind::f16, K = 16 per MMA at the same 64
// cycles: half the MMAs, half the weight bytes; weights scaled per output row by a power of two, activations must fit fp16).
// These are the q / kv / proj / fc1 / fc2 layers of the block (backbone/aff.py:62-70,103-106,181-189) and the merge Linear
// (aff.py:282) in fp32 inference, where cuBLAS answers the skinny shapes (K = 32..1536, N = 32..2304, R = 4 096..2 097 152) with
// SIMT sgemm kernels plus a separate bias kernel: 64 % of the device time of the AFF-Mini forward
// (profiles/r2_launches_default.md).  DESIGN.md section 5b has the history, the measurements and what bounds the kernel.
//
// One persistent CTA per SM, 15 warps, five roles (the canonical sm_100 pipeline; every hand-off is an mbarrier):
//   warp 0    X producer: per 32-wide K chunk one TMA box of X [128 rows x 128 B], SWIZZLE_128B, into a 4-8 stage ring (X comes from
//             HBM: this ring covers its latency); the loop runs warp-uniform, an elect.sync lane issues
//   warp 14   W producer: the matching boxes of the pre-split weights Wh, Wl [BN rows], into a 3-4 stage ring (L2-resident operand)
//   warps 2-5 split: thread = row; 8 swizzled LDS.128, optional LayerNorm, hi / lo, tcgen05.st into a 4-stage A ring IN TENSOR MEMORY,
//             fence.proxy.async before the X slot is released
//   warp 1    issuer: 12 (TF32) or 6 (fp16) tcgen05.mma (M = 128, N = BN; A from TMEM, B by shared-memory descriptor) per chunk --
//             small terms first -- into one of two TMEM accumulators; tcgen05.commit frees the W slot and the A stage and, at the end
//             of a chain, hands the accumulator to the epilogue
//   warps 6-13 drain + epilogue (two warps per TMEM lane quadrant, half the columns each): tcgen05.ld, fp32 running sum in registers
//             (the tensor core accumulates with truncation, so a chain is cut after `chain` chunks), then weight-scale / bias /
//             q-scale / GELU / residual into a swizzled staging tile (the residual tile is TMA-loaded into it ahead), TMA store
// The rings and the accumulator pair run across tile boundaries, so the next tile's loads, split and MMAs overlap the epilogue.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace clusten {

namespace tc {

constexpr int BM = 128, BK = 32;                       // rows per tile, K elements per chunk (= one 128-byte swizzle row)
constexpr int THREADS = 480;                         // 15 warps: X producer, MMA issuer, 4 x split, 8 x epilogue, W producer
constexpr int EPI_THREADS = 256;
constexpr int A_BYTES = BM * BK * 4;                   // 16 KiB

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
// (a lost completion must not hang the device: after ~2^24 probes the CTA traps and the launch reports an error)
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin)
        if (spin > (1u << 24)) __trap();
}
// one lane of a converged warp (elect.sync): the warp runs the role's loop uniformly, so the operands of the instructions issued under
// this predicate stay in uniform registers (a loop entered by `lane == 0` alone makes ptxas wrap every UTCHMMA / UTMALDG in a
// per-lane waterfall, ~100 cycles each)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tma_box_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
// shared-memory matrix descriptor, K-major, SWIZZLE_128B: rows of 128 bytes, 8-row groups 1024 bytes apart (SBO), version 1
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3fffu) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
}
// {lo half = fp16(a), hi half = fp16(b)}, round to nearest even, values beyond the fp16 range -> +-65504
__device__ __forceinline__ uint32_t f16x2_sat(float a, float b) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
// round to TF32 (nearest, ties away -- cvt.rna) on the bit pattern: the low 13 mantissa bits end up zero
__device__ __forceinline__ float tf32_rn(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }

// D[tmem] (+)= A[tmem] . B[smem]^T; H: fp16 operands (kind::f16, K = 16 per instruction), else TF32 (kind::tf32, K = 8)
template <bool H>
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if (H)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
                     ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
                     ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// K-major SWIZZLE_64B descriptor (rows of 64 bytes, 8-row groups 512 bytes apart): the fp16 weight boxes [BN x 32 halves]
__device__ __forceinline__ uint64_t umma_desc64(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3fffu) | (1ull << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
                   "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
                 "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};\n"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
                   "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]),
                   "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]),
                   "r"(v[30]), "r"(v[31]) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void epi_bar() {
    __syncwarp();
    asm volatile("bar.sync 1, 256;" ::: "memory");
}     // the eight epilogue warps
// byte offset of 16-byte chunk c of row r inside a [rows x 128 B] box written / read by the TMA unit with SWIZZLE_128B
__device__ __forceinline__ uint32_t swz(int r, int c) { return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); }

constexpr int NA_MAX = 8;                              // A-operand stages in tensor memory: 4 x 64 columns (TF32) or 8 x 32 columns (fp16)
constexpr int A_COL0 = 256;                            // tensor-memory columns 0..255: the two accumulators, 256..511: the A ring
constexpr int SUB_BYTES = BM * 128;                    // one [128 rows x 32 columns] piece of the output tile, 16 KiB

// H: the operands are split into fp16 hi / lo (11 + 11 significand bits, as the TF32 split) and multiplied with kind::f16 -- K = 16 per
// instruction at the same 64 cycles, so half the MMAs, half the weight bytes and half the TMEM columns of the TF32 form.
template <int BN, bool H> struct Cfg {
    static constexpr int B_BYTES = BN * BK * (H ? 2 : 4);
    static constexpr int WSTAGE = 2 * B_BYTES;                           // wh | wl
    static constexpr int WS = BN >= 96 && !H ? 3 : 4;                    // weight ring (L2-resident operand: short latency)
    static constexpr int NSUB = BN / 32;                                 // output pieces per tile
    static constexpr int CSETS = BN >= 96 ? 1 : 2;                       // tile-sized output staging sets
    static constexpr int CBUF = CSETS * NSUB * SUB_BYTES;
    static constexpr int TAIL = 2560;                                    // barriers, TMEM slot, bias / gamma / weight scale of the tile
    // the X ring takes what is left of 227 KiB, at most 8 stages: X comes from HBM and its latency is what the ring has to cover
    static constexpr int XS_FIT = (232448 - 1024 - TAIL - CBUF - WS * WSTAGE) / A_BYTES;
    static constexpr int XS = XS_FIT > 8 ? 8 : XS_FIT;
    static constexpr int ACC_COLS = BN == 96 ? 128 : BN;                 // column pitch of the two accumulators
    static constexpr int A_COLS = H ? 32 : 64;                           // TMEM columns of one A stage: xh | xl
    static constexpr int NA = 256 / A_COLS;                              // A stages: tensor-memory columns 256..511
    static constexpr int KSTEPS = H ? BK / 16 : BK / 8;                  // MMA K steps per chunk (8 TMEM columns / 32 B of B each)
    static constexpr uint32_t IDESC = (1u << 4) | ((H ? 0u : 2u) << 7) | ((H ? 0u : 2u) << 10) | ((uint32_t)(BN >> 3) << 17) |
                                      ((uint32_t)(BM >> 4) << 24);
    static constexpr int XRING = XS * A_BYTES, RING = XRING + WS * WSTAGE;
    static constexpr size_t SMEM = (size_t)RING + CBUF + TAIL + 1024 /* alignment slack */;
    static_assert(XS >= 3, "X ring too shallow");
};

enum { EPI_BIAS = 0, EPI_GELU = 1, EPI_RES = 2 };

// -DCLUSTEN_TC_PROFILE: cycle counters of the issuer warp and of the first epilogue warp, per CTA (tools/lin_profile.py reads them
// through clusten_linear_tc_profile).  Not compiled into the shipped library.
#ifdef CLUSTEN_TC_PROFILE
__device__ long long tc_prof[148 * 16];
#define TC_CLK(v) const long long v = clock64()
#define TC_ADD(slot, a_, b_) prof[slot] += (b_) - (a_)
#else
#define TC_CLK(v)
#define TC_ADD(slot, a_, b_)
#endif

struct Args {
    const float *bias, *gamma;
    const float *ln_mean, *ln_rstd, *ln_gamma, *ln_beta;   // LayerNorm of the X rows applied while they are split (or NULL)
    int R, K, N;
    int tiles_m, tiles_n, chain;
    float alpha;
    int alpha_cols;
    const float *w_inv_scale;                              // fp16 split: 1 / (power-of-two scale of weight row n), [N]
    int resident;                                          // 1: a CTA owns ROW tiles and walks their column tiles with the split rows kept in tensor memory
    int dbg;                                               // experiment switches (CLUSTEN_TC_DBG; wrong results): 1 = no X loads, 2 = no W loads, 4 = no split, 8 = no epilogue
};

template <int BN, int EPI, bool H>
__global__ void __launch_bounds__(THREADS, 1)
linear_tc_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapWh,
                 const __grid_constant__ CUtensorMap mapWl, const __grid_constant__ CUtensorMap mapY,
                 const __grid_constant__ CUtensorMap mapRes, const Args a) {
    using C = Cfg<BN, H>;
    constexpr int XS = C::XS, WS = C::WS, NA = C::NA;
    extern __shared__ uint8_t tc_smem_raw[];
    const uint32_t base = (smem_u32(tc_smem_raw) + 1023u) & ~1023u;
    uint8_t *gbase = tc_smem_raw + (base - smem_u32(tc_smem_raw));
    const uint32_t wring = base + C::XRING, cbuf = base + C::RING, tail = cbuf + C::CBUF;
    const uint32_t bar_xfull = tail, bar_xempty = tail + 8 * XS, bar_wfull = tail + 16 * XS, bar_wempty = bar_wfull + 8 * WS;
    const uint32_t bar_afull = bar_wempty + 8 * WS, bar_aempty = bar_afull + 8 * NA_MAX, bar_accf = bar_aempty + 8 * NA_MAX, bar_acce = bar_accf + 16;
    const uint32_t bar_cfull = bar_acce + 16;                            // CSETS * NSUB barriers
    uint8_t *gtail = gbase + C::RING + C::CBUF;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(gtail + 512);
    float *s_bias = reinterpret_cast<float *>(gtail + 1024), *s_gamma = s_bias + 128, *s_wsc = s_bias + 256;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < XS; ++s) {
            mbar_init(bar_xfull + 8 * s, 1);
            mbar_init(bar_xempty + 8 * s, 128);
        }
        for (int s = 0; s < WS; ++s) {
            mbar_init(bar_wfull + 8 * s, 1);
            mbar_init(bar_wempty + 8 * s, 1);
        }
        for (int s = 0; s < NA; ++s) {
            mbar_init(bar_afull + 8 * s, 128);
            mbar_init(bar_aempty + 8 * s, 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(bar_accf + 8 * b, 1);
            mbar_init(bar_acce + 8 * b, EPI_THREADS);
        }
        for (int j = 0; j < C::CSETS * C::NSUB; ++j) mbar_init(bar_cfull + 8 * j, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;

    const int KC = a.K / BK, tiles = a.tiles_m * a.tiles_n;
    const int chain = a.chain;
    // Tile walk of this CTA.  Default: tiles t = blockIdx.x, + gridDim.x, ... with the column tile running fastest across CTAs; every
    // tile loads and splits its X rows.  Resident (K <= NA chunks, more than one column tile): the CTA owns ROW tiles o = blockIdx.x,
    // + gridDim.x, ... and walks their column tiles i = 0 .. tiles_n - 1 itself; the rows are loaded and split ONCE, stay in their
    // tensor-memory stages for all column tiles, and only the weights stream (what an SM ingests per chunk halves, section 5b).
    const bool res = a.resident != 0;
    const int outer_n = res ? a.tiles_m : tiles, inner_n = res ? a.tiles_n : 1;

    if (warp == 0) {                                                     // ---- TMA producer, X (from HBM: the deep ring) ----
        uint32_t s = 0, ph = 0;
        for (int o = blockIdx.x; o < outer_n; o += gridDim.x) {
            const int m0 = (res ? o : o / a.tiles_n) * BM;
            for (int kc = 0; kc < KC; ++kc) {
                mbar_wait(bar_xempty + 8 * s, ph ^ 1u);
                if (elect_one()) {
                    if (a.dbg & 1) mbar_arrive(bar_xfull + 8 * s);
                    else {
                        mbar_expect_tx(bar_xfull + 8 * s, A_BYTES);
                        tma_box_2d(base + s * A_BYTES, &mapX, bar_xfull + 8 * s, kc * BK, m0);
                    }
                }
                __syncwarp();
                if (++s == XS) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 14) {                                             // ---- TMA producer, pre-split weights (L2) ----
        uint32_t s = 0, ph = 0;
        for (int o = blockIdx.x; o < outer_n; o += gridDim.x)
        for (int i = 0; i < inner_n; ++i) {
            const int n0 = (res ? i : o % a.tiles_n) * BN;
            for (int kc = 0; kc < KC; ++kc) {
                const uint32_t st = wring + s * C::WSTAGE;
                mbar_wait(bar_wempty + 8 * s, ph ^ 1u);
                if (elect_one()) {
                    if (a.dbg & 2) mbar_arrive(bar_wfull + 8 * s);
                    else {
                        mbar_expect_tx(bar_wfull + 8 * s, C::WSTAGE);
                        tma_box_2d(st, &mapWh, bar_wfull + 8 * s, kc * BK, n0);
                        tma_box_2d(st + C::B_BYTES, &mapWl, bar_wfull + 8 * s, kc * BK, n0);
                    }
                }
                __syncwarp();
                if (++s == WS) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {                                              // ---- MMA issuer ----
        // This warp's loop is the clock of the kernel: ncu showed it ~100 % busy at ~1100 cycles per chunk with the tensor pipe idle
        // in between (ring slots by % and /, `kc % chain` by a runtime divisor, descriptors rebuilt from addresses, every chunk).
        // Ring positions and parities are now carried, descriptors are a base plus a stride, the chain position is a counter.
        const uint64_t wdesc0 = H ? umma_desc64(wring) : umma_desc(wring);
        constexpr uint32_t WSTEP = C::WSTAGE >> 4, WLO = C::B_BYTES >> 4;            // descriptor address units (16 bytes)
        const uint32_t xa0 = tmem + A_COL0;
        uint32_t s = 0, ph = 0, sa = 0, pa = 0, buf = 0, pacc = 1;        // W slot / parity, A stage / parity, accumulator / its "drained" parity
#ifdef CLUSTEN_TC_PROFILE
        long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const long long prof_t0 = clock64();
#endif
        for (int o = blockIdx.x; o < outer_n; o += gridDim.x) {
            const uint32_t sa0 = sa, pa0 = pa;                           // resident: every column tile revisits the A stages of its row tile
        for (int i = 0; i < inner_n; ++i) {
            sa = sa0; pa = pa0;
            const bool free_a = i == inner_n - 1;                        // ... and the last one hands them back to the split warps
            int cc = 0;                                                  // chunk inside the current accumulator chain
            for (int kc = 0; kc < KC; ++kc) {
                const bool first = cc == 0, last = cc == chain - 1 || kc == KC - 1;
                TC_CLK(c0);
                if (first) mbar_wait(bar_acce + 8 * buf, pacc);          // the epilogue has drained this accumulator
                TC_CLK(c1);
                mbar_wait(bar_wfull + 8 * s, ph);
                TC_CLK(c2);
                mbar_wait(bar_afull + 8 * sa, pa);
                TC_CLK(c3);
                TC_ADD(0, c0, c1); TC_ADD(1, c1, c2); TC_ADD(2, c2, c3);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (elect_one()) {
                    const uint64_t wh = wdesc0 + s * WSTEP, wl = wh + WLO;
                    const uint32_t xh = xa0 + sa * C::A_COLS, xl = xh + C::A_COLS / 2;
                    const uint32_t d = tmem + buf * C::ACC_COLS;
#pragma unroll
                    for (int k = 0; k < C::KSTEPS; ++k)                  // one K step: 8 TMEM columns of A, 32 bytes (+2) of B
                        umma_ts<H>(d, xl + 8 * k, wh + 2 * k, C::IDESC, (first && k == 0) ? 0u : 1u);
#pragma unroll
                    for (int k = 0; k < C::KSTEPS; ++k) umma_ts<H>(d, xh + 8 * k, wl + 2 * k, C::IDESC, 1u);
#pragma unroll
                    for (int k = 0; k < C::KSTEPS; ++k) umma_ts<H>(d, xh + 8 * k, wh + 2 * k, C::IDESC, 1u);
                    umma_commit(bar_wempty + 8 * s);                     // both rings are free once these MMAs have read them
                    if (free_a) umma_commit(bar_aempty + 8 * sa);
                    if (last) umma_commit(bar_accf + 8 * buf);
                }
                __syncwarp();
                TC_CLK(c4);
                TC_ADD(3, c3, c4);
#ifdef CLUSTEN_TC_PROFILE
                prof[4] += 1;
#endif
                if (++s == WS) { s = 0; ph ^= 1u; }
                if (++sa == NA) { sa = 0; pa ^= 1u; }
                ++cc;
                if (last) {
                    cc = 0;
                    if (buf) pacc ^= 1u;                                 // both accumulators used once more: the parity to wait for flips
                    buf ^= 1u;
                }
            }
        }
        }
#ifdef CLUSTEN_TC_PROFILE
        if (lane == 0 && blockIdx.x < 148) {
            for (int x = 0; x < 5; ++x) tc_prof[blockIdx.x * 16 + x] = prof[x];
            tc_prof[blockIdx.x * 16 + 5] = clock64() - prof_t0;
        }
#endif
    } else if (warp < 6) {                                               // ---- split x -> (xh, xl), row per thread, into TMEM ----
        const int q = warp & 3, row = q * 32 + lane;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        uint32_t s = 0, ph = 0, sa = 0, pa = 0;
        const bool ln = a.ln_mean != nullptr, ln_affine = a.ln_gamma != nullptr;
        for (int o = blockIdx.x; o < outer_n; o += gridDim.x) {
            const int64_t r = (int64_t)(res ? o : o / a.tiles_n) * BM + row;
            const float mu = ln && r < a.R ? __ldg(a.ln_mean + r) : 0.f, rs = ln && r < a.R ? __ldg(a.ln_rstd + r) : 0.f;
            for (int kc = 0; kc < KC; ++kc) {
                mbar_wait(bar_xfull + 8 * s, ph);
                if (a.dbg & 4) {                                         // experiment: no LDS / split / tcgen05.st, hand-offs only
                    mbar_arrive(bar_xempty + 8 * s);
                    mbar_wait(bar_aempty + 8 * sa, pa ^ 1u);
                    mbar_arrive(bar_afull + 8 * sa);
                    if (++s == XS) { s = 0; ph ^= 1u; }
                    if (++sa == NA) { sa = 0; pa ^= 1u; }
                    continue;
                }
                const uint8_t *xs = gbase + s * A_BYTES;
                float4 x[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) x[c] = *reinterpret_cast<const float4 *>(xs + swz(row, c));
                if (ln && !ln_affine) {                                  // gamma / beta folded into the weights and the bias by the caller
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        x[c].x = (x[c].x - mu) * rs; x[c].y = (x[c].y - mu) * rs; x[c].z = (x[c].z - mu) * rs; x[c].w = (x[c].w - mu) * rs;
                    }
                } else if (ln) {                                         // y = (x - mean) * rstd * gamma + beta, as ln_fwd_kernel rounds it
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const float4 g = __ldg(reinterpret_cast<const float4 *>(a.ln_gamma + kc * BK + 4 * c));
                        const float4 b = __ldg(reinterpret_cast<const float4 *>(a.ln_beta + kc * BK + 4 * c));
                        x[c].x = fmaf((x[c].x - mu) * rs, g.x, b.x); x[c].y = fmaf((x[c].y - mu) * rs, g.y, b.y);
                        x[c].z = fmaf((x[c].z - mu) * rs, g.z, b.z); x[c].w = fmaf((x[c].w - mu) * rs, g.w, b.w);
                    }
                }
                uint32_t h[H ? 16 : 32], l[H ? 16 : 32];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float xv[4] = {x[c].x, x[c].y, x[c].z, x[c].w};
                    if (H) {                                             // fp16 hi / lo, two K elements per 32-bit TMEM column (even k low)
#pragma unroll
                        for (int e = 0; e < 4; e += 2) {
                            // saturating conversions (F2FP.SATFINITE, one instruction per pair): |x| beyond the fp16 range gives
                            // hi = +-65504 and a lo that saturates as well -- a finite result instead of inf.  (The explicit clamp
                            // this replaces was 64 of the ~180 instructions of a chunk, and these warps are what the MMAs wait for.)
                            const uint32_t hh = f16x2_sat(xv[e], xv[e + 1]);
                            const float2 hf = __half22float2(*reinterpret_cast<const __half2 *>(&hh));
                            h[2 * c + e / 2] = hh;
                            l[2 * c + e / 2] = f16x2_sat(xv[e] - hf.x, xv[e + 1] - hf.y);
                        }
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float hv = tf32_rn(xv[e]);
                            h[4 * c + e] = __float_as_uint(hv);
                            l[4 * c + e] = __float_as_uint(tf32_rn(xv[e] - hv));
                        }
                    }
                }
                // Generic-proxy reads, then async-proxy (TMA) writes to the same slot: the release needs a proxy fence.  Without it
                // the arrive (scheduled right behind the ISSUE of the eight LDS) ran ahead of their completion and the next box
                // landed under loads in flight -- one K chunk of a row wrong in ~0.1 % of the rows once the X ring was freed by
                // this arrive alone.
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive(bar_xempty + 8 * s);
                mbar_wait(bar_aempty + 8 * sa, pa ^ 1u);                 // the MMAs of NA chunks ago are done with this A stage
                __syncwarp();                                            // (.sync.aligned below: the lanes left the spin loops apart)
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t ta = tmem + lane_off + A_COL0 + sa * C::A_COLS;
                if constexpr (H) {
                    tmem_st16(ta, h);
                    tmem_st16(ta + 16, l);
                } else {
                    tmem_st32(ta, h);
                    tmem_st32(ta + 32, l);
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                mbar_arrive(bar_afull + 8 * sa);
                if (++s == XS) { s = 0; ph ^= 1u; }
                if (++sa == NA) { sa = 0; pa ^= 1u; }
            }
        }
    } else if (warp < 14) {                                              // ---- drain + epilogue ----
        // two warps per TMEM lane quadrant; each keeps half of the tile's columns (HC) of its 32 rows
        const int q = warp & 3, row = q * 32 + lane, et = threadIdx.x - 192, hf = (warp - 6) >> 2;     // et: 0..255
        constexpr int HC = BN / 2;
        const bool elected = et == 0;
        uint32_t ch = 0, nt = 0;                                         // chains, tiles of this CTA so far
        float acc[HC];
#ifdef CLUSTEN_TC_PROFILE
        long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const long long prof_t0 = clock64();
#endif
        if (EPI == EPI_RES && elected && (int)blockIdx.x < outer_n) {    // residual pieces of the first tile
            const int m0 = (res ? blockIdx.x : blockIdx.x / a.tiles_n) * BM, n0 = (res ? 0 : blockIdx.x % a.tiles_n) * BN;
#pragma unroll
            for (int j = 0; j < C::NSUB; ++j) {
                mbar_expect_tx(bar_cfull + 8 * j, SUB_BYTES);
                tma_box_2d(cbuf + j * SUB_BYTES, &mapRes, bar_cfull + 8 * j, n0 + 32 * j, m0);
            }
        }
        for (int o = blockIdx.x; o < outer_n; o += gridDim.x)
        for (int ci = 0; ci < inner_n; ++ci, ++nt) {
            const int m0 = (res ? o : o / a.tiles_n) * BM, n0 = (res ? ci : o % a.tiles_n) * BN;
            const uint32_t set = C::CSETS == 2 ? (nt & 1u) : 0u;
            if (EPI == EPI_RES && C::CSETS == 1 && nt > 0 && elected) {
                // one staging set: the residual pieces of this tile go in as soon as the previous tile's stores have read it, and
                // arrive while the MMAs of this tile run
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
#pragma unroll
                for (int j = 0; j < C::NSUB; ++j) {
                    mbar_expect_tx(bar_cfull + 8 * j, SUB_BYTES);
                    tma_box_2d(cbuf + j * SUB_BYTES, &mapRes, bar_cfull + 8 * j, n0 + 32 * j, m0);
                }
            }
            if (et < BN) {                                               // bias / gamma of this column block (the previous tile's
                const bool in = n0 + et < a.N;                           // readers all passed the barrier before its stores)
                s_bias[et] = a.bias && in ? __ldg(a.bias + n0 + et) : 0.f;
                if (EPI == EPI_RES) s_gamma[et] = a.gamma && in ? __ldg(a.gamma + n0 + et) : 1.f;
                if (H) s_wsc[et] = in ? __ldg(a.w_inv_scale + n0 + et) : 1.f;     // powers of two: undo the row scaling of the fp16 weights
            }
            for (int kc0 = 0; kc0 < KC; kc0 += chain, ++ch) {
                const uint32_t buf = ch & 1u;
                TC_CLK(e0);
                mbar_wait(bar_accf + 8 * buf, (ch >> 1) & 1u);
                TC_CLK(e1);
                TC_ADD(0, e0, e1);
                __syncwarp();
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + buf * C::ACC_COLS + hf * HC;
#pragma unroll
                for (int c0 = 0; c0 < HC; c0 += 16) {
                    uint32_t v[16];
                    tmem_ld16(taddr + c0, v);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int j = 0; j < 16; ++j) acc[c0 + j] = kc0 == 0 ? __uint_as_float(v[j]) : acc[c0 + j] + __uint_as_float(v[j]);
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                mbar_arrive(bar_acce + 8 * buf);
                TC_CLK(e2);
                TC_ADD(1, e1, e2);
            }
            TC_CLK(e3);
            // the staging set of this tile: the stores of its previous user must have read it (one set: the previous tile, two sets:
            // the tile before that, whose stores were committed one group earlier)
            if (EPI != EPI_RES && (a.dbg & 8)) continue;                 // experiment: drains only, no epilogue arithmetic / stores
            if (elected) {
                if (C::CSETS == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
            epi_bar();                                                   // bias in place; the staging set may be written
            TC_CLK(e4);
            TC_ADD(2, e3, e4);
            const uint32_t cset = cbuf + set * C::NSUB * SUB_BYTES;
#pragma unroll
            for (int i = 0; i < HC / 4; ++i) {                           // 16-byte cells of this thread: columns hf * HC + 4 i ..
                const int col = hf * HC + 4 * i, j = col >> 5, c = (col >> 2) & 7;
                if (EPI == EPI_RES && (i == 0 || c == 0))
                    mbar_wait(bar_cfull + 8 * (set * C::NSUB + j), (C::CSETS == 2 ? (nt >> 1) : nt) & 1u);
                float4 *cell = reinterpret_cast<float4 *>(gbase + C::RING + (set * C::NSUB + j) * SUB_BYTES + swz(row, c));
                const float4 bv = *reinterpret_cast<const float4 *>(s_bias + col);
                float o[4];
                if (H) {
                    const float4 wv = *reinterpret_cast<const float4 *>(s_wsc + col);
                    o[0] = fmaf(acc[4 * i], wv.x, bv.x); o[1] = fmaf(acc[4 * i + 1], wv.y, bv.y);
                    o[2] = fmaf(acc[4 * i + 2], wv.z, bv.z); o[3] = fmaf(acc[4 * i + 3], wv.w, bv.w);
                } else {
                    o[0] = acc[4 * i] + bv.x; o[1] = acc[4 * i + 1] + bv.y; o[2] = acc[4 * i + 2] + bv.z; o[3] = acc[4 * i + 3] + bv.w;
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    if (EPI == EPI_BIAS) {
                        if (n0 + col + e < a.alpha_cols) o[e] *= a.alpha;
                    } else if (EPI == EPI_GELU) {
                        o[e] = 0.5f * o[e] * (1.f + erff(o[e] * 0.70710678118654752440f));
                    }
                }
                if (EPI == EPI_RES) {
                    const float4 rv = *cell, gv = *reinterpret_cast<const float4 *>(s_gamma + col);
                    o[0] = rv.x + gv.x * o[0]; o[1] = rv.y + gv.y * o[1]; o[2] = rv.z + gv.z * o[2]; o[3] = rv.w + gv.w * o[3];
                }
                *cell = make_float4(o[0], o[1], o[2], o[3]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");         // generic-proxy writes -> visible to the TMA unit
            TC_CLK(e5);
            TC_ADD(3, e4, e5);
            epi_bar();
            TC_CLK(e6);
            TC_ADD(4, e5, e6);
            if (elected) {
#pragma unroll
                for (int j = 0; j < C::NSUB; ++j)
                    if (n0 + 32 * j < a.N) tma_store_2d(&mapY, cset + j * SUB_BYTES, n0 + 32 * j, m0);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                if (EPI == EPI_RES && C::CSETS == 2) {                   // residual pieces of the next tile into the other set, whose
                    // last stores (the previous tile's) must have read it
                    const bool same_row = ci + 1 < inner_n;              // the next tile of this CTA's walk
                    const int on = same_row ? o : o + (int)gridDim.x, in_ = same_row ? ci + 1 : 0;
                    if (on < outer_n) {
                        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                        const int m1 = (res ? on : on / a.tiles_n) * BM, n1 = (res ? in_ : on % a.tiles_n) * BN;
                        const uint32_t so = (set ^ 1u) * C::NSUB;
#pragma unroll
                        for (int j = 0; j < C::NSUB; ++j) {
                            mbar_expect_tx(bar_cfull + 8 * (so + j), SUB_BYTES);
                            tma_box_2d(cbuf + (so + j) * SUB_BYTES, &mapRes, bar_cfull + 8 * (so + j), n1 + 32 * j, m1);
                        }
                    }
                }
            }
        }
        if (elected) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all stores complete before the CTA exits
#ifdef CLUSTEN_TC_PROFILE
        if (elected && blockIdx.x < 148) {
            for (int x = 0; x < 5; ++x) tc_prof[blockIdx.x * 16 + 8 + x] = prof[x];
            tc_prof[blockIdx.x * 16 + 13] = clock64() - prof_t0;
        }
#endif
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

__global__ void tf32_split_kernel(const float *__restrict__ w, float *__restrict__ hi, float *__restrict__ lo, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const float x = w[i], h = tf32_rn(x);
        hi[i] = h;
        lo[i] = tf32_rn(x - h);
    }
}

// hi = fp16(w * s_n), lo = fp16(w * s_n - hi) with s_n = the power of two that brings the largest |w| of OUTPUT ROW n to [512, 1024):
// both halves stay in the normal fp16 range for every weight within 2^-13 of the row's largest (smaller ones lose bits that are below
// fp32 resolution of that output's sum); inv_scale[n] = 1 / s_n is applied to column n in the epilogue
__global__ void f16_split_kernel(const float *__restrict__ w, __half *__restrict__ hi, __half *__restrict__ lo, int64_t n, int K,
                                 const float *__restrict__ amax, float *__restrict__ inv_scale) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const int64_t row = i / K;
        const float sc = exp2f(floorf(log2f(1024.f / fmaxf(amax[row], 1e-30f))));
        if (i == row * K) inv_scale[row] = 1.f / sc;
        const float x = w[i] * sc;
        const __half h = __float2half_rn(x);
        hi[i] = h;
        lo[i] = __float2half_rn(x - __half2float(h));
    }
}

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

// matrix [rows, K] with row stride ld (elements), fp32 or fp16; box = [32 columns x box_rows] (one 128- or 64-byte swizzle row), rows
// past the end read as 0
static bool map_2d(CUtensorMap *m, const void *ptr, int64_t rows, int K, int64_t ld, int box_rows, bool half = false) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows}, strides[1] = {(cuuint64_t)ld * (half ? 2 : 4)};
    const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows}, estr[2] = {1u, 1u};
    return enc(m, half ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(ptr), dims, strides, box,
               estr, CU_TENSOR_MAP_INTERLEAVE_NONE, half ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int sm_count() {
    static int n = 0;
    if (!n) {
        if (const char *e = getenv("CLUSTEN_TC_GRID")) { n = atoi(e); if (n > 0) return n; n = 0; }   // experiment: fewer persistent CTAs
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    }
    return n;
}

template <int BN, int EPI, bool H>
static int launch(const CUtensorMap *m, const Args &a, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(linear_tc_kernel<BN, EPI, H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg<BN, H>::SMEM) != cudaSuccess)
            return set_error(CLUSTEN_EUNSUPPORTED, "linear_tc: cannot reserve %zu bytes of shared memory", Cfg<BN, H>::SMEM);
        attr_set = true;
    }
    const int tiles = a.tiles_m * a.tiles_n;
    linear_tc_kernel<BN, EPI, H><<<std::min(a.resident ? a.tiles_m : tiles, sm_count()), THREADS, Cfg<BN, H>::SMEM, st>>>(m[0], m[1], m[2], m[3], m[4], a);
    note_launches(1);
    return check_launch("linear_tc");
}

template <int BN, bool H>
static int launch_epi(int epi, const CUtensorMap *m, const Args &a, cudaStream_t st) {
    switch (epi) {
        case EPI_BIAS: return launch<BN, EPI_BIAS, H>(m, a, st);
        case EPI_GELU: return launch<BN, EPI_GELU, H>(m, a, st);
        case EPI_RES: return launch<BN, EPI_RES, H>(m, a, st);
    }
    return set_error(CLUSTEN_EINVAL, "linear_tc: unknown epilogue %d", epi);
}

template <bool H>
static int launch_bn(int BN, int epi, const CUtensorMap *m, const Args &a, cudaStream_t st) {
    switch (BN) {
        case 128: return launch_epi<128, H>(epi, m, a, st);
        case 96: return launch_epi<96, H>(epi, m, a, st);
        case 64: return launch_epi<64, H>(epi, m, a, st);
        default: return launch_epi<32, H>(epi, m, a, st);
    }
}

}  // namespace tc
}  // namespace clusten

using namespace clusten;

// hi = w rounded to TF32, lo = (w - hi) rounded to TF32: the two weight operands clusten_linear_tc_f32 reads.  n elements.
extern "C" int clusten_tf32_split(const float *w, float *hi, float *lo, int64_t n, void *stream) {
    if (n < 0 || (n > 0 && (!w || !hi || !lo))) return set_error(CLUSTEN_EINVAL, "tf32_split: bad arguments");
    if (n == 0) return 0;
    tc::tf32_split_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(w, hi, lo, n);
    note_launches(1);
    return check_launch("tf32_split");
}

// Y = epilogue(X . W^T + bias).  X [R,K] (row stride ldx), w_hi / w_lo [N,K] contiguous (clusten_tf32_split of W), bias [N] or NULL,
// Y [R,N] (row stride ldy); fp32.  epi 0: y = (acc + bias), columns < alpha_cols multiplied by alpha afterwards; epi 1: exact GELU of
// (acc + bias); epi 2: y = res + gamma * (acc + bias) (res [R,N] row stride ldres, gamma [N] or NULL = 1).  chain = K chunks of 32
// summed inside the tensor-core accumulator before it is added to the fp32 running sum (<= 0: the default, 4).
// ln_mean / ln_rstd [R] + ln_gamma / ln_beta [K] (or all NULL): the rows of X are LayerNorm-ed while they are split, with the
// statistics clusten_layer_norm_fwd(y = NULL) wrote -- the `self.norm1(x)` / `self.norm2(x)` / `self.norm(x)` in front of the layer.
// ln_gamma = ln_beta = NULL with statistics given: the rows are only normalised, (x - mean) * rstd -- for callers that folded the
// affine part into the layer (W' = W * gamma per input column, bias' = bias + W beta), which takes ~90 instructions and 16 loads per
// chunk off the split warps, the role the MMAs wait for.
// w_fp16 = 1: w_hi / w_lo are the fp16 operands of clusten_f16_split and w_inv_scale [N] its per-row factors; X is split into fp16 hi / lo
// as well (values beyond +-65504 saturate) and the products run as kind::f16 -- half the MMAs of the TF32 form at the same accuracy.
// Needs K % 32 == 0, N % 4 == 0, 16-byte aligned rows; anything else returns CLUSTEN_EUNSUPPORTED.
// fp16 form of clusten_tf32_split for a weight [N, K]: hi / lo fp16 of w[n,:] * s_n, s_n the power of two that brings amax[n] (device,
// max |w[n,:]|) to [512, 1024); writes 1 / s_n to inv_scale[n] (device, [N]) for the epilogue of clusten_linear_tc_f32(w_fp16 = 1).
extern "C" int clusten_f16_split(const float *w, void *hi, void *lo, int64_t N, int K, const float *amax, float *inv_scale, void *stream) {
    if (N < 0 || K <= 0 || !amax || !inv_scale || (N > 0 && (!w || !hi || !lo))) return set_error(CLUSTEN_EINVAL, "f16_split: bad arguments");
    if (N == 0) return 0;
    const int64_t n = N * K;
    tc::f16_split_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(w, (__half *)hi, (__half *)lo, n, K, amax, inv_scale);
    note_launches(1);
    return check_launch("f16_split");
}

extern "C" int clusten_linear_tc_f32(const float *x, const void *w_hi, const void *w_lo, const float *bias, const float *res,
                                     const float *gamma, float *y, int64_t R, int K, int N, int64_t ldx, int64_t ldy, int64_t ldres,
                                     int epi, float alpha, int alpha_cols, int chain, const float *ln_mean, const float *ln_rstd,
                                     const float *ln_gamma, const float *ln_beta, int w_fp16, const float *w_inv_scale, void *stream) {
    if (R < 0 || K <= 0 || N <= 0 || ldx < K || ldy < N || !x || !w_hi || !w_lo || !y || (epi == tc::EPI_RES && (!res || ldres < N)))
        return set_error(CLUSTEN_EINVAL, "linear_tc: bad arguments R=%lld K=%d N=%d", (long long)R, K, N);
    if ((ln_mean != nullptr) != (ln_rstd != nullptr) || (ln_gamma != nullptr) != (ln_beta != nullptr) || (ln_gamma && !ln_mean) ||
        (ln_gamma && (!aligned16(ln_gamma) || !aligned16(ln_beta))))
        return set_error(CLUSTEN_EINVAL, "linear_tc: LayerNorm needs mean, rstd [R] and, unless they are folded into the weights, 16-byte aligned gamma, beta [K]");
    if (R == 0) return 0;
    if (K % tc::BK || N % 4 || ldx % 4 || ldy % 4 || (epi == tc::EPI_RES && ldres % 4) || !aligned16(x) || !aligned16(w_hi) || !aligned16(w_lo) ||
        !aligned16(y) || (bias && !aligned16(bias)) || (res && !aligned16(res)) || (gamma && !aligned16(gamma)) || R > (1LL << 31) - tc::BM)
        return set_error(CLUSTEN_EUNSUPPORTED, "linear_tc: needs K %% 32 == 0, N %% 4 == 0 and 16-byte aligned rows (K=%d N=%d)", K, N);
    // widest tile that wastes the fewest columns (the X tile is re-read and re-split once per column tile)
    int BN = 32, waste = 1 << 30;
    for (int cand : {128, 96, 64, 32}) {
        const int w = (N + cand - 1) / cand * cand - N;
        if (w < waste) { waste = w; BN = cand; }
    }
    CUtensorMap m[5];
    const bool half = w_fp16 != 0;
    if (!tc::map_2d(&m[0], x, R, K, ldx, tc::BM) || !tc::map_2d(&m[1], w_hi, N, K, K, BN, half) || !tc::map_2d(&m[2], w_lo, N, K, K, BN, half) ||
        !tc::map_2d(&m[3], y, R, N, ldy, tc::BM) || !tc::map_2d(&m[4], epi == tc::EPI_RES ? res : y, R, N, epi == tc::EPI_RES ? ldres : ldy, tc::BM))
        return set_error(CLUSTEN_EUNSUPPORTED, "linear_tc: cuTensorMapEncodeTiled failed");
    tc::Args a;
    a.bias = bias; a.gamma = gamma;
    a.ln_mean = ln_mean; a.ln_rstd = ln_rstd; a.ln_gamma = ln_gamma; a.ln_beta = ln_beta;
    a.R = (int)R; a.K = K; a.N = N;
    a.tiles_m = (int)((R + tc::BM - 1) / tc::BM); a.tiles_n = (N + BN - 1) / BN;
    a.chain = chain > 0 ? chain : 4;
    a.alpha = alpha; a.alpha_cols = epi == tc::EPI_BIAS ? alpha_cols : 0;
    if (half && !w_inv_scale) return set_error(CLUSTEN_EINVAL, "linear_tc: the fp16 split needs w_inv_scale [N]");
    a.w_inv_scale = half ? w_inv_scale : nullptr;
    // resident rows: K fits the A stages of tensor memory (8 chunks fp16, 4 chunks TF32), there is more than one column tile to walk,
    // and the row tiles alone fill most of the SMs (else spreading row x column tiles over all SMs wins).  CLUSTEN_TC_RESIDENT=0: off
    {
        const char *e = getenv("CLUSTEN_TC_RESIDENT");
        const int kc = K / tc::BK, na = half ? 8 : 4;
        a.resident = !(e && e[0] == '0') && kc <= na && a.tiles_n >= 2 && a.tiles_m * 4 >= tc::sm_count() * 3;
        const char *d = getenv("CLUSTEN_TC_DBG");
        a.dbg = d ? atoi(d) : 0;
    }
    cudaStream_t st = (cudaStream_t)stream;
    return half ? tc::launch_bn<true>(BN, epi, m, a, st) : tc::launch_bn<false>(BN, epi, m, a, st);
}

#ifdef CLUSTEN_TC_PROFILE
// per CTA (148 x 16 counters): issuer [0] wait drained accumulator, [1] wait W, [2] wait A, [3] issue MMAs + commits, [4] chunks, [5] role
// cycles; first epilogue thread [8] wait accumulator, [9] drain, [10] wait staging + barrier, [11] arithmetic + STS, [12] barrier, [13] role cycles
extern "C" int clusten_linear_tc_profile(long long *host_out) {
    cudaDeviceSynchronize();
    return cudaMemcpyFromSymbol(host_out, tc::tc_prof, sizeof(long long) * 148 * 16) == cudaSuccess ? 0 : -1;
}
#endif
