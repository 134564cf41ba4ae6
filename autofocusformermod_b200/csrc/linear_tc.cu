// fp32 Linear layer  Y[R,N] = epilogue( X[R,K] . W[N,K]^T + bias[N] )  on the 5th-generation tensor cores (tcgen05.mma kind::tf32,
// accumulators in tensor memory) with the 3xTF32 split that keeps fp32-level accuracy:
//     x = xh + xl,  w = wh + wl   (xh, wh = the value rounded to TF32, xl, wl = the rounded remainder)
//     x.w ~ xl.wh + xh.wl + xh.wh                                   (the dropped xl.wl term is ~2^-22 of the product)
// These are the q / kv / proj / fc1 / fc2 layers of the block (backbone/aff.py:62-70,103-106,181-189) and the merge Linear
// (aff.py:282) in fp32 inference, where cuBLAS answers the skinny shapes (K = 32..1536, N = 32..2304, R = 4 096..2 097 152) with
// SIMT sgemm kernels plus a separate bias kernel: 64 % of the device time of the AFF-Mini forward
// (profiles/r2_launches_default.md).
//
// One persistent CTA per SM, 10 warps, four roles (the canonical sm_100 pipeline; every hand-off is an mbarrier):
//   warp 0 (one lane)   TMA producer: per 32-wide K chunk one box of X [128 rows x 128 B] and the matching boxes of the pre-split
//                       weights Wh, Wl [BN rows x 128 B], SWIZZLE_128B, into a ring of STAGES buffers
//   warps 2-5           split the X box in place: xh back over x, xl into the second A buffer (element-wise, so the swizzle does not
//                       matter), fence.proxy.async, arrive
//   warp 1 (one lane)   12 tcgen05.mma (M = 128, N = BN, K = 8) per chunk -- small terms first -- into one of two TMEM accumulators;
//                       tcgen05.commit frees the ring slot and, at the end of a chain, hands the accumulator to the epilogue
//   warps 6-9           drain: tcgen05.ld of their 32 TMEM lanes, add to the running fp32 sum in registers (the tensor core
//                       accumulates with truncation, so a chain is cut after `chain` chunks -- DESIGN.md section 7), then
//                       bias / GELU / scaled residual and 16-byte stores of their row
// The ring and the accumulator pair run across tile boundaries, so the next tile's loads, split and MMAs overlap the epilogue.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace clusten {

namespace tc {

constexpr int BM = 128, BK = 32;                       // rows per tile, K elements per chunk (= one 128-byte swizzle row)
constexpr int THREADS = 320;
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
// round to TF32 (nearest, ties away -- cvt.rna) on the bit pattern: the low 13 mantissa bits end up zero
__device__ __forceinline__ float tf32_rn(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }

template <int BN> struct Cfg {
    static constexpr int B_BYTES = BN * BK * 4;
    static constexpr int STAGE = 2 * A_BYTES + 2 * B_BYTES;              // xh | xl | wh | wl
    static constexpr int STAGES = BN >= 96 ? 3 : 4;
    static constexpr int ACC_COLS = BN == 96 ? 128 : BN;                 // column pitch of the two accumulators
    static constexpr int TMEM_COLS = 2 * ACC_COLS;                       // a power of two >= 32
    static constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    static constexpr size_t SMEM = (size_t)STAGES * STAGE + 1024 /* alignment slack */ + 256 /* barriers */;
};

enum { EPI_BIAS = 0, EPI_GELU = 1, EPI_RES = 2 };

struct Args {
    const float *bias, *res, *gamma;
    float *y;
    int R, K, N;
    int64_t ldy, ldres;
    int tiles_m, tiles_n, chain;
    float alpha;
    int alpha_cols;
};

template <int BN, int EPI>
__global__ void __launch_bounds__(THREADS, 1)
linear_tc_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapWh,
                 const __grid_constant__ CUtensorMap mapWl, const Args a) {
    using C = Cfg<BN>;
    extern __shared__ uint8_t tc_smem_raw[];
    const uint32_t base = (smem_u32(tc_smem_raw) + 1023u) & ~1023u;
    uint8_t *gbase = tc_smem_raw + (base - smem_u32(tc_smem_raw));
    const uint32_t bars = base + C::STAGES * C::STAGE;                   // 8-byte barriers
    const uint32_t bar_full = bars, bar_conv = bars + 8 * C::STAGES, bar_empty = bars + 16 * C::STAGES;
    const uint32_t bar_accf = bars + 24 * C::STAGES, bar_acce = bar_accf + 16;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(gbase + C::STAGES * C::STAGE + 24 * C::STAGES + 32);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < C::STAGES; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_conv + 8 * s, 128);
            mbar_init(bar_empty + 8 * s, 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(bar_accf + 8 * b, 1);
            mbar_init(bar_acce + 8 * b, 128);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(C::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;

    const int KC = a.K / BK, tiles = a.tiles_m * a.tiles_n;
    const int chain = a.chain;

    if (warp == 0) {
        if (lane == 0) {                                                 // ---- TMA producer ----
            uint32_t it = 0;
            for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
                const int m0 = (t / a.tiles_n) * BM, n0 = (t % a.tiles_n) * BN;
                for (int kc = 0; kc < KC; ++kc, ++it) {
                    const uint32_t s = it % C::STAGES, ph = (it / C::STAGES) & 1u;
                    mbar_wait(bar_empty + 8 * s, ph ^ 1u);
                    const uint32_t st = base + s * C::STAGE, full = bar_full + 8 * s;
                    mbar_expect_tx(full, A_BYTES + 2 * C::B_BYTES);
                    tma_box_2d(st, &mapX, full, kc * BK, m0);
                    tma_box_2d(st + 2 * A_BYTES, &mapWh, full, kc * BK, n0);
                    tma_box_2d(st + 2 * A_BYTES + C::B_BYTES, &mapWl, full, kc * BK, n0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {                                                 // ---- MMA issuer ----
            uint32_t it = 0, ch = 0;
            for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
                for (int kc = 0; kc < KC; ++kc, ++it) {
                    const uint32_t s = it % C::STAGES, ph = (it / C::STAGES) & 1u;
                    const uint32_t buf = ch & 1u;
                    const bool first = kc % chain == 0, last = (kc + 1) % chain == 0 || kc + 1 == KC;
                    if (first) {
                        mbar_wait(bar_acce + 8 * buf, ((ch >> 1) & 1u) ^ 1u);     // the epilogue has drained this accumulator
                    }
                    mbar_wait(bar_conv + 8 * s, ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t st = base + s * C::STAGE;
                    const uint64_t xh = umma_desc(st), xl = umma_desc(st + A_BYTES);
                    const uint64_t wh = umma_desc(st + 2 * A_BYTES), wl = umma_desc(st + 2 * A_BYTES + C::B_BYTES);
                    const uint32_t d = tmem + buf * C::ACC_COLS;
#pragma unroll
                    for (int k = 0; k < BK / 8; ++k)                     // one K = 8 step is 32 bytes: +2 in the address field
                        umma_tf32(d, xl + 2 * k, wh + 2 * k, C::IDESC, (first && k == 0) ? 0u : 1u);
#pragma unroll
                    for (int k = 0; k < BK / 8; ++k) umma_tf32(d, xh + 2 * k, wl + 2 * k, C::IDESC, 1u);
#pragma unroll
                    for (int k = 0; k < BK / 8; ++k) umma_tf32(d, xh + 2 * k, wh + 2 * k, C::IDESC, 1u);
                    umma_commit(bar_empty + 8 * s);                      // the ring slot is free once these MMAs have read it
                    if (last) {
                        umma_commit(bar_accf + 8 * buf);
                        ++ch;
                    }
                }
            }
        }
    } else if (warp < 6) {                                               // ---- split x -> (xh, xl) in place ----
        const int c = threadIdx.x - 64;                                  // 0..127
        uint32_t it = 0;
        for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
            for (int kc = 0; kc < KC; ++kc, ++it) {
                const uint32_t s = it % C::STAGES, ph = (it / C::STAGES) & 1u;
                mbar_wait(bar_full + 8 * s, ph);
                float4 *xh = reinterpret_cast<float4 *>(gbase + s * C::STAGE), *xl = xh + A_BYTES / 16;
#pragma unroll
                for (int j = 0; j < A_BYTES / 16 / 128; ++j) {
                    const float4 v = xh[c + 128 * j];
                    float4 h, l;
                    h.x = tf32_rn(v.x); l.x = tf32_rn(v.x - h.x);
                    h.y = tf32_rn(v.y); l.y = tf32_rn(v.y - h.y);
                    h.z = tf32_rn(v.z); l.z = tf32_rn(v.z - h.z);
                    h.w = tf32_rn(v.w); l.w = tf32_rn(v.w - h.w);
                    xh[c + 128 * j] = h;
                    xl[c + 128 * j] = l;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the tensor core
                mbar_arrive(bar_conv + 8 * s);
            }
        }
    } else {                                                             // ---- drain + epilogue ----
        const int q = warp & 3, row = q * 32 + lane;                     // TMEM lane quadrant of this warp
        uint32_t ch = 0;
        float acc[BN];
        for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
            const int m0 = (t / a.tiles_n) * BM, n0 = (t % a.tiles_n) * BN;
            const int64_t r = (int64_t)m0 + row;
            const float *__restrict__ rr = EPI == EPI_RES ? a.res + r * a.ldres + n0 : nullptr;
            constexpr int G4 = BN >= 128 ? 2 : BN >= 96 ? 4 : 8;                               // float4 per thread and column group
            constexpr bool PRE = EPI == EPI_RES && BN <= 64;             // the first 32 residual columns are fetched ahead of the MMAs
            float4 pre[8];
            if (PRE && r < a.R) {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    pre[i] = n0 + 4 * i < a.N ? __ldg(reinterpret_cast<const float4 *>(rr + 4 * i)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            for (int kc0 = 0; kc0 < KC; kc0 += chain, ++ch) {
                const uint32_t buf = ch & 1u;
                mbar_wait(bar_accf + 8 * buf, (ch >> 1) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + buf * C::ACC_COLS;
#pragma unroll
                for (int c0 = 0; c0 < BN; c0 += 16) {
                    uint32_t v[16];
                    tmem_ld16(taddr + c0, v);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int j = 0; j < 16; ++j) acc[c0 + j] = kc0 == 0 ? __uint_as_float(v[j]) : acc[c0 + j] + __uint_as_float(v[j]);
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                mbar_arrive(bar_acce + 8 * buf);
            }
            if (r < a.R) {
                float *__restrict__ yr = a.y + r * a.ldy + n0;
#pragma unroll
                for (int g0 = 0; g0 < BN; g0 += 4 * G4) {                // a group of columns at a time: all loads first, then math + stores
                    float4 rv[G4], bv[G4], gv[G4];
                    if (EPI == EPI_RES && !(PRE && g0 == 0)) {
#pragma unroll
                        for (int i = 0; i < G4; ++i)
                            rv[i] = n0 + g0 + 4 * i < a.N ? __ldg(reinterpret_cast<const float4 *>(rr + g0 + 4 * i)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                    for (int i = 0; i < G4; ++i) {
                        const bool in = n0 + g0 + 4 * i < a.N;           // N % 4 == 0: the quad is inside or outside together
                        bv[i] = a.bias && in ? __ldg(reinterpret_cast<const float4 *>(a.bias + n0 + g0 + 4 * i)) : make_float4(0.f, 0.f, 0.f, 0.f);
                        if (EPI == EPI_RES)
                            gv[i] = a.gamma && in ? __ldg(reinterpret_cast<const float4 *>(a.gamma + n0 + g0 + 4 * i)) : make_float4(1.f, 1.f, 1.f, 1.f);
                    }
#pragma unroll
                    for (int i = 0; i < G4; ++i) {
                        const int c0 = g0 + 4 * i;
                        if (n0 + c0 < a.N) {
                            float o[4] = {acc[c0] + bv[i].x, acc[c0 + 1] + bv[i].y, acc[c0 + 2] + bv[i].z, acc[c0 + 3] + bv[i].w};
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                if (EPI == EPI_BIAS) {
                                    if (n0 + c0 + j < a.alpha_cols) o[j] *= a.alpha;
                                } else if (EPI == EPI_GELU) {
                                    o[j] = 0.5f * o[j] * (1.f + erff(o[j] * 0.70710678118654752440f));
                                }
                            }
                            if (EPI == EPI_RES) {
                                const float4 q = PRE && g0 == 0 ? pre[i] : rv[i];
                                o[0] = q.x + gv[i].x * o[0]; o[1] = q.y + gv[i].y * o[1];
                                o[2] = q.z + gv[i].z * o[2]; o[3] = q.w + gv[i].w * o[3];
                            }
                            *reinterpret_cast<float4 *>(yr + c0) = make_float4(o[0], o[1], o[2], o[3]);
                        }
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(C::TMEM_COLS) : "memory");
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

// fp32 matrix [rows, K] with row stride ld (elements); box = [32 columns x box_rows], 128-byte swizzle, rows past the end read as 0
static bool map_2d(CUtensorMap *m, const float *ptr, int64_t rows, int K, int64_t ld, int box_rows) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows}, strides[1] = {(cuuint64_t)ld * 4};
    const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows}, estr[2] = {1u, 1u};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int sm_count() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    }
    return n;
}

template <int BN, int EPI>
static int launch(const CUtensorMap &mx, const CUtensorMap &mh, const CUtensorMap &ml, const Args &a, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(linear_tc_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg<BN>::SMEM) != cudaSuccess)
            return set_error(CLUSTEN_EUNSUPPORTED, "linear_tc: cannot reserve %zu bytes of shared memory", Cfg<BN>::SMEM);
        attr_set = true;
    }
    const int tiles = a.tiles_m * a.tiles_n;
    linear_tc_kernel<BN, EPI><<<std::min(tiles, sm_count()), THREADS, Cfg<BN>::SMEM, st>>>(mx, mh, ml, a);
    note_launches(1);
    return check_launch("linear_tc");
}

template <int BN>
static int launch_epi(int epi, const CUtensorMap &mx, const CUtensorMap &mh, const CUtensorMap &ml, const Args &a, cudaStream_t st) {
    switch (epi) {
        case EPI_BIAS: return launch<BN, EPI_BIAS>(mx, mh, ml, a, st);
        case EPI_GELU: return launch<BN, EPI_GELU>(mx, mh, ml, a, st);
        case EPI_RES: return launch<BN, EPI_RES>(mx, mh, ml, a, st);
    }
    return set_error(CLUSTEN_EINVAL, "linear_tc: unknown epilogue %d", epi);
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
// Needs K % 32 == 0, N % 4 == 0, 16-byte aligned rows; anything else returns CLUSTEN_EUNSUPPORTED.
extern "C" int clusten_linear_tc_f32(const float *x, const float *w_hi, const float *w_lo, const float *bias, const float *res,
                                     const float *gamma, float *y, int64_t R, int K, int N, int64_t ldx, int64_t ldy, int64_t ldres,
                                     int epi, float alpha, int alpha_cols, int chain, void *stream) {
    if (R < 0 || K <= 0 || N <= 0 || ldx < K || ldy < N || !x || !w_hi || !w_lo || !y || (epi == tc::EPI_RES && (!res || ldres < N)))
        return set_error(CLUSTEN_EINVAL, "linear_tc: bad arguments R=%lld K=%d N=%d", (long long)R, K, N);
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
    CUtensorMap mx, mh, ml;
    if (!tc::map_2d(&mx, x, R, K, ldx, tc::BM) || !tc::map_2d(&mh, w_hi, N, K, K, BN) || !tc::map_2d(&ml, w_lo, N, K, K, BN))
        return set_error(CLUSTEN_EUNSUPPORTED, "linear_tc: cuTensorMapEncodeTiled failed");
    tc::Args a;
    a.bias = bias; a.res = res; a.gamma = gamma; a.y = y;
    a.R = (int)R; a.K = K; a.N = N; a.ldy = ldy; a.ldres = ldres;
    a.tiles_m = (int)((R + tc::BM - 1) / tc::BM); a.tiles_n = (N + BN - 1) / BN;
    a.chain = chain > 0 ? chain : 4;
    a.alpha = alpha; a.alpha_cols = epi == tc::EPI_BIAS ? alpha_cols : 0;
    cudaStream_t st = (cudaStream_t)stream;
    switch (BN) {
        case 128: return tc::launch_epi<128>(epi, mx, mh, ml, a, st);
        case 96: return tc::launch_epi<96>(epi, mx, mh, ml, a, st);
        case 64: return tc::launch_epi<64>(epi, mx, mh, ml, a, st);
        default: return tc::launch_epi<32>(epi, mx, mh, ml, a, st);
    }
}
