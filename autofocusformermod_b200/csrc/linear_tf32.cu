// fp32 Linear layer  Y[R,N] = X[R,K] . W[N,K]^T + bias[N]  on the tensor cores with the 3xTF32 split (fp32-level accuracy:
// x = hi + lo, x.w ~ hi.hi + hi.lo + lo.hi, the same scheme as the fp32 tile kernels of clusten_tile.cu) -- the q / kv / proj /
// fc1 / fc2 layers of the block (backbone/aff.py:62-70,103-106) in fp32 inference, where cuBLAS answers these skinny shapes
// (K = 32..768, N = 32..2304, R = 4 096..2 097 152) with SIMT sgemm kernels plus a separate bias kernel: 63 % of the device
// time of the AFF-Mini forward (profiles/r1_launches_aff_mini_fwd_b16_v12.md).
//
// OPT-IN (CLUSTEN_TC_LINEAR=1 in the Python layer); first parity cases green on a B200, not yet timed -- see DESIGN.md section 7.
//
// CTA = 128 rows x 64 columns, 8 warps of 16 rows each; K in chunks of 32 through shared memory (cp.async, double buffered,
// rows padded to 36 floats: the canonical m16n8k8 fragment reads are bank-conflict free); per chunk and warp: four A fragments
// (split once) against 8 x 4 B fragments -> 96 mma, each chunk summed from zero and added to the running sum outside the
// tensor cores (their accumulation truncates).  Bias in the epilogue, 8-byte stores.
#include <algorithm>

#include "common.cuh"

namespace clusten {

constexpr int LT_BM = 128, LT_BN = 64, LT_BK = 32, LT_LD = LT_BK + 4;
constexpr int LT_STAGE = (LT_BM + LT_BN) * LT_LD;       // floats per stage

__device__ __forceinline__ void lt_split(float x, uint32_t &hi, uint32_t &lo) {
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
    const float r = x - __uint_as_float(hi);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}
__device__ __forceinline__ void lt_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void lt_cp16(float *smem, const float *gmem, bool pred) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    const int sz = pred ? 16 : 0;                        // size 0: the 16 bytes are zero-filled, nothing is read
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}

__global__ void __launch_bounds__(256)
linear_tf32x3_kernel(const float *__restrict__ X, const float *__restrict__ W, const float *__restrict__ bias, float *__restrict__ Y,
                     int R, int K, int N, int64_t ldx, int64_t ldy) {
    extern __shared__ __align__(16) float lt_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int64_t r0 = (int64_t)blockIdx.x * LT_BM;
    const int n0 = blockIdx.y * LT_BN;
    // stage kc: X rows r0..r0+127 and W rows n0..n0+63, columns 32 kc .. 32 kc + 31; 8 threads cover one 128-byte row segment
    auto stage = [&](int kc, int buf) {
        float *Xs = lt_smem + buf * LT_STAGE, *Ws = Xs + LT_BM * LT_LD;
        const int c4 = (threadIdx.x & 7) * 4, rr = threadIdx.x >> 3;            // 32 rows per pass
#pragma unroll
        for (int p = 0; p < LT_BM / 32; ++p) {
            const int row = rr + 32 * p;
            const bool ok = r0 + row < R;
            lt_cp16(Xs + row * LT_LD + c4, X + (ok ? (r0 + row) * ldx + (int64_t)kc * LT_BK + c4 : 0), ok);
        }
#pragma unroll
        for (int p = 0; p < LT_BN / 32; ++p) {
            const int row = rr + 32 * p;
            const bool ok = n0 + row < N;
            lt_cp16(Ws + row * LT_LD + c4, W + (ok ? (int64_t)(n0 + row) * K + (int64_t)kc * LT_BK + c4 : 0), ok);
        }
        asm volatile("cp.async.commit_group;\n" ::);
    };
    float acc[LT_BN / 8][4];
#pragma unroll
    for (int n = 0; n < LT_BN / 8; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;
    const int KC = K / LT_BK;
    stage(0, 0);
    for (int kc = 0; kc < KC; ++kc) {
        if (kc + 1 < KC) {
            stage(kc + 1, (kc + 1) & 1);
            asm volatile("cp.async.wait_group 1;\n" ::);
        } else {
            asm volatile("cp.async.wait_group 0;\n" ::);
        }
        __syncthreads();
        const float *Xs = lt_smem + (kc & 1) * LT_STAGE, *Ws = Xs + LT_BM * LT_LD;
        const float *xa = Xs + (warp * 16 + g) * LT_LD + t, *xb = xa + 8 * LT_LD;
        const float *wr = Ws + g * LT_LD + t;
        uint32_t ah[LT_BK / 8][4], al[LT_BK / 8][4];
#pragma unroll
        for (int k8 = 0; k8 < LT_BK / 8; ++k8) {
            lt_split(xa[8 * k8], ah[k8][0], al[k8][0]);          // (row g,   k = t)
            lt_split(xb[8 * k8], ah[k8][1], al[k8][1]);          // (row g+8, k = t)
            lt_split(xa[8 * k8 + 4], ah[k8][2], al[k8][2]);      // (row g,   k = t + 4)
            lt_split(xb[8 * k8 + 4], ah[k8][3], al[k8][3]);      // (row g+8, k = t + 4)
        }
        // The tensor cores accumulate with truncation, so a long chain inside the mma accumulator drifts: the first cut, which kept
        // one chain over all of K, passed K <= 128 and left the 2e-6 band at K = 768 on the B200.  Each K chunk is therefore summed
        // from zero (12 mma) and added to the running sum with a rounded FADD (not re-run on hardware yet; a CPU emulation with a
        // truncating accumulator gives 6.7e-6 for the single chain and 4.1e-7 with this flush, tools/emulate_tf32_chain.py).
#pragma unroll
        for (int n = 0; n < LT_BN / 8; ++n) {
            float part[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int k8 = 0; k8 < LT_BK / 8; ++k8) {
                uint32_t b0h, b0l, b1h, b1l;
                lt_split(wr[n * 8 * LT_LD + 8 * k8], b0h, b0l);          // (k = t,     n = g): W[n0 + 8n + g][k]
                lt_split(wr[n * 8 * LT_LD + 8 * k8 + 4], b1h, b1l);      // (k = t + 4, n = g)
                lt_mma(part, al[k8], b0h, b1h);
                lt_mma(part, ah[k8], b0l, b1l);
                lt_mma(part, ah[k8], b0h, b1h);
            }
            acc[n][0] += part[0]; acc[n][1] += part[1]; acc[n][2] += part[2]; acc[n][3] += part[3];
        }
        __syncthreads();                                 // the buffer is refilled two iterations later
    }
    // epilogue: c0, c1 = (row g, columns 2t, 2t+1), c2, c3 = (row g+8, ...) of each 8-column block
    const int64_t ra = r0 + warp * 16 + g, rb = ra + 8;
#pragma unroll
    for (int n = 0; n < LT_BN / 8; ++n) {
        const int col = n0 + 8 * n + 2 * t;
        if (col < N) {                                   // N % 2 == 0: the pair is inside or outside together
            const float b0 = bias ? __ldg(bias + col) : 0.f, b1 = bias ? __ldg(bias + col + 1) : 0.f;
            if (ra < R) *reinterpret_cast<float2 *>(Y + ra * ldy + col) = make_float2(acc[n][0] + b0, acc[n][1] + b1);
            if (rb < R) *reinterpret_cast<float2 *>(Y + rb * ldy + col) = make_float2(acc[n][2] + b0, acc[n][3] + b1);
        }
    }
}

}  // namespace clusten

using namespace clusten;

// X [R,K] (row stride ldx), W [N,K] contiguous, bias [N] or NULL, Y [R,N] (row stride ldy); fp32; K % 32 == 0, N % 2 == 0,
// 16-byte aligned X / W rows, 8-byte aligned Y rows.
extern "C" int clusten_linear_f32(const float *x, const float *weight, const float *bias, float *y, int64_t R, int K, int N,
                                  int64_t ldx, int64_t ldy, void *stream) {
    if (R < 0 || K <= 0 || N <= 0 || ldx < K || ldy < N)
        return set_error(CLUSTEN_EINVAL, "bad sizes R=%lld K=%d N=%d ldx=%lld ldy=%lld", (long long)R, K, N, (long long)ldx, (long long)ldy);
    if (!x || !weight || !y) return set_error(CLUSTEN_EINVAL, "null pointer");
    if (R == 0) return 0;
    if (K % LT_BK || N % 2 || ldx % 4 || ldy % 2 || !aligned16(x) || !aligned16(weight) || (reinterpret_cast<uintptr_t>(y) & 7u) ||
        R >= (1LL << 31) - LT_BM || (R + LT_BM - 1) / LT_BM > 2147483647LL || (N + LT_BN - 1) / LT_BN > 65535)
        return set_error(CLUSTEN_EUNSUPPORTED, "linear_f32: needs K %% 32 == 0, N %% 2 == 0, 16-byte aligned rows (K=%d N=%d)", K, N);
    const size_t smem = (size_t)2 * LT_STAGE * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(linear_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_set = true;
    }
    const dim3 grid((unsigned)((R + LT_BM - 1) / LT_BM), (unsigned)((N + LT_BN - 1) / LT_BN));
    linear_tf32x3_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(x, weight, bias, y, (int)R, K, N, ldx, ldy);
    note_launches(1);
    return check_launch("linear_f32");
}
