// Column sum of a row-major [R, C] matrix into fp32 [C] -- the bias gradient of the backbone's Linear layers
// (grad_bias = grad_output.sum(0); 184 of them per AFF-Tiny training step).  ATen's generic reduce_kernel spends 42 us per
// call on these tall-skinny bf16 sums (7.8 ms = 14 % of the graphed step, benchmarks/profile_step.py); here a CTA owns a
// slab of rows and a group of 32 16-byte channel chunks: coalesced 512-byte row segments, four rows in flight per thread,
// partial sums meet in shared memory, one fp32 atomic per (CTA, channel).
#include <algorithm>

#include "common.cuh"

namespace clusten {

template <typename T>
__global__ void __launch_bounds__(256)
col_sum_kernel(const T *__restrict__ x, float *__restrict__ out, int R, int C, int64_t ld, int rows_per_cta) {
    constexpr int VPT = Vec<T>::VPT;                     // channels per 16-byte chunk
    __shared__ float red[8][32][VPT + 1];
    const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;          // chunk lane, row lane
    const int c0 = (blockIdx.y * 32 + cl) * VPT;
    const bool act = c0 < C;
    float acc[VPT];
#pragma unroll
    for (int v = 0; v < VPT; ++v) acc[v] = 0.f;
    const int r0 = blockIdx.x * rows_per_cta, r1 = min(r0 + rows_per_cta, R);
    if (act) {
        const T *xp = x + c0;
        for (int r = r0 + rl; r < r1; r += 32) {                     // 4 rows (8 row lanes apart) in flight
            float f[4][VPT];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int rr = r + 8 * u;
                if (rr < r1) load16(xp + (int64_t)rr * ld, f[u]);
                else {
#pragma unroll
                    for (int v = 0; v < VPT; ++v) f[u][v] = 0.f;
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int v = 0; v < VPT; ++v) acc[v] += f[u][v];
        }
    }
#pragma unroll
    for (int v = 0; v < VPT; ++v) red[rl][cl][v] = acc[v];
    __syncthreads();
    for (int t = threadIdx.x; t < 32 * VPT; t += blockDim.x) {
        const int k = t / VPT, v = t - k * VPT;
        const int c = (blockIdx.y * 32 + k) * VPT + v;
        float s = 0.f;
#pragma unroll
        for (int l = 0; l < 8; ++l) s += red[l][k][v];
        if (c < C && s != 0.f) atomicAdd(out + c, s);
    }
}

}  // namespace clusten

using namespace clusten;

extern "C" int clusten_col_sum(const void *x, float *out, int64_t R, int C, int64_t ld, int dtype, void *stream) {
    if (R < 0 || C <= 0 || ld < C) return set_error(CLUSTEN_EINVAL, "bad sizes R=%lld C=%d ld=%lld", (long long)R, C, (long long)ld);
    if (!x || !out) return set_error(CLUSTEN_EINVAL, "null pointer");
    if (R == 0) return 0;
    const int vpt = dtype == CLUSTEN_F32 ? 4 : 8;
    if (R >= (1LL << 31) || C % vpt || ld % vpt || !aligned16(x))
        return set_error(CLUSTEN_EUNSUPPORTED, "col_sum: needs C and the row stride multiples of %d elements, a 16-byte aligned base, R < 2^31", vpt);
    cudaStream_t st = (cudaStream_t)stream;
    const int gy = ceil_div(C, 32 * vpt);
    // ~4 CTAs per SM overall, at least 64 rows per CTA
    int rows_per_cta = (int)std::max<int64_t>(64, (R * gy + 148 * 4 - 1) / (148 * 4));
    rows_per_cta = (rows_per_cta + 31) / 32 * 32;
    const dim3 grid(ceil_div(R, rows_per_cta), gy);
    CLUSTEN_DISPATCH_DTYPE(dtype, (col_sum_kernel<T><<<grid, 256, 0, st>>>((const T *)x, out, (int)R, C, ld, rows_per_cta)));
    note_launches(1);
    return check_launch("col_sum");
}
