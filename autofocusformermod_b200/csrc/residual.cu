// Residual add with layer scale and stochastic depth in one pass -- the two `x = shortcut + drop_path(gamma * x)` lines of the
// reference block (backbone/aff.py:230,236; DropPath = timm 0.6.12, aff.py:10):
//
//     out[b, i, c] = res[b, i, c] + x[b, i, c] * gamma[c] * sample_scale[b]
//
// ATen runs this as up to three elementwise kernels, two of them through the slow broadcasting TensorIterator path (116
// launches, 2.1 ms of the 43.8 ms AFF-Tiny training step, profiles/r1_torch_profiler_tiny_train_graphed_v11.log), and the
// backward repeats them plus a reduction for d_gamma.  Here a CTA works inside ONE sample (grid.y = b), a thread owns four
// fixed channels: gamma and the sample's scale live in registers, rows stream through 8 / 16-byte accesses.
// Rounding follows the op-by-op formulation (product, product, sum -- no contraction), so fp32 results equal ATen's bit for bit.
#include <algorithm>

#include "common.cuh"

namespace clusten {

__device__ __forceinline__ void ld4(const float *p, float (&o)[4]) {
    const float4 v = *reinterpret_cast<const float4 *>(p);
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}
__device__ __forceinline__ void ld4(const __half *p, float (&o)[4]) {
    const uint2 v = *reinterpret_cast<const uint2 *>(p);
    const float2 a = __half22float2(*reinterpret_cast<const __half2 *>(&v.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2 *>(&v.y));
    o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
}
__device__ __forceinline__ void ld4(const __nv_bfloat16 *p, float (&o)[4]) {
    const uint2 v = *reinterpret_cast<const uint2 *>(p);
    o[0] = __uint_as_float(v.x << 16); o[1] = __uint_as_float(v.x & 0xffff0000u);
    o[2] = __uint_as_float(v.y << 16); o[3] = __uint_as_float(v.y & 0xffff0000u);
}
__device__ __forceinline__ void st4(float *p, const float (&o)[4]) {
    *reinterpret_cast<float4 *>(p) = make_float4(o[0], o[1], o[2], o[3]);
}
__device__ __forceinline__ void st4(__half *p, const float (&o)[4]) {
    uint2 v;
    *reinterpret_cast<__half2 *>(&v.x) = __floats2half2_rn(o[0], o[1]);
    *reinterpret_cast<__half2 *>(&v.y) = __floats2half2_rn(o[2], o[3]);
    *reinterpret_cast<uint2 *>(p) = v;
}
__device__ __forceinline__ void st4(__nv_bfloat16 *p, const float (&o)[4]) {
    uint2 v;
    *reinterpret_cast<__nv_bfloat162 *>(&v.x) = __floats2bfloat162_rn(o[0], o[1]);
    *reinterpret_cast<__nv_bfloat162 *>(&v.y) = __floats2bfloat162_rn(o[2], o[3]);
    *reinterpret_cast<uint2 *>(p) = v;
}

// thread layout shared by both kernels: tpr = C / 4 threads per row, rpb = 256 / tpr rows per CTA pass
struct ResGeom { int tpr, rpb; };

template <typename TR, typename TX, typename TO>
__global__ void __launch_bounds__(256)
scale_residual_fwd_kernel(const TR *__restrict__ res, const TX *__restrict__ x, const float *__restrict__ gamma,
                          const float *__restrict__ sample_scale, TO *__restrict__ out, int rows, int C, ResGeom g) {
    const int col = threadIdx.x % g.tpr, rl = threadIdx.x / g.tpr;
    if (rl >= g.rpb) return;
    float gm[4] = {1.f, 1.f, 1.f, 1.f};
    if (gamma) ld4(gamma + 4 * col, gm);
    const float s = sample_scale ? sample_scale[blockIdx.y] : 1.f;
    const int64_t base = (int64_t)blockIdx.y * rows * C + 4 * col;
    const int step = gridDim.x * g.rpb;
    for (int r = blockIdx.x * g.rpb + rl; r < rows; r += 2 * step) {       // two rows in flight
        const int r2 = r + step;
        const bool two = r2 < rows;
        const int64_t o1 = base + (int64_t)r * C, o2 = base + (int64_t)(two ? r2 : r) * C;
        float a1[4], v1[4], a2[4], v2[4];
        ld4(res + o1, a1); ld4(x + o1, v1);
        ld4(res + o2, a2); ld4(x + o2, v2);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            a1[k] = __fadd_rn(a1[k], __fmul_rn(__fmul_rn(v1[k], gm[k]), s));
            a2[k] = __fadd_rn(a2[k], __fmul_rn(__fmul_rn(v2[k], gm[k]), s));
        }
        st4(out + o1, a1);
        if (two) st4(out + o2, a2);
    }
}

// d_x = g * gamma * s;  d_gamma[c] += sum over rows of g * x * s  (fp32 atomics: one per CTA and channel)
template <typename TG, typename TX>
__global__ void __launch_bounds__(256)
scale_residual_bwd_kernel(const TG *__restrict__ gout, const TX *__restrict__ x, const float *__restrict__ gamma,
                          const float *__restrict__ sample_scale, TX *__restrict__ d_x, float *__restrict__ d_gamma,
                          int rows, int C, ResGeom g) {
    __shared__ float red[256][4];
    const int col = threadIdx.x % g.tpr, rl = threadIdx.x / g.tpr;
    const bool act = rl < g.rpb;
    float gm[4] = {1.f, 1.f, 1.f, 1.f}, acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (gamma && act) ld4(gamma + 4 * col, gm);
    const float s = sample_scale ? sample_scale[blockIdx.y] : 1.f;
    const int64_t base = (int64_t)blockIdx.y * rows * C + 4 * col;
    const int step = gridDim.x * g.rpb;
    if (act && !(d_gamma == nullptr && d_x == nullptr)) {
        for (int r = blockIdx.x * g.rpb + rl; r < rows; r += step) {
            const int64_t o = base + (int64_t)r * C;
            float gv[4], xv[4], dv[4];
            ld4(gout + o, gv);
            if (d_gamma) {
                ld4(x + o, xv);
#pragma unroll
                for (int k = 0; k < 4; ++k) acc[k] = fmaf(gv[k] * xv[k], s, acc[k]);
            }
            if (d_x) {
#pragma unroll
                for (int k = 0; k < 4; ++k) dv[k] = __fmul_rn(__fmul_rn(gv[k], s), gm[k]);      // autograd's order: (g * scale) * gamma
                st4(d_x + o, dv);
            }
        }
    }
    if (d_gamma == nullptr) return;
#pragma unroll
    for (int k = 0; k < 4; ++k) red[threadIdx.x][k] = acc[k];
    __syncthreads();
    if (threadIdx.x < g.tpr) {                              // row lane 0 of every channel group folds the other row lanes
        float t[4] = {0.f, 0.f, 0.f, 0.f};
        for (int l = 0; l < g.rpb; ++l)
#pragma unroll
            for (int k = 0; k < 4; ++k) t[k] += red[l * g.tpr + threadIdx.x][k];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (t[k] != 0.f) atomicAdd(d_gamma + 4 * threadIdx.x + k, t[k]);
    }
}

static int res_check(int64_t B, int64_t rows, int C, ResGeom *g) {
    if (B < 0 || rows < 0 || C <= 0) return set_error(CLUSTEN_EINVAL, "bad sizes B=%lld rows=%lld C=%d", (long long)B, (long long)rows, C);
    if (C % 4 || C > 1024 || B > 65535 || rows >= (1LL << 30))
        return set_error(CLUSTEN_EUNSUPPORTED, "scale_residual: needs C %% 4 == 0, C <= 1024, B <= 65535, rows < 2^30 (C=%d B=%lld)", C,
                         (long long)B);
    g->tpr = C / 4;
    g->rpb = 256 / g->tpr;
    return 0;
}

static int res_grid_x(int64_t B, int64_t rows, const ResGeom &g, int rows_in_flight) {
    const int64_t want = (rows + (int64_t)g.rpb * rows_in_flight - 1) / ((int64_t)g.rpb * rows_in_flight);
    const int64_t cap = std::max<int64_t>(1, (148 * 8 + B - 1) / B);          // ~8 CTAs per SM over the whole grid
    return (int)std::max<int64_t>(1, std::min(want, cap));
}

}  // namespace clusten

using namespace clusten;

#define RES_CASE(R_, X_, O_, TR_, TX_, TO_)                                                                                    \
    if (res_dtype == R_ && x_dtype == X_ && out_dtype == O_) {                                                                 \
        scale_residual_fwd_kernel<TR_, TX_, TO_><<<grid, 256, 0, st>>>((const TR_ *)res, (const TX_ *)x, gamma, sample_scale,  \
                                                                       (TO_ *)out, (int)rows, C, g);                           \
        launched = true;                                                                                                       \
    }

extern "C" int clusten_scale_residual_fwd(const void *res, const void *x, const float *gamma, const float *sample_scale, void *out,
                                          int64_t B, int64_t rows, int C, int res_dtype, int x_dtype, int out_dtype, void *stream) {
    ResGeom g;
    if (int rc = res_check(B, rows, C, &g)) return rc;
    if (!res || !x || !out) return set_error(CLUSTEN_EINVAL, "null pointer");
    if (B == 0 || rows == 0) return 0;
    if (!aligned16(res) || !aligned16(x) || !aligned16(out) || (gamma && !aligned16(gamma)))
        return set_error(CLUSTEN_EUNSUPPORTED, "scale_residual: 16-byte aligned bases required");
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 grid(res_grid_x(B, rows, g, 2), (unsigned)B);
    bool launched = false;
    RES_CASE(CLUSTEN_F32, CLUSTEN_F32, CLUSTEN_F32, float, float, float)
    RES_CASE(CLUSTEN_F32, CLUSTEN_F16, CLUSTEN_F32, float, __half, float)
    RES_CASE(CLUSTEN_F32, CLUSTEN_BF16, CLUSTEN_F32, float, __nv_bfloat16, float)
    RES_CASE(CLUSTEN_F16, CLUSTEN_F16, CLUSTEN_F16, __half, __half, __half)
    RES_CASE(CLUSTEN_BF16, CLUSTEN_BF16, CLUSTEN_BF16, __nv_bfloat16, __nv_bfloat16, __nv_bfloat16)
    RES_CASE(CLUSTEN_F16, CLUSTEN_F16, CLUSTEN_F32, __half, __half, float)
    RES_CASE(CLUSTEN_BF16, CLUSTEN_BF16, CLUSTEN_F32, __nv_bfloat16, __nv_bfloat16, float)
    if (!launched)
        return set_error(CLUSTEN_EDTYPE, "scale_residual: unsupported dtype combination res=%d x=%d out=%d", res_dtype, x_dtype, out_dtype);
    note_launches(1);
    return check_launch("scale_residual_fwd");
}
#undef RES_CASE

#define RES_CASE(G_, X_, TG_, TX_)                                                                                             \
    if (g_dtype == G_ && x_dtype == X_) {                                                                                      \
        scale_residual_bwd_kernel<TG_, TX_><<<grid, 256, 0, st>>>((const TG_ *)d_out, (const TX_ *)x, gamma, sample_scale,     \
                                                                  (TX_ *)d_x, d_gamma, (int)rows, C, g);                       \
        launched = true;                                                                                                       \
    }

// d_gamma is accumulated INTO (caller zeroes it); x may be NULL when d_gamma is NULL; d_x may be NULL
extern "C" int clusten_scale_residual_bwd(const void *d_out, const void *x, const float *gamma, const float *sample_scale, void *d_x,
                                          float *d_gamma, int64_t B, int64_t rows, int C, int g_dtype, int x_dtype, void *stream) {
    ResGeom g;
    if (int rc = res_check(B, rows, C, &g)) return rc;
    if (!d_out || (d_gamma && !x) || (!d_x && !d_gamma)) return set_error(CLUSTEN_EINVAL, "null pointer");
    if (B == 0 || rows == 0) return 0;
    if (!aligned16(d_out) || (x && !aligned16(x)) || (d_x && !aligned16(d_x)) || (gamma && !aligned16(gamma)) ||
        (d_gamma && !aligned16(d_gamma)))
        return set_error(CLUSTEN_EUNSUPPORTED, "scale_residual: 16-byte aligned bases required");
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 grid(res_grid_x(B, rows, g, 1), (unsigned)B);
    bool launched = false;
    RES_CASE(CLUSTEN_F32, CLUSTEN_F32, float, float)
    RES_CASE(CLUSTEN_F32, CLUSTEN_F16, float, __half)
    RES_CASE(CLUSTEN_F32, CLUSTEN_BF16, float, __nv_bfloat16)
    RES_CASE(CLUSTEN_F16, CLUSTEN_F16, __half, __half)
    RES_CASE(CLUSTEN_BF16, CLUSTEN_BF16, __nv_bfloat16, __nv_bfloat16)
    if (!launched) return set_error(CLUSTEN_EDTYPE, "scale_residual: unsupported dtype combination d_out=%d x=%d", g_dtype, x_dtype);
    note_launches(1);
    return check_launch("scale_residual_bwd");
}
#undef RES_CASE
