// LayerNorm over the channel dimension of token rows, forward + backward (sm_100a).
//
// The AFF backbone normalises [B*n, C] token matrices with C = 32 ... 1024 before every attention / MLP / merge
// (mask2former/modeling/backbone/aff.py:196-199, 258, 617-620).  ATen's vectorised kernel spends one CTA per row, which at
// C = 32 and B*n = 262 144 rows is launch- and latency-bound (16.7 % of the AFF-Mini forward in
// profiles/r1_launches_aff_mini_fwd_b16_v3.md).  Here one WARP owns a row (C <= 1024): the row lives in registers, mean and
// variance are two shuffle reductions, HBM traffic is one read and one write.  Input / output may be fp32, fp16 or bf16
// independently; statistics and arithmetic are fp32.
//   fwd: y = (x - mean) * rstd * gamma + beta;   saves mean, rstd (fp32 [R]) when asked
//   bwd: dx = rstd * (g - mean_c(g) - xhat * mean_c(g * xhat)), g = dy * gamma;  dgamma += sum_r dy * xhat;  dbeta += sum_r dy
//        (dgamma / dbeta: per-CTA shared-memory partials, then one fp32 atomic per channel and CTA)
#include <algorithm>

#include "common.cuh"

namespace clusten {

constexpr int LN_MAX_PER_LANE = 32;      // C <= 1024

template <typename T> __device__ __forceinline__ float ln_ld(const T *p) { return to_f(*p); }

__device__ __forceinline__ float ln_warp_sum(float x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(FULL, x, o);
    return x;
}

// PER = elements per lane (C = 32 * PER exactly when FULLROW, else C <= 32 * PER with a tail predicate)
// Rows are walked grid-stride, TWO per warp and iteration (their loads are issued together, their shuffle reductions
// interleave): with one short-lived CTA per 8 rows the kernel sat at 2.0 TB/s at C = 64 (ncu, AFF-Tiny stage 0) -- CTA
// turnover, not bandwidth.
template <typename TI, typename TO, int PER>
__global__ void __launch_bounds__(256)
ln_fwd_kernel(const TI *__restrict__ x, const float *__restrict__ gamma, const float *__restrict__ beta, TO *__restrict__ y,
              float *__restrict__ mean_out, float *__restrict__ rstd_out, int64_t R, int C, float eps) {
    const int lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * 8;
    constexpr bool KEEP = PER <= 8;                     // gamma / beta live in registers for narrow rows only
    float gm[KEEP ? PER : 1], bt[KEEP ? PER : 1];
    if constexpr (KEEP) {
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int c = lane + 32 * k;
            gm[k] = c < C ? __ldg(gamma + c) : 0.f;
            bt[k] = c < C ? __ldg(beta + c) : 0.f;
        }
    }
    const float invC = 1.f / (float)C;
    for (int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); row < R; row += 2 * stride) {
        const int64_t row2 = row + stride;
        const bool two = row2 < R;
        const TI *xa = x + row * C, *xb = x + (two ? row2 : row) * C;
        float va[PER], vb[PER];
        float sa = 0.f, sb = 0.f;
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int c = lane + 32 * k;
            va[k] = c < C ? ln_ld(xa + c) : 0.f;
            vb[k] = c < C ? ln_ld(xb + c) : 0.f;
            sa += va[k];
            sb += vb[k];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { sa += __shfl_xor_sync(FULL, sa, o); sb += __shfl_xor_sync(FULL, sb, o); }
        const float ma = sa * invC, mb = sb * invC;
        float qa = 0.f, qb = 0.f;
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int c = lane + 32 * k;
            const float da = c < C ? va[k] - ma : 0.f, db = c < C ? vb[k] - mb : 0.f;
            qa = fmaf(da, da, qa);
            qb = fmaf(db, db, qb);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { qa += __shfl_xor_sync(FULL, qa, o); qb += __shfl_xor_sync(FULL, qb, o); }
        const float ra = rsqrtf(qa * invC + eps), rb = rsqrtf(qb * invC + eps);
        TO *ya = y + row * C, *yb = y + row2 * C;
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int c = lane + 32 * k;
            if (c < C && y) {                           // y == NULL: statistics only (the consumer normalises on the fly)
                const float g = KEEP ? gm[KEEP ? k : 0] : __ldg(gamma + c), be = KEEP ? bt[KEEP ? k : 0] : __ldg(beta + c);
                ya[c] = from_f<TO>(fmaf((va[k] - ma) * ra, g, be));
                if (two) yb[c] = from_f<TO>(fmaf((vb[k] - mb) * rb, g, be));
            }
        }
        if (lane == 0 && mean_out) {
            mean_out[row] = ma; rstd_out[row] = ra;
            if (two) { mean_out[row2] = mb; rstd_out[row2] = rb; }
        }
    }
}
template <typename TI, typename TG, typename TO, int PER>
__global__ void __launch_bounds__(256)
ln_bwd_kernel(const TG *__restrict__ dy, const TI *__restrict__ x, const float *__restrict__ gamma,
              const float *__restrict__ mean_in, const float *__restrict__ rstd_in, TO *__restrict__ dx,
              float *__restrict__ dgamma, float *__restrict__ dbeta, int64_t R, int C) {
    extern __shared__ float ln_s[];                      // [2][C] partial dgamma / dbeta of this CTA
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) ln_s[c] = 0.f;
    __syncthreads();
    float ag[PER], ab[PER], gm[PER];
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const int c = lane + 32 * k;
        ag[k] = ab[k] = 0.f;
        gm[k] = c < C ? __ldg(gamma + c) : 0.f;
    }
    // narrow rows (C <= 256): two rows per warp and iteration -- their loads are in flight together, their reductions
    // interleave (see ln_fwd_kernel); wide rows keep one row per iteration (registers)
    constexpr bool TWO = PER <= 8;
    const int64_t stride = (int64_t)gridDim.x * 8;
    const float invC = 1.f / (float)C;
    for (int64_t row = (int64_t)blockIdx.x * 8 + warp; row < R; row += (TWO ? 2 : 1) * stride) {
        const int64_t row2 = row + stride;
        const bool two = TWO && row2 < R;
        const int64_t rb = two ? row2 : row;
        const TI *xa = x + row * C, *xb = x + rb * C;
        const TG *ga = dy + row * C, *gb = dy + rb * C;
        const float mean_a = mean_in[row], rstd_a = rstd_in[row], mean_b = mean_in[rb], rstd_b = rstd_in[rb];
        float xha[PER], gva[PER], xhb[TWO ? PER : 1], gvb[TWO ? PER : 1];
        float s1a = 0.f, s2a = 0.f, s1b = 0.f, s2b = 0.f;
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int c = lane + 32 * k;
            const bool in = c < C;
            const float da = in ? to_f(ga[c]) : 0.f;
            xha[k] = in ? (ln_ld(xa + c) - mean_a) * rstd_a : 0.f;
            gva[k] = da * gm[k];
            s1a += gva[k];
            s2a = fmaf(gva[k], xha[k], s2a);
            ag[k] = fmaf(da, xha[k], ag[k]);
            ab[k] += da;
            if constexpr (TWO) {
                const float db = (in && two) ? to_f(gb[c]) : 0.f;
                xhb[k] = in ? (ln_ld(xb + c) - mean_b) * rstd_b : 0.f;
                gvb[k] = db * gm[k];
                s1b += gvb[k];
                s2b = fmaf(gvb[k], xhb[k], s2b);
                ag[k] = fmaf(db, xhb[k], ag[k]);
                ab[k] += db;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s1a += __shfl_xor_sync(FULL, s1a, o); s2a += __shfl_xor_sync(FULL, s2a, o);
            if constexpr (TWO) { s1b += __shfl_xor_sync(FULL, s1b, o); s2b += __shfl_xor_sync(FULL, s2b, o); }
        }
        s1a *= invC; s2a *= invC; s1b *= invC; s2b *= invC;
        TO *da_ = dx + row * C, *db_ = dx + row2 * C;
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int c = lane + 32 * k;
            if (c < C) {
                da_[c] = from_f<TO>(rstd_a * (gva[k] - s1a - xha[k] * s2a));
                if constexpr (TWO) { if (two) db_[c] = from_f<TO>(rstd_b * (gvb[k] - s1b - xhb[k] * s2b)); }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const int c = lane + 32 * k;
        if (c < C) { atomicAdd(ln_s + c, ag[k]); atomicAdd(ln_s + C + c, ab[k]); }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        atomicAdd(dgamma + c, ln_s[c]);
        atomicAdd(dbeta + c, ln_s[C + c]);
    }
}

template <typename TI, typename TO>
static int ln_fwd_launch(const TI *x, const float *g, const float *b, TO *y, float *mean, float *rstd, int64_t R, int C, float eps,
                         cudaStream_t st) {
    const int grid = (int)std::min<int64_t>((R + 7) / 8, 148 * 8);
    const int per = (C + 31) / 32;
#define LN_F(P_) ln_fwd_kernel<TI, TO, P_><<<grid, 256, 0, st>>>(x, g, b, y, mean, rstd, R, C, eps)
    if (per <= 1) LN_F(1); else if (per <= 2) LN_F(2); else if (per <= 4) LN_F(4); else if (per <= 8) LN_F(8);
    else if (per <= 12) LN_F(12); else if (per <= 16) LN_F(16); else if (per <= 24) LN_F(24); else LN_F(32);
#undef LN_F
    note_launches(1);
    return check_launch("layer_norm_fwd");
}

template <typename TI, typename TG, typename TO>
static int ln_bwd_launch(const TG *dy, const TI *x, const float *g, const float *mean, const float *rstd, TO *dx, float *dg, float *db,
                         int64_t R, int C, cudaStream_t st) {
    const int grid = (int)std::min<int64_t>((R + 7) / 8, 148 * 8);
    const int per = (C + 31) / 32;
    const size_t smem = (size_t)2 * C * sizeof(float);
#define LN_B(P_) ln_bwd_kernel<TI, TG, TO, P_><<<grid, 256, smem, st>>>(dy, x, g, mean, rstd, dx, dg, db, R, C)
    if (per <= 1) LN_B(1); else if (per <= 2) LN_B(2); else if (per <= 4) LN_B(4); else if (per <= 8) LN_B(8);
    else if (per <= 16) LN_B(16); else LN_B(32);
#undef LN_B
    note_launches(1);
    return check_launch("layer_norm_bwd");
}

}  // namespace clusten

using namespace clusten;

#define LN_DISPATCH2(d0, d1, ...)                                                                         \
    switch (d0) {                                                                                         \
        case CLUSTEN_F32: { using TA = float; LN_DISPATCH1(d1, __VA_ARGS__); break; }                     \
        case CLUSTEN_F16: { using TA = __half; LN_DISPATCH1(d1, __VA_ARGS__); break; }                    \
        case CLUSTEN_BF16: { using TA = __nv_bfloat16; LN_DISPATCH1(d1, __VA_ARGS__); break; }            \
        default: return set_error(CLUSTEN_EDTYPE, "unknown dtype %d", d0);                                \
    }
#define LN_DISPATCH1(d1, ...)                                                                             \
    switch (d1) {                                                                                         \
        case CLUSTEN_F32: { using TB = float; __VA_ARGS__; break; }                                       \
        case CLUSTEN_F16: { using TB = __half; __VA_ARGS__; break; }                                      \
        case CLUSTEN_BF16: { using TB = __nv_bfloat16; __VA_ARGS__; break; }                              \
        default: return set_error(CLUSTEN_EDTYPE, "unknown dtype %d", d1);                                \
    }

extern "C" int clusten_layer_norm_fwd(const void *x, const float *gamma, const float *beta, void *y, float *mean, float *rstd,
                                      int64_t R, int C, float eps, int x_dtype, int y_dtype, void *stream) {
    if (R < 0 || C <= 0 || C > 32 * LN_MAX_PER_LANE) return set_error(CLUSTEN_EUNSUPPORTED, "layer norm: C=%d outside 1..1024", C);
    if (!x || !gamma || !beta || (!y && !mean) || (mean == nullptr) != (rstd == nullptr)) return set_error(CLUSTEN_EINVAL, "null pointer");
    if (R == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    LN_DISPATCH2(x_dtype, y_dtype, return ln_fwd_launch<TA, TB>((const TA *)x, gamma, beta, (TB *)y, mean, rstd, R, C, eps, st));
    return 0;
}

// d_x has x's dtype, d_y its own (g_dtype); d_gamma / d_beta fp32 [C], accumulated INTO (caller zeroes them)
extern "C" int clusten_layer_norm_bwd(const void *d_y, const void *x, const float *gamma, const float *mean, const float *rstd,
                                      void *d_x, float *d_gamma, float *d_beta, int64_t R, int C, int x_dtype, int g_dtype,
                                      void *stream) {
    if (R < 0 || C <= 0 || C > 32 * LN_MAX_PER_LANE) return set_error(CLUSTEN_EUNSUPPORTED, "layer norm: C=%d outside 1..1024", C);
    if (!d_y || !x || !gamma || !mean || !rstd || !d_x || !d_gamma || !d_beta) return set_error(CLUSTEN_EINVAL, "null pointer");
    if (R == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    LN_DISPATCH2(x_dtype, g_dtype, return ln_bwd_launch<TA, TB, TA>((const TB *)d_y, (const TA *)x, gamma, mean, rstd, (TA *)d_x, d_gamma, d_beta, R, C, st));
    return 0;
}
