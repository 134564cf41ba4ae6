// Linear(F -> H) over the rows of the relative-position feature table, F <= 8 (the reference's pos_embed = Linear(5, heads),
// backbone/aff.py:101,129), restricted to the rows a stage references (device-side count, see clusten_stage_prepare).
//
// cuBLAS answers a K = 5 GEMM over the 65 025-row upper bound of the table with an unaligned legacy kernel (14.7 us per
// block of the AFF-Tiny step, plus the casts around it under autocast and a split-K GEMM + reduction for the weight gradient
// in the backward).  Here: one thread per (row, head) forward; backward = per-CTA partial sums of g^T [feat | 1] over the
// referenced rows only, folded in shared memory, one fp32 atomic per CTA and weight.  fp32 throughout.
#include <algorithm>

#include "common.cuh"

namespace clusten {

constexpr int TL_FMAX = 8, TL_HMAX = 32;

__global__ void __launch_bounds__(256)
table_linear_fwd_kernel(const float *__restrict__ feat, const float *__restrict__ W, const float *__restrict__ bias,
                        float *__restrict__ out, int R, int F, int H, const int *__restrict__ count) {
    const int U = count ? min(*count, R) : R;
    const int total = R * H;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < total; i += gridDim.x * 256) {
        const int r = i / H, h = i - r * H;
        float acc = 0.f;                                  // rows past the count are written as zeros (never gathered)
        if (r < U) {
            for (int j = 0; j < F; ++j) acc = fmaf(__ldg(feat + r * F + j), __ldg(W + h * F + j), acc);
            if (bias) acc += __ldg(bias + h);
        }
        out[i] = acc;
    }
}

// HP = H rounded up to a power of two (<= 32): thread = (row lane, head)
__global__ void __launch_bounds__(256)
table_linear_bwd_kernel(const float *__restrict__ g, const float *__restrict__ feat, float *__restrict__ dW, float *__restrict__ db,
                        int R, int F, int H, int HP, const int *__restrict__ count) {
    __shared__ float red[256][TL_FMAX + 1];
    const int U = count ? min(*count, R) : R;
    const int h = threadIdx.x & (HP - 1), rl = threadIdx.x / HP, RL = 256 / HP;
    float acc[TL_FMAX + 1];
#pragma unroll
    for (int k = 0; k <= TL_FMAX; ++k) acc[k] = 0.f;
    if (h < H) {
        for (int r = blockIdx.x * RL + rl; r < U; r += gridDim.x * RL) {
            const float gv = g[r * H + h];
#pragma unroll
            for (int j = 0; j < TL_FMAX; ++j)
                if (j < F) acc[j] = fmaf(gv, __ldg(feat + r * F + j), acc[j]);
            acc[TL_FMAX] += gv;
        }
    }
#pragma unroll
    for (int k = 0; k <= TL_FMAX; ++k) red[threadIdx.x][k] = acc[k];
    __syncthreads();
    for (int t = threadIdx.x; t < HP * (TL_FMAX + 1); t += 256) {
        const int h2 = t & (HP - 1), k = t / HP;
        if (h2 >= H || (k >= F && k != TL_FMAX)) continue;
        float s = 0.f;
        for (int l = 0; l < RL; ++l) s += red[l * HP + h2][k];
        if (s == 0.f) continue;
        if (k < F) atomicAdd(dW + h2 * F + k, s);
        else if (db) atomicAdd(db + h2, s);
    }
}

static int tl_check(int R, int F, int H) {
    if (R < 0 || F <= 0 || H <= 0) return set_error(CLUSTEN_EINVAL, "bad sizes R=%d F=%d H=%d", R, F, H);
    if (F > TL_FMAX || H > TL_HMAX || (int64_t)R * H >= (1LL << 31))
        return set_error(CLUSTEN_EUNSUPPORTED, "table_linear: needs F <= %d, H <= %d (F=%d H=%d)", TL_FMAX, TL_HMAX, F, H);
    return 0;
}

}  // namespace clusten

using namespace clusten;

extern "C" int clusten_table_linear_fwd(const float *feat, const float *weight, const float *bias, float *out, int R, int F, int H,
                                        const int32_t *count, void *stream) {
    if (int rc = tl_check(R, F, H)) return rc;
    if (!feat || !weight || !out) return set_error(CLUSTEN_EINVAL, "null pointer");
    if (R == 0) return 0;
    const int grid = std::min(ceil_div((int64_t)R * H, 256), 148 * 4);
    table_linear_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(feat, weight, bias, out, R, F, H, count);
    note_launches(1);
    return check_launch("table_linear_fwd");
}

// d_weight [H,F] and d_bias [H] (or NULL) are accumulated INTO (caller zeroes them); rows >= *count are ignored
extern "C" int clusten_table_linear_bwd(const float *d_out, const float *feat, float *d_weight, float *d_bias, int R, int F, int H,
                                        const int32_t *count, void *stream) {
    if (int rc = tl_check(R, F, H)) return rc;
    if (!d_out || !feat || !d_weight) return set_error(CLUSTEN_EINVAL, "null pointer");
    if (R == 0) return 0;
    int HP = 1;
    while (HP < H) HP *= 2;
    const int RL = 256 / HP;
    const int grid = std::max(1, std::min(ceil_div(R, RL * 8), 128));
    table_linear_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_out, feat, d_weight, d_bias, R, F, H, HP, count);
    note_launches(1);
    return check_launch("table_linear_bwd");
}
