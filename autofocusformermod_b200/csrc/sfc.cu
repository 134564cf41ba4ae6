// Balanced space-filling-curve clustering for sm_100a -- replaces the ~25 small torch kernels + segmented sort of
// space_filling_cluster (mask2former/modeling/backbone/point_utils.py:135-287; default branch sf_type='',
// use_anchor=True, no_reorder=False) by four launches + one radix sort.
//
// Bit-exactness contract: the fp32 sort key of the reference,
//     key = fl(fl(rank * (max_batch(ratio) + 1)) + ratio),   ratio = d2(prev anchor) / (d2(next anchor) + 1e-5),
// is reproduced operation by operation with round-to-nearest intrinsics (no FMA contraction), the maximum is taken over
// the WHOLE tensor handed in (under data parallelism that is the per-GPU batch, exactly like the reference under DDP),
// and the sort is a stable LSD radix sort on the key bits, so ties resolve to the lower original index (the oracle's
// stable=True rule).  The boustrophedon anchor order (point_utils.py:203-212) is evaluated in closed form instead of
// through argsort/scatter tables: rank(x, y) = y*npw + (y even ? x : npw-1-x).
#include <cmath>

#include "common.cuh"

namespace clusten {

struct SfcGrid {
    int nph, npw;        // anchors per column / row                                   (point_utils.py:174-175)
    float plw, plh;      // patch_len_w, patch_len_h as stored in the fp32 tensor       (point_utils.py:215)
};

__device__ __forceinline__ float2 anchor_centre(int r, const SfcGrid g) {
    const int gy = r / g.npw, t = r - gy * g.npw;
    const int gx = (gy & 1) ? g.npw - 1 - t : t;
    // init_pos_means = ordered_grid * patch_len_hw + patch_len_hw / 2 - 0.5                             (:217)
    const float cx = __fsub_rn(__fadd_rn(__fmul_rn((float)gx, g.plw), __fmul_rn(g.plw, 0.5f)), 0.5f);
    const float cy = __fsub_rn(__fadd_rn(__fmul_rn((float)gy, g.plh), __fmul_rn(g.plh, 0.5f)), 0.5f);
    return make_float2(cx, cy);
}

__global__ void sfc_ratio_kernel(const float2 *__restrict__ pos, int64_t total, SfcGrid g,
                                 int *__restrict__ rank_out, float *__restrict__ ratio_out, unsigned *__restrict__ max_bits) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float ratio = 0.f;
    if (p < total) {
        const float2 q = pos[p];
        const int nump = g.nph * g.npw;
        int cx = (int)floorf(__fdiv_rn(q.x, g.plw)), cy = (int)floorf(__fdiv_rn(q.y, g.plh));      // :227
        cx = min(max(cx, 0), g.npw - 1);
        cy = min(max(cy, 0), g.nph - 1);
        const int r = cy * g.npw + ((cy & 1) ? g.npw - 1 - cx : cx);                               // :228-229
        const float2 c0 = anchor_centre(r, g);
        float2 pv, nx;
        if (r > 0) pv = anchor_centre(r - 1, g);
        else {                                                                                     // :222
            const float2 c1 = anchor_centre(1, g);
            pv = make_float2(__fsub_rn(c0.x, __fsub_rn(c1.x, c0.x)), __fsub_rn(c0.y, __fsub_rn(c1.y, c0.y)));
        }
        if (r < nump - 1) nx = anchor_centre(r + 1, g);
        else {                                                                                     // :225
            const float2 cm = anchor_centre(nump - 2, g);
            nx = make_float2(__fadd_rn(c0.x, __fsub_rn(c0.x, cm.x)), __fadd_rn(c0.y, __fsub_rn(c0.y, cm.y)));
        }
        const float ax = __fsub_rn(q.x, pv.x), ay = __fsub_rn(q.y, pv.y);
        const float bx = __fsub_rn(q.x, nx.x), by = __fsub_rn(q.y, nx.y);
        const float dprev = __fadd_rn(__fmul_rn(ax, ax), __fmul_rn(ay, ay));                       // :233
        const float dnext = __fadd_rn(__fmul_rn(bx, bx), __fmul_rn(by, by));                       // :234
        ratio = __fdiv_rn(dprev, __fadd_rn(dnext, 1e-5f));                                         // :235
        rank_out[p] = r;
        ratio_out[p] = ratio;
    }
    // ratio >= 0 so the unsigned bit pattern is order-preserving; one atomic per warp
    unsigned bits = __float_as_uint(ratio);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) bits = max(bits, __shfl_xor_sync(FULL, bits, o));
    if ((threadIdx.x & 31) == 0) atomicMax(max_bits, bits);
}

__global__ void sfc_key_kernel(const int *__restrict__ rank, const float *__restrict__ ratio,
                               const unsigned *__restrict__ max_bits, uint32_t *__restrict__ keys, int64_t total) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= total) return;
    const float mx1 = __fadd_rn(__uint_as_float(*max_bits), 1.0f);                                 // dist_ratio.max()+1
    keys[p] = __float_as_uint(__fadd_rn(__fmul_rn((float)rank[p], mx1), ratio[p]));                // :237
}

__global__ void sfc_gather_kernel(const float2 *__restrict__ pos, const uint32_t *__restrict__ order, int n,
                                  float2 *__restrict__ pos_sorted, int64_t *__restrict__ ranking) {
    const int b = blockIdx.y;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const uint32_t src = order[(int64_t)b * n + p];
    pos_sorted[(int64_t)b * n + p] = pos[(int64_t)b * n + src];                                    // :260
    ranking[(int64_t)b * n + p] = (int64_t)src;
}

__global__ void sfc_cluster_kernel(const float2 *__restrict__ pos, const uint32_t *__restrict__ order, int n, int m, int k,
                                   float2 *__restrict__ mean_pos, int64_t *__restrict__ member_idx,
                                   int64_t *__restrict__ cluster_mask) {
    const int b = blockIdx.y;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= k) return;
    float sx = 0.f, sy = 0.f;
    int cnt = 0;
    for (int t = 0; t < m; ++t) {
        const int p = c * m + t;
        const bool in = p < n;
        if (in) {
            const float2 q = pos[(int64_t)b * n + order[(int64_t)b * n + p]];
            sx = __fadd_rn(sx, q.x);
            sy = __fadd_rn(sy, q.y);
            ++cnt;
        }
        member_idx[((int64_t)b * k + c) * m + t] = in ? p : 0;                                     // :282-283
        if (cluster_mask) cluster_mask[((int64_t)b * k + c) * m + t] = in ? 1 : 0;                 // :268-270
    }
    mean_pos[(int64_t)b * k + c] = make_float2(__fdiv_rn(sx, (float)cnt), __fdiv_rn(sy, (float)cnt));  // :264 / :271
}

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace clusten

using namespace clusten;

extern "C" size_t clusten_sfc_workspace_bytes(int B, int n) {
    const size_t tot = (size_t)B * n;
    return 7 * align256(tot * 4) + 256 + radix_sort_workspace_bytes(B, n) + 256;
}

extern "C" int clusten_sfc_cluster(const float *pos, int B, int n, int m, int h, int w,
                                   float *pos_sorted, float *mean_pos, int64_t *member_idx, int64_t *cluster_mask,
                                   int64_t *pos_ranking, void *workspace, size_t workspace_bytes, void *stream) {
    if (B <= 0 || n <= 0 || m <= 0 || h <= 0 || w <= 0)
        return set_error(CLUSTEN_EINVAL, "bad sizes B=%d n=%d m=%d h=%d w=%d", B, n, m, h, w);
    if (!pos || !pos_sorted || !mean_pos || !member_idx || !pos_ranking || !workspace)
        return set_error(CLUSTEN_EINVAL, "null pointer");
    const int k = (n + m - 1) / m;                                          // point_utils.py:167
    if ((int64_t)k * m != n && !cluster_mask) return set_error(CLUSTEN_EINVAL, "cluster_mask required when k*m != n");
    if (workspace_bytes < clusten_sfc_workspace_bytes(B, n))
        return set_error(CLUSTEN_EWORKSPACE, "workspace too small: %zu < %zu", workspace_bytes, clusten_sfc_workspace_bytes(B, n));
    // host scalars in double, Python round() == round-half-even == rint in the default FP environment (:173-176)
    const double patch_len = std::pow((double)h * (double)w / (double)k, 0.5);   // Python float ** 0.5 is C pow()
    SfcGrid g;
    g.nph = (int)std::rint((double)h / patch_len);
    g.npw = (int)std::rint((double)w / patch_len);
    if (g.nph < 1 || g.npw < 1 || g.nph * g.npw < 3 || g.npw > w)
        return set_error(CLUSTEN_EUNSUPPORTED, "degenerate anchor grid %dx%d for n=%d m=%d h=%d w=%d", g.nph, g.npw, n, m, h, w);
    g.plh = (float)((double)h / g.nph);
    g.plw = (float)((double)w / g.npw);

    cudaStream_t st = (cudaStream_t)stream;
    const size_t tot = (size_t)B * n;
    const size_t stride = align256(tot * 4);
    char *ws = reinterpret_cast<char *>(workspace);
    int *rank = reinterpret_cast<int *>(ws);
    float *ratio = reinterpret_cast<float *>(ws + stride);
    uint32_t *k0 = reinterpret_cast<uint32_t *>(ws + 2 * stride);
    uint32_t *k1 = reinterpret_cast<uint32_t *>(ws + 3 * stride);
    uint32_t *k2 = reinterpret_cast<uint32_t *>(ws + 4 * stride);
    uint32_t *v1 = reinterpret_cast<uint32_t *>(ws + 5 * stride);
    uint32_t *v2 = reinterpret_cast<uint32_t *>(ws + 6 * stride);
    unsigned *max_bits = reinterpret_cast<unsigned *>(ws + 7 * stride);
    void *hist = ws + 7 * stride + 256;

    cudaMemsetAsync(max_bits, 0, sizeof(unsigned), st);
    const int blocks = ceil_div((int64_t)tot, 256);
    const float2 *pos2 = reinterpret_cast<const float2 *>(pos);
    sfc_ratio_kernel<<<blocks, 256, 0, st>>>(pos2, (int64_t)tot, g, rank, ratio, max_bits);
    sfc_key_kernel<<<blocks, 256, 0, st>>>(rank, ratio, max_bits, k0, (int64_t)tot);
    if (int e = radix_sort_pairs(k0, nullptr, k1, v1, k2, v2, B, n, 32, hist, st)) return e;   // stable: ties -> lower index (:238)
    sfc_gather_kernel<<<dim3(ceil_div(n, 256), B), 256, 0, st>>>(pos2, v2, n, reinterpret_cast<float2 *>(pos_sorted), pos_ranking);
    sfc_cluster_kernel<<<dim3(ceil_div(k, 128), B), 128, 0, st>>>(pos2, v2, n, m, k, reinterpret_cast<float2 *>(mean_pos),
                                                                   member_idx, ((int64_t)k * m != n) ? cluster_mask : nullptr);
    note_launches(4);
    return check_launch("sfc_cluster");
}
