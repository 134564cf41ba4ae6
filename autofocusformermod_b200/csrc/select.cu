// Selection primitives of the adaptive downsampling (ClusterMerging.forward, mask2former/modeling/backbone/aff.py:320-324).
//
//   clusten_topk_select : replaces final_prob.topk(k, sorted=False) (aff.py:320).  The reference leaves the order of the
//                         k picks (and the choice among tied scores) to ATen; the canonical rule here is "first k of a
//                         stable descending sort" = descending score, ties -> lower index.  Implemented
//                         as the stable radix sort of sort.cu on an order-reversing bit transform of the fp32 score.
//   clusten_mask_select : replaces reserve_mask.nonzero()[1].reshape(b, reserve_num) (aff.py:323), a dynamic-shape op
//                         that forces a device->host sync in the reference; the count is known a priori, so this is
//                         a fixed-size ordered compaction with no sync.
#include "common.cuh"

namespace clusten {

__global__ void topk_keys_kernel(const float *__restrict__ score, uint32_t *__restrict__ keys, int64_t total) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= total) return;
    uint32_t u = __float_as_uint(score[p]);
    if (u == 0x80000000u) u = 0u;                                   // -0.0 == +0.0 for ordering
    const uint32_t asc = (u & 0x80000000u) ? ~u : (u | 0x80000000u);   // monotone increasing in the float value
    keys[p] = ~asc;                                                 // ascending sort of ~asc == descending by score
}

__global__ void topk_emit_kernel(const uint32_t *__restrict__ order, int n, int k, int64_t *__restrict__ out, int64_t out_stride) {
    const int b = blockIdx.y;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < k) out[(int64_t)b * out_stride + p] = (int64_t)order[(int64_t)b * n + p];
}

__global__ void __launch_bounds__(1024)
mask_select_kernel(const float *__restrict__ mask, int n, int count, int64_t *__restrict__ out, int64_t out_stride) {
    __shared__ int warp_tot[32];
    __shared__ int carry_s;
    const int b = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float *mrow = mask + (int64_t)b * n;
    int64_t *orow = out + (int64_t)b * out_stride;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int p = base + threadIdx.x;
        const bool f = p < n && mrow[p] != 0.f;
        const unsigned bal = __ballot_sync(FULL, f);
        if (lane == 0) warp_tot[warp] = __popc(bal);
        __syncthreads();
        if (warp == 0) {
            int wv = warp_tot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(FULL, wv, o);
                if (lane >= o) wv += t;
            }
            warp_tot[lane] = wv;                                     // inclusive
        }
        __syncthreads();
        const int carry = carry_s;
        const int dst = carry + (warp ? warp_tot[warp - 1] : 0) + __popc(bal & ((1u << lane) - 1u));
        if (f && dst < count) orow[dst] = p;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + warp_tot[31];
        __syncthreads();
    }
    for (int x = carry_s + threadIdx.x; x < count; x += 1024) orow[x] = 0;
}

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }


// Scores of the adaptive downsampling, ClusterMerging.forward (aff.py:292-315), in ONE pass over the tokens:
//   grid_prob    = all(pos.long() % s == 0),  s = stride (stride == 2) or the token's own 2 ** (ceil(log2(d_nearest)) + 1)   :297-302
//   final_prob   = grid_prob + learned_prob * alpha                                                                  :307-310
//   reserve_mask = all(pos.long() % (2 stride) == 0);  final_prob += reserve_mask * (-100)                            :313-315
// The reference spends ~22 element-wise / reduction launches of 1-3 us per merge on it (55 of the ~300 launches of an AFF-Mini
// forward).  Every fp32 operation is rounded separately, in the reference's order (no FMA contraction): the scores feed a top-k.
__global__ void __launch_bounds__(256)
merge_scores_kernel(const float2 *__restrict__ pos, const float *__restrict__ min_dist, int dist_stride, const float *__restrict__ lp,
                    float alpha, int stride, int reserve_on, float *__restrict__ final_prob, float *__restrict__ reserve_mask, int64_t total) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const float2 p = pos[e];
    const long long px = (long long)p.x, py = (long long)p.y;           // .long(): towards zero
    long long s = stride;
    if (min_dist) {
        const float k = __fadd_rn(ceilf(log2f(min_dist[e * dist_stride + 1])), 1.f);
        const float a = k < -149.f ? 0.f : k > 127.f ? INFINITY : ldexpf(1.f, (int)k);       // 2 ** k, exact
        s = (long long)a;
    }
    // (positions are non-negative, so the remainder of ATen -- sign of the divisor -- is the C one; s = 0 cannot occur for distinct
    // integer positions, whose nearest distance is >= 1: it is scored as "not on the grid" instead of dividing by zero)
    const float grid = (s != 0 && px % s == 0 && py % s == 0) ? 1.f : 0.f;
    float f = grid;
    if (lp) f = __fadd_rn(f, __fmul_rn(lp[e], alpha));
    if (reserve_on) {
        const long long s2 = 2LL * stride;
        const float r = (px % s2 == 0 && py % s2 == 0) ? 1.f : 0.f;
        f = __fadd_rn(f, __fmul_rn(r, -100.f));
        reserve_mask[e] = r;
    }
    final_prob[e] = f;
}

}  // namespace clusten

using namespace clusten;

extern "C" size_t clusten_topk_workspace_bytes(int B, int n) {
    const size_t tot = (size_t)B * n;
    return 5 * align256(tot * 4) + radix_sort_workspace_bytes(B, n) + 256;
}

extern "C" int clusten_topk_select(const float *score, int B, int n, int k, int64_t *idx_out, int64_t out_stride,
                                   void *workspace, size_t workspace_bytes, void *stream) {
    if (B < 0 || n <= 0 || k < 0 || k > n) return set_error(CLUSTEN_EINVAL, "bad sizes B=%d n=%d k=%d", B, n, k);
    if (!score || !idx_out || !workspace) return set_error(CLUSTEN_EINVAL, "null pointer");
    if (workspace_bytes < clusten_topk_workspace_bytes(B, n))
        return set_error(CLUSTEN_EWORKSPACE, "workspace too small: %zu < %zu", workspace_bytes, clusten_topk_workspace_bytes(B, n));
    if (B == 0 || k == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t tot = (size_t)B * n;
    const size_t stride = align256(tot * 4);
    char *ws = reinterpret_cast<char *>(workspace);
    uint32_t *k0 = reinterpret_cast<uint32_t *>(ws);
    uint32_t *k1 = reinterpret_cast<uint32_t *>(ws + stride);
    uint32_t *k2 = reinterpret_cast<uint32_t *>(ws + 2 * stride);
    uint32_t *v1 = reinterpret_cast<uint32_t *>(ws + 3 * stride);
    uint32_t *v2 = reinterpret_cast<uint32_t *>(ws + 4 * stride);
    void *hist = ws + 5 * stride;
    topk_keys_kernel<<<ceil_div((int64_t)tot, 256), 256, 0, st>>>(score, k0, (int64_t)tot);
    if (int e = radix_sort_pairs(k0, nullptr, k1, v1, k2, v2, B, n, 32, hist, st)) return e;
    topk_emit_kernel<<<dim3(ceil_div(k, 256), B), 256, 0, st>>>(v2, n, k, idx_out, out_stride);
    note_launches(2);
    return check_launch("topk_select");
}

extern "C" int clusten_mask_select(const float *mask, int B, int n, int count, int64_t *idx_out, int64_t out_stride, void *stream) {
    if (B < 0 || n <= 0 || count < 0) return set_error(CLUSTEN_EINVAL, "bad sizes B=%d n=%d count=%d", B, n, count);
    if (!mask || !idx_out) return set_error(CLUSTEN_EINVAL, "null pointer");
    if (B == 0 || count == 0) return 0;
    mask_select_kernel<<<B, 1024, 0, (cudaStream_t)stream>>>(mask, n, count, idx_out, out_stride);
    note_launches(1);
    return check_launch("mask_select");
}

extern "C" int clusten_merge_scores(const float *pos, const float *min_dist, int dist_stride, const float *learned_prob, float alpha,
                                    int stride, int reserve_on, float *final_prob, float *reserve_mask, int B, int n, void *stream) {
    if (B < 0 || n < 0 || stride <= 0 || (min_dist && dist_stride < 2)) return set_error(CLUSTEN_EINVAL, "merge_scores: bad sizes B=%d n=%d stride=%d", B, n, stride);
    if (!pos || !final_prob || (reserve_on && !reserve_mask)) return set_error(CLUSTEN_EINVAL, "null pointer");
    if (reinterpret_cast<uintptr_t>(pos) & 7u) return set_error(CLUSTEN_EUNSUPPORTED, "positions must be 8-byte aligned");
    const int64_t total = (int64_t)B * n;
    if (total == 0) return 0;
    merge_scores_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float2 *>(pos), min_dist, dist_stride, learned_prob, alpha, stride, reserve_on, final_prob, reserve_mask, total);
    note_launches(1);
    return check_launch("merge_scores");
}
