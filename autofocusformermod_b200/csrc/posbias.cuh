// Relative-position bias computed from token positions inside the attention kernels (opt-in: the clusten_attn_pos_* entry
// points; first parity cases green on a B200, not yet timed -- see DESIGN.md section 7).
//
// The reference gathers it from a table: bias[b,h,i,j] = pos_embed(pre_table)[pe_idx[b,i,j], h] with
//   rel    = clamp(pos[idx[b,i,j]] - (pos[i] - 511), 0, 1022) truncated to an integer        (backbone/aff.py:481-485)
//   row    = rel.y * 1023 + rel.x  ->  (dx, dy) = (rel.x - 511, rel.y - 511)                  (aff.py:21-24)
//   feat   = (dx, dy, dist, dy/dist, dx/dist), the 0/0 centre zeroed                          (aff.py:25-31)
//   bias_h = W[h,:] . feat + b[h]                                                             (aff.py:101,129)
// i.e. five multiply-adds per head and one rsqrt per (token, neighbour) from two positions the tile already addresses:
// no bias_idx operand, no table gathers, and the weight gradient becomes partial sums of dS * [feat | 1].
#pragma once
#include "common.cuh"

namespace clusten {

constexpr int PB_PARTS = 1024;           // partial-sum slots of the pos_embed gradient ([PB_PARTS][H][6] fp32, atomics spread over them)

struct PosBiasW { float w0, w1, w2, w3, w4, b; };

__device__ __forceinline__ PosBiasW pos_bias_load(const float *__restrict__ pe_w, const float *__restrict__ pe_b, int h) {
    PosBiasW w;
    w.w0 = __ldg(pe_w + 5 * h); w.w1 = __ldg(pe_w + 5 * h + 1); w.w2 = __ldg(pe_w + 5 * h + 2);
    w.w3 = __ldg(pe_w + 5 * h + 3); w.w4 = __ldg(pe_w + 5 * h + 4);
    w.b = pe_b ? __ldg(pe_b + h) : 0.f;
    return w;
}

struct RelFeat { float dx, dy, d, rd; };            // dist = d, 1/dist = rd (0 at the centre)

// (the three functions below also compile for the host: tests/test_posbias_host.py checks their arithmetic against the oracle's table)
__host__ __device__ __forceinline__ RelFeat rel_feat(float2 q, float2 k) {
    RelFeat f;
    f.dx = truncf(fminf(fmaxf(k.x - (q.x - 511.f), 0.f), 1022.f)) - 511.f;
    f.dy = truncf(fminf(fmaxf(k.y - (q.y - 511.f), 0.f), 1022.f)) - 511.f;
    const float s = fmaf(f.dx, f.dx, f.dy * f.dy);
#ifdef __CUDA_ARCH__
    f.rd = s > 0.f ? rsqrtf(s) : 0.f;
#else
    f.rd = s > 0.f ? 1.f / sqrtf(s) : 0.f;
#endif
    f.d = s * f.rd;
    return f;
}

__host__ __device__ __forceinline__ float pos_bias(const PosBiasW &w, float2 q, float2 k) {
    const RelFeat f = rel_feat(q, k);
    const float lin = fmaf(w.w0, f.dx, fmaf(w.w1, f.dy, w.b));
    const float ang = fmaf(w.w3, f.dy, w.w4 * f.dx);
    return fmaf(ang, f.rd, fmaf(w.w2, f.d, lin));
}

// acc += ds * [dx, dy, dist, dy/dist, dx/dist, 1]
__host__ __device__ __forceinline__ void pos_bias_grad(float (&acc)[6], float2 q, float2 k, float ds) {
    const RelFeat f = rel_feat(q, k);
    acc[0] = fmaf(ds, f.dx, acc[0]);
    acc[1] = fmaf(ds, f.dy, acc[1]);
    acc[2] = fmaf(ds, f.d, acc[2]);
    acc[3] = fmaf(ds, f.dy * f.rd, acc[3]);
    acc[4] = fmaf(ds, f.dx * f.rd, acc[4]);
    acc[5] += ds;
}

// warp-reduce the six partial sums and add them into slot `part` of the [PB_PARTS][H][6] buffer
__device__ __forceinline__ void pos_bias_grad_flush(float (&acc)[6], float *__restrict__ parts, int part, int H, int h, int lane) {
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        float v = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
        if (lane == k && v != 0.f) atomicAdd(parts + ((int64_t)(part & (PB_PARTS - 1)) * H + h) * 6 + k, v);
    }
}

}  // namespace clusten
