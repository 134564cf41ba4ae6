// Pieces shared by the fused attention forward kernels (clusten_fused.cu: one warp per (tile, head), operands straight from
// global memory; clusten_fused_tma.cu: one CTA per 64-token group, key / value octets staged once by TMA box loads).
#pragma once
#include "posbias.cuh"
#include "t2.cuh"

namespace clusten {

template <typename T> __device__ __forceinline__ float f_exp(float x) {
    if constexpr (sizeof(T) == 4) return expf(x); else return __expf(x);
}

struct FusedArgs {
    const void *q, *k, *v;
    const int64_t *idx;
    const float *bias_tab;
    const int32_t *bias_idx;
    const uint8_t *mask;
    const void *blank_k, *blank_v;
    void *out;
    float *probs, *lse;
    int B, H, Nq, Nk, C, M;
    int64_t q_sb, q_sh, q_sn, k_sb, k_sh, k_sn, v_sb, v_sh, v_sn, o_sb, o_sh, o_sn;
};
// position-bias variant (PB kernels, clusten_attn_pos_fwd): bias from positions instead of bias_tab / bias_idx (posbias.cuh).
// A separate type so that the kernels of the table variant keep their parameter block exactly as validated.
struct FusedArgsPB : FusedArgs {
    const float *pos_q, *pos_k, *pe_w, *pe_b;            // [B,Nq,2], [B,Nk,2], [H,5], [H] or NULL
};
template <bool PB> using FArgsOf = std::conditional_t<PB, FusedArgsPB, FusedArgs>;

// One (token, head) computed the slow way by one warp; `sm` = M + 2 floats of shared scratch.  Also THE generic kernel body.
template <typename T, bool PB = false>
__device__ __forceinline__ void fused_row_generic(const FArgsOf<PB> &a, int b, int h, int i, float *sm, int lane) {
    const T *q = reinterpret_cast<const T *>(a.q) + b * a.q_sb + h * a.q_sh + (int64_t)i * a.q_sn;
    const T *kb = reinterpret_cast<const T *>(a.k) + b * a.k_sb + h * a.k_sh;
    const T *vb = reinterpret_cast<const T *>(a.v) + b * a.v_sb + h * a.v_sh;
    const T *bk = reinterpret_cast<const T *>(a.blank_k) + h * a.C;
    const T *bv = reinterpret_cast<const T *>(a.blank_v) + h * a.C;
    const int64_t *irow = a.idx + ((int64_t)b * a.Nq + i) * a.M;
    const int32_t *bi = PB ? nullptr : a.bias_idx + ((int64_t)b * a.Nq + i) * a.M;
    const uint8_t *mk = a.mask ? a.mask + ((int64_t)b * a.Nq + i) * a.M : nullptr;
    const int M = a.M, C = a.C;
    PosBiasW pw = {};
    float2 pq = make_float2(0.f, 0.f);
    const float2 *PK = nullptr;
    if constexpr (PB) {
        pw = pos_bias_load(a.pe_w, a.pe_b, h);
        pq = __ldg(reinterpret_cast<const float2 *>(a.pos_q) + (int64_t)b * a.Nq + i);
        PK = reinterpret_cast<const float2 *>(a.pos_k) + (int64_t)b * a.Nk;
    }
    float mx = -INFINITY;
    for (int j = lane; j <= M; j += 32) {
        float s = 0.f;
        if (j < M) {
            const T *kr = kb + irow[j] * a.k_sn;
            for (int ch = 0; ch < C; ++ch) s = fmaf(to_f(q[ch]), to_f(kr[ch]), s);
            if constexpr (PB) s += pos_bias(pw, pq, __ldg(PK + irow[j]));
            else s += a.bias_tab[(int64_t)bi[j] * a.H + h];
            if (mk && !mk[j]) s += -100.f;
        } else {
            for (int ch = 0; ch < C; ++ch) s = fmaf(to_f(q[ch]), to_f(bk[ch]), s);
        }
        sm[j] = s;
        mx = fmaxf(mx, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, o));
    float sum = 0.f;
    for (int j = lane; j <= M; j += 32) {
        const float e = f_exp<T>(sm[j] - mx);
        sm[j] = e;
        sum += e;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(FULL, sum, o);
    const float inv = 1.f / sum;
    if (a.lse && lane == 0) a.lse[((int64_t)b * a.H + h) * a.Nq + i] = mx + logf(sum);
    __syncwarp();
    T *orow = reinterpret_cast<T *>(a.out) + b * a.o_sb + h * a.o_sh + (int64_t)i * a.o_sn;
    for (int ch = lane; ch < C; ch += 32) {
        float acc = sm[M] * to_f(bv[ch]);
        for (int j = 0; j < M; ++j) acc = fmaf(sm[j], to_f(vb[irow[j] * a.v_sn + ch]), acc);
        orow[ch] = from_f<T>(acc * inv);
    }
    if (a.probs) {
        float *pr = a.probs + (((int64_t)b * a.H + h) * a.Nq + i) * (M + 1);
        for (int j = lane; j <= M; j += 32) pr[j] = sm[j] * inv;
    }
    __syncwarp();
}

// TMA-staged kernel (clusten_fused_tma.cu).  *taken = true when it took the call (its kernel is enqueued and exits at once when
// the pack routes the tensor to the generic kernels), false when the shape / layout is outside what it supports (the caller
// falls back to the per-warp tile kernel).  Returns the usual error code.
int fused_tma_launch(const FusedArgsPB &a, bool pos_bias, int dtype, const void *pack, cudaStream_t st, bool *taken);

}  // namespace clusten
