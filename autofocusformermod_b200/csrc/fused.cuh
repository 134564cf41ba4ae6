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

// ---- masked neighbours (rare path, shared by the two tile forward kernels) -------------------------------------------------------
// The reference does not remove a masked neighbour: it subtracts 100 from its logit (aff.py:137) and still reads key / value row
// idx[b,i,j] (= 0 for the padded tail of the last cluster, point_utils.py:283).  The mask-aware pack treats such entries as
// wildcards, so the tile phases leave their logit without the q.k term (S holds the bias only) and their value row unread.  The two
// passes below put both back.  They are ROW-PARALLEL (two lanes per token row, like the softmax) and evaluate the dot product / the
// value row once per run of equal key rows: the masked entries of a row nearly always share one key row (the padding points at row
// 0), so a padded cluster of 16 slots costs each of its tokens one dot and one axpy.  The tiles that touch the padded cluster are the
// LAST tiles of a sample -- a slow rare path there is the tail of the whole launch (measured: 2x on every AFF-Base call).

// S[row][j] += q[row] . k[idx[row][j]] - 100 (+ the position bias of that key row when posq_t != NULL) for every masked (row, j) of
// the tile.  Scalar arguments only (a kernel-parameter struct passed by reference to a non-inlined function lands on the stack and
// costs the hot path registers): qrow0 / krows = row 0 of the tile's queries / of the (b, h) key slice, idx_t / mask_t = the tile's
// [rows][M] blocks, posq_t = the tile's query positions, posk = the sample's key positions.
template <typename T>
__device__ __noinline__ void masked_logit_pass(const T *qrow0, int64_t q_sn, const T *krows, int64_t k_sn, const int64_t *idx_t,
                                               const uint8_t *mask_t, int M, int C, int Nk, int rows, uint32_t impm, float *S, int MP,
                                               int lane, const float2 *posq_t, const float2 *posk, const float *pe_w, const float *pe_b, int h) {
    const int row = lane >> 1, half = lane & 1, Mh = M >> 1, j0 = half * Mh;
    if (row < rows && !((impm >> row) & 1u)) {
        PosBiasW pw = {};
        if (posq_t) pw = pos_bias_load(pe_w, pe_b, h);
        const uint32_t *mk4 = reinterpret_cast<const uint32_t *>(mask_t + (int64_t)row * M + j0);      // M % 8 == 0: 4-byte aligned
        const int64_t *ir = idx_t + (int64_t)row * M + j0;
        const T *qr = qrow0 + (int64_t)row * q_sn;
        float *Sr = S + row * MP + j0;
        int64_t cached_k = -1;
        float cached_v = 0.f;
        for (int j = 0; j < Mh; j += 4) {
            const uint32_t m4 = mk4[j >> 2];
            if (m4 == 0x01010101u) continue;
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                if ((m4 >> (8 * x)) & 0xffu) continue;
                const int64_t kidx = min(max(ir[j + x], (int64_t)0), (int64_t)Nk - 1);
                if (kidx != cached_k) {
                    const T *kr = krows + kidx * k_sn;
                    float p = 0.f;
                    for (int c = 0; c < C; ++c) p = fmaf(to_f(qr[c]), to_f(kr[c]), p);
                    if (posq_t) p += pos_bias(pw, __ldg(posq_t + row), __ldg(posk + kidx));
                    cached_k = kidx;
                    cached_v = p - 100.f;
                }
                Sr[j + x] += cached_v;
            }
        }
    }
    __syncwarp();
}
template <typename T, bool PB>
__device__ __forceinline__ void masked_logit_pass(const FArgsOf<PB> &a, int b, int h, int i0, int rows, uint32_t impm, float *S, int MP, int lane) {
    const float2 *pq = nullptr, *pk = nullptr;
    const float *w = nullptr, *bb = nullptr;
    if constexpr (PB) {
        pq = reinterpret_cast<const float2 *>(a.pos_q) + (int64_t)b * a.Nq + i0;
        pk = reinterpret_cast<const float2 *>(a.pos_k) + (int64_t)b * a.Nk;
        w = a.pe_w; bb = a.pe_b;
    }
    masked_logit_pass<T>(reinterpret_cast<const T *>(a.q) + b * a.q_sb + h * a.q_sh + (int64_t)i0 * a.q_sn, a.q_sn,
                         reinterpret_cast<const T *>(a.k) + b * a.k_sb + h * a.k_sh, a.k_sn, a.idx + ((int64_t)b * a.Nq + i0) * a.M,
                         a.mask + ((int64_t)b * a.Nq + i0) * a.M, a.M, a.C, a.Nk, rows, impm, S, MP, lane, pq, pk, w, bb, h);
}

// Row-parallel scan of the tile's masked entries: e = S[row][j] leaves S (its octet column must not pull in the wildcard row) and is
// summed per row; returns through (key_row, e_sum) of the lane pair that owns the row: key_row = the one key row all masked entries
// of the row point at, -1 when the row has none, -2 when they point at several (the caller then takes masked_value_slow_row).
static __device__ __noinline__ void masked_value_scan(const int64_t *idx_t, const uint8_t *mask_t, int M, int Nk, int rows, uint32_t impm, float *S,
                                               int MP, int lane, int &key_row, float &e_sum) {
    const int row = lane >> 1, half = lane & 1, Mh = M >> 1, j0 = half * Mh;
    int k0 = -1;
    if (row < rows && !((impm >> row) & 1u)) {
        const uint32_t *mk4 = reinterpret_cast<const uint32_t *>(mask_t + (int64_t)row * M + j0);
        const int64_t *ir = idx_t + (int64_t)row * M + j0;
        for (int j = 0; j < Mh; j += 4) {
            const uint32_t m4 = mk4[j >> 2];
            if (m4 == 0x01010101u) continue;
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                if ((m4 >> (8 * x)) & 0xffu) continue;
                const int kidx = (int)min(max(ir[j + x], (int64_t)0), (int64_t)Nk - 1);
                k0 = k0 == -1 ? kidx : (k0 == kidx ? k0 : -2);
            }
        }
    }
    const int ko = __shfl_xor_sync(FULL, k0, 1);
    if (k0 == -1) k0 = ko;
    else if (ko != -1 && ko != k0) k0 = -2;
    float e = 0.f;
    if (k0 >= 0 && row < rows && !((impm >> row) & 1u)) {
        const uint32_t *mk4 = reinterpret_cast<const uint32_t *>(mask_t + (int64_t)row * M + j0);
        float *Sr = S + row * MP + j0;
        for (int j = 0; j < Mh; j += 4) {
            const uint32_t m4 = mk4[j >> 2];
            if (m4 == 0x01010101u) continue;
#pragma unroll
            for (int x = 0; x < 4; ++x)
                if (!((m4 >> (8 * x)) & 0xffu)) { e += Sr[j + x]; Sr[j + x] = 0.f; }
        }
    }
    e += __shfl_xor_sync(FULL, e, 1);
    key_row = k0;
    e_sum = e;
    __syncwarp();
}

// masked entries of the tile -> `flush(row, key_row, e_sum)`, called in warp-uniform control flow once per run of equal key rows;
// the caller adds e_sum * v[key_row] to its accumulators
template <typename T, bool PB, typename Flush>
__device__ __forceinline__ void masked_value_pass(const FArgsOf<PB> &a, int b, int i0, int rows, uint32_t impm, float *S, int MP, int lane, Flush flush) {
    const int M = a.M;
    const int64_t *idx_t = a.idx + ((int64_t)b * a.Nq + i0) * M;
    const uint8_t *mask_t = a.mask + ((int64_t)b * a.Nq + i0) * M;
    int key_row;
    float e_sum;
    masked_value_scan(idx_t, mask_t, M, a.Nk, rows, impm, S, MP, lane, key_row, e_sum);
    for (int r = 0; r < rows; ++r) {
        const int kk = __shfl_sync(FULL, key_row, 2 * r);
        const float ee = __shfl_sync(FULL, e_sum, 2 * r);
        if (kk >= 0) {
            if (ee != 0.f) flush(r, (int64_t)kk, ee);
        } else if (kk == -2) {                               // several key rows behind the row's masked entries (arbitrary masks): entry by entry
            const uint8_t *mk = mask_t + (int64_t)r * M;
            const int64_t *ir = idx_t + (int64_t)r * M;
            for (int x0 = 0; x0 < M; x0 += 32) {
                unsigned bal = __ballot_sync(FULL, x0 + lane < M && !mk[x0 + lane]);
                while (bal) {
                    const int j = x0 + __ffs(bal) - 1;
                    bal &= bal - 1;
                    const float e = S[r * MP + j];
                    __syncwarp();
                    if (lane == 0) S[r * MP + j] = 0.f;
                    if (e != 0.f) flush(r, min(max(ir[j], (int64_t)0), (int64_t)a.Nk - 1), e);
                }
            }
        }
    }
    __syncwarp();
}

// TMA-staged kernel (clusten_fused_tma.cu).  *taken = true when it took the call (its kernel is enqueued and exits at once when
// the pack routes the tensor to the generic kernels), false when the shape / layout is outside what it supports (the caller
// falls back to the per-warp tile kernel).  Returns the usual error code.
int fused_tma_launch(const FusedArgsPB &a, bool pos_bias, int dtype, const void *pack, cudaStream_t st, bool *taken);

}  // namespace clusten
