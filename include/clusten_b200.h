/*
 * clusten_b200.h -- C ABI of libclusten_b200.so: the B200 (sm_100a) implementation of
 * AutoFocusFormer's CLUSTEN neighbourhood-attention hot path.
 *
 * Every entry point replaces one piece of the reference's native / library interface
 * (Eiphodos/autofocusformerMod; paths relative to mask2former/modeling/):
 *
 *   clusten_qk_fwd / clusten_qk_bwd   <- clusten/src/clustenqk_cuda.cpp:25-45  (pybind forward/backward, :48-51)
 *   clusten_av_fwd / clusten_av_bwd   <- clusten/src/clustenav_cuda.cpp:25-45
 *   clusten_wf_fwd / clusten_wf_bwd   <- clusten/src/clustenwf_cuda.cpp:25-45
 *   clusten_wg_fwd / clusten_wg_bwd   <- clusten/src/weighted_gather_cuda.cpp:25-45
 *   clusten_msdetrpc_fwd / _bwd       <- clusten/src/msdetrpc_cuda.cpp:25-51
 *   clusten_csr_*                     <- replaces the fastAtomicAdd scatter of the reference backward kernels
 *                                        (clustenqk_cuda_kernel.cu:125, clustenav_cuda_kernel.cu:121,
 *                                         clustenwf_cuda_kernel.cu:129, weighted_gather_cuda_kernel.cu:115)
 *   clusten_knn                       <- backbone/point_utils.py:28-60 (knn_keops; pykeops argKmin / Kmin_argKmin)
 *   clusten_sfc_cluster               <- backbone/point_utils.py:135-287 (space_filling_cluster, default branch)
 *   clusten_topk_select, clusten_mask_select <- backbone/aff.py:320,323 (topk(sorted=False), nonzero)
 *
 * Conventions
 *   - Plain C: raw DEVICE pointers, sizes, element strides, a CUDA stream handle (cudaStream_t passed as void*).
 *     No torch types.  All calls are asynchronous on `stream`; none synchronises the host.
 *   - Return value: 0 = ok; > 0 = cudaError_t of the failing launch; < 0 = CLUSTEN_E* argument error.
 *     clusten_last_error() returns a thread-local message for the last non-zero return.
 *   - dtype: CLUSTEN_F32 / CLUSTEN_F16 / CLUSTEN_BF16 for all floating tensors of a call; accumulation is fp32.
 *   - "rows" tensors (q, k, v, feat, and their gradients) are addressed as
 *         base + b*sb + h*sh + n*sn + c      (strides in ELEMENTS, innermost stride must be 1)
 *     so the non-contiguous [B,N,H,C] -> [B,H,N,C] views the reference model produces (aff.py:111-113) are
 *     consumed and produced without the .contiguous() copies of clusten.py:25-33.
 *   - neighbour index tensors are int64 [B, Nq, M] contiguous, exactly what the reference passes
 *     (clusten.py:29); values must lie in [0, Nk) (unchecked, as in the reference).
 *   - Outputs are fully overwritten (callers may allocate them uninitialised).
 */
#ifndef CLUSTEN_B200_H
#define CLUSTEN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CLUSTEN_ABI_VERSION 6

enum { CLUSTEN_F32 = 0, CLUSTEN_F16 = 1, CLUSTEN_BF16 = 2 };

enum {
    CLUSTEN_EINVAL = -1,      /* bad size / null pointer */
    CLUSTEN_EDTYPE = -2,      /* unknown dtype */
    CLUSTEN_EUNSUPPORTED = -3,/* shape outside the supported range (message says which) */
    CLUSTEN_EWORKSPACE = -4   /* workspace too small */
};

int clusten_abi_version(void);
const char *clusten_last_error(void);
/* number of CUDA kernels this library has launched in this process (bench.py reports it as gpu_launches) */
long long clusten_kernel_launches(void);

/* ---- inverse neighbour list (CSR over key rows), built once per index tensor and reused by every backward ----
 * offsets: int32 [B, Nk+1]; entries: uint32 [B, Nq*M], entry = (i << 8) | j  (query i, slot j), ascending per row.
 * Requires M <= 256 and Nq < 2^24.  Deterministic (stable radix sort, no atomics on the data path). */
size_t clusten_csr_workspace_bytes(int B, int Nq, int M, int Nk);
int clusten_csr_build(const int64_t *nbhd_idx, int B, int Nq, int M, int Nk,
                      int32_t *offsets, uint32_t *entries, void *workspace, size_t workspace_bytes,
                      const void *pack /* or NULL: when given, the build is skipped on the device unless the pack routes to the generic kernels */,
                      void *stream);

/* ---- tile pack: per-index-tensor structure behind the tensor-core ("tile-union") kernels.  Built once per index tensor
 * (it is constant across the blocks of an AFF stage, aff.py:487-493) and passed as `pack` to the QK / AV entry points;
 * pack == NULL selects the generic row-gather kernels.  The pack carries a device-side flag: index tensors without
 * octet structure (M % 8 != 0, impure runs, too little locality) fall back to the generic kernels with no host sync. */
size_t clusten_pack_bytes(int B, int Nq, int M, int Nk);
/* mask (uint8 [B,Nq,M], 0 = padded neighbour slot, aff.py:480; or NULL): ONLY for packs handed to the fused attention entry
 * points (clusten_attn_fwd / _bwd and the clusten_scatter_rows calls behind them), which apply the same mask: masked entries
 * become wildcards, so the padded last cluster of point_utils.py:282-283 stays on the tensor-core path.  Packs for
 * clusten_qk_* / clusten_av_* must be built with mask == NULL (those ops honour every index literally). */
int clusten_pack_build(const int64_t *nbhd_idx, const uint8_t *mask, int B, int Nq, int M, int Nk, void *pack, size_t pack_bytes,
                       void *stream);
/* inverse lists (key octet -> referencing tiles) of an already built pack: needed by the backward entry points only */
int clusten_pack_inverse(void *pack, size_t pack_bytes, int B, int Nq, int M, int Nk, void *stream);

/* ---- QK: attn[b,h,i,j] = sum_c q[b,h,i,c] * k[b,h,idx[b,i,j],c]            (clustenqk_cuda_kernel.cu:38-45) */
int clusten_qk_fwd(const void *q, const void *k, const int64_t *nbhd_idx, const void *pack /* or NULL */,
                   void *attn /* [B,H,Nq,M] contiguous */,
                   int B, int H, int Nq, int Nk, int C, int M,
                   int64_t q_sb, int64_t q_sh, int64_t q_sn, int64_t k_sb, int64_t k_sh, int64_t k_sn,
                   int dtype, void *stream);
/* d_q[b,h,i,:] = sum_j d_attn[b,h,i,j] k[b,h,idx,:];  d_k[b,h,r,:] = sum_{(i,j): idx[b,i,j]=r} d_attn[b,h,i,j] q[b,h,i,:]
 *                                                                        (clustenqk_cuda_kernel.cu:118-128) */
int clusten_qk_bwd(const void *d_attn /* [B,H,Nq,M] contiguous */, const void *q, const void *k,
                   const int64_t *nbhd_idx, const int32_t *csr_offsets, const uint32_t *csr_entries,
                   const void *pack /* or NULL */, void *d_q, void *d_k,
                   int B, int H, int Nq, int Nk, int C, int M,
                   int64_t q_sb, int64_t q_sh, int64_t q_sn, int64_t k_sb, int64_t k_sh, int64_t k_sn,
                   int64_t dq_sb, int64_t dq_sh, int64_t dq_sn, int64_t dk_sb, int64_t dk_sh, int64_t dk_sn,
                   int dtype, void *stream);

/* ---- AV: feat[b,h,i,c] = sum_j attn[b,h,i,j] * v[b,h,idx[b,i,j],c]          (clustenav_cuda_kernel.cu:40-46)
 * attn is addressed base + b*a_sb + h*a_sh + i*a_sn + j (so the attn[..., :-1] slice of aff.py:146 needs no copy). */
int clusten_av_fwd(const void *attn, const void *v, const int64_t *nbhd_idx, const void *pack /* or NULL */, void *feat,
                   int B, int H, int Nq, int Nk, int C, int M,
                   int64_t a_sb, int64_t a_sh, int64_t a_sn, int64_t v_sb, int64_t v_sh, int64_t v_sn,
                   int64_t f_sb, int64_t f_sh, int64_t f_sn, int dtype, void *stream);
/* d_attn[b,h,i,j] = sum_c v[b,h,idx,c] d_feat[b,h,i,c];  d_v[b,h,r,:] = sum_{(i,j)->r} attn[b,h,i,j] d_feat[b,h,i,:]
 *                                                                 (clustenav_cuda_kernel.cu:117-123,152-156) */
int clusten_av_bwd(const void *d_feat, const void *attn, const void *v,
                   const int64_t *nbhd_idx, const int32_t *csr_offsets, const uint32_t *csr_entries,
                   const void *pack /* or NULL */, void *d_attn /* [B,H,Nq,M] contiguous */, void *d_v,
                   int B, int H, int Nq, int Nk, int C, int M,
                   int64_t df_sb, int64_t df_sh, int64_t df_sn, int64_t a_sb, int64_t a_sh, int64_t a_sn,
                   int64_t v_sb, int64_t v_sh, int64_t v_sn, int64_t dv_sb, int64_t dv_sh, int64_t dv_sn,
                   int dtype, void *stream);

/* ---- fused ClusterAttention core, forward (aff.py:114-155): QK + position bias + cluster mask + blank
 * token + softmax + AV in one kernel.
 *   logit[b,h,i,j] = q[b,h,i,:].k[b,h,idx[b,i,j],:] + bias_tab[bias_idx[b,i,j]*H + h] + (mask && !mask[b,i,j] ? -100 : 0), j < M
 *   logit[b,h,i,M] = q[b,h,i,:].blank_k[h*C:(h+1)*C];   p = softmax(logit[b,h,i,0..M])
 *   out[b,h,i,:]   = sum_j p[j] v[b,h,idx[b,i,j],:] + p[M] blank_v[h*C:(h+1)*C]
 * q is expected pre-scaled (aff.py:104).  q/k/v/out: strided rows like clusten_qk_fwd; bias_tab fp32 [R,H]; bias_idx int32
 * [B,Nq,M]; mask uint8 [B,Nq,M] or NULL; blank_k/blank_v: [H*C] of the call's dtype; probs: fp32 [B,H,Nq,M+1] or NULL
 * (the softmax output); lse: fp32 [B,H,Nq] or NULL (log-sum-exp of the M+1 logits, what clusten_attn_bwd recomputes the
 * probabilities from).  pack may be NULL (generic kernel only). */
int clusten_attn_fwd(const void *q, const void *k, const void *v, const int64_t *nbhd_idx, const void *pack,
                     const float *bias_tab, const int32_t *bias_idx, const uint8_t *mask,
                     const void *blank_k, const void *blank_v, void *out, float *probs, float *lse,
                     int B, int H, int Nq, int Nk, int C, int M,
                     int64_t q_sb, int64_t q_sh, int64_t q_sn, int64_t k_sb, int64_t k_sh, int64_t k_sn,
                     int64_t v_sb, int64_t v_sh, int64_t v_sn, int64_t o_sb, int64_t o_sh, int64_t o_sn,
                     int dtype, void *stream);

/* ---- fused ClusterAttention core, backward (fp16 / bf16): token-tile part.  Recomputes p = exp(logit - lse) and writes
 *   d_q (strided rows), probs = p[:, :, :, 0:M] and d_logits = p (dp - D) as [B,H,Nq,M] tensors of the call's dtype, and the
 *   blank-token column p_blank / ds_blank as fp32 [B,H,Nq].  The caller finishes with
 *   d_k = clusten_scatter_rows(d_logits, q), d_v = clusten_scatter_rows(probs, d_out), d_bias_tab = clusten_table_grad(d_logits,
 *   bias_idx), d_blank_k[h] = sum_i ds_blank q_i, d_blank_v[h] = sum_i p_blank d_out_i.   (aff.py:114-155 backwards) */
int clusten_attn_bwd(const void *d_out, const void *out, const float *lse, const void *q, const void *k, const void *v,
                     const int64_t *nbhd_idx, const void *pack, const float *bias_tab, const int32_t *bias_idx,
                     const uint8_t *mask, const void *blank_k, const void *blank_v,
                     void *d_q, void *probs, void *d_logits, float *p_blank, float *ds_blank,
                     int B, int H, int Nq, int Nk, int C, int M,
                     int64_t q_sb, int64_t q_sh, int64_t q_sn, int64_t k_sb, int64_t k_sh, int64_t k_sn,
                     int64_t v_sb, int64_t v_sh, int64_t v_sn, int64_t do_sb, int64_t do_sh, int64_t do_sn,
                     int64_t o_sb, int64_t o_sh, int64_t o_sn, int64_t dq_sb, int64_t dq_sh, int64_t dq_sn,
                     int dtype, void *stream);
/* ---- the fused core with the relative-position bias COMPUTED from token positions instead of gathered from a table
 * (opt-in: CLUSTEN_INKERNEL_BIAS=1 in the Python layer; see DESIGN.md section 7 for its validation status):
 *   bias[b,h,i,j] = pe_weight[h,:] . feat(rel) + pe_bias[h],  rel = trunc(clamp(pos_k[idx[b,i,j]] - (pos_q[i] - 511), 0, 1022)) - 511,
 *   feat = (dx, dy, dist, dy/dist, dx/dist) with the centre zeroed -- exactly pos_embed(pre_table)[pe_idx] of aff.py:17-31,129-132,
 *   481-485.  pos_q [B,Nq,2] / pos_k [B,Nk,2] fp32 (x, y), pe_weight fp32 [H,5], pe_bias fp32 [H] or NULL; everything else as in
 *   clusten_attn_fwd / clusten_attn_bwd.  Backward: pe_grad_parts fp32 [1024][H][6], accumulated INTO (caller zeroes it and sums
 *   over the first dimension): [.., h, 0:5] = d_pe_weight[h,:], [.., h, 5] = d_pe_bias[h]. */
int clusten_attn_pos_fwd(const void *q, const void *k, const void *v, const int64_t *nbhd_idx, const void *pack,
                         const float *pos_q, const float *pos_k, const float *pe_weight, const float *pe_bias,
                         const uint8_t *mask, const void *blank_k, const void *blank_v, void *out, float *probs, float *lse,
                         int B, int H, int Nq, int Nk, int C, int M,
                         int64_t q_sb, int64_t q_sh, int64_t q_sn, int64_t k_sb, int64_t k_sh, int64_t k_sn,
                         int64_t v_sb, int64_t v_sh, int64_t v_sn, int64_t o_sb, int64_t o_sh, int64_t o_sn,
                         int dtype, void *stream);
int clusten_attn_pos_bwd(const void *d_out, const void *out, const float *lse, const void *q, const void *k, const void *v,
                         const int64_t *nbhd_idx, const void *pack, const float *pos_q, const float *pos_k,
                         const float *pe_weight, const float *pe_bias,
                         const uint8_t *mask, const void *blank_k, const void *blank_v,
                         void *d_q, void *probs, void *d_logits, float *p_blank, float *ds_blank, float *pe_grad_parts,
                         int B, int H, int Nq, int Nk, int C, int M,
                         int64_t q_sb, int64_t q_sh, int64_t q_sn, int64_t k_sb, int64_t k_sh, int64_t k_sn,
                         int64_t v_sb, int64_t v_sh, int64_t v_sn, int64_t do_sb, int64_t do_sh, int64_t do_sn,
                         int64_t o_sb, int64_t o_sh, int64_t o_sn, int64_t dq_sb, int64_t dq_sh, int64_t dq_sn,
                         int dtype, void *stream);

/* blank-token parameter gradients of the fused core (aff.py:138-146 backwards), 16-bit q / d_out as strided [B,H,N,C] views,
 * dS_blank / P_blank fp32 [B,H,N] as written by clusten_attn_bwd; d_blank_k / d_blank_v fp32 [H*C], accumulated INTO
 * (caller zeroes them).  One pass over q and d_out instead of two skinny GEMMs. */
int clusten_blank_grad(const void *q, const void *d_out, const float *dS_blank, const float *P_blank,
                       float *d_blank_k, float *d_blank_v, int B, int H, int N, int C,
                       int64_t q_sb, int64_t q_sh, int64_t q_sn, int64_t g_sb, int64_t g_sh, int64_t g_sn,
                       int dtype, void *stream);

/* out[b,h,r,:] = sum_{(i,j): idx[b,i,j]=r} w[b,h,i,j] x[b,h,i,:]: the scatter half of clusten_qk_bwd / clusten_av_bwd on its own
 * (deterministic: inverse lists, no atomics).  w addressed base + b*w_sb + h*w_sh + i*w_sn + j. */
int clusten_scatter_rows(const void *w, const void *x, const int32_t *csr_offsets, const uint32_t *csr_entries,
                         const void *pack /* or NULL */, void *out, int B, int H, int Nq, int Nk, int C, int M,
                         int64_t w_sb, int64_t w_sh, int64_t w_sn, int64_t x_sb, int64_t x_sh, int64_t x_sn,
                         int64_t o_sb, int64_t o_sh, int64_t o_sn, int dtype, void *stream);

/* ---- relative-position table lookup (aff.py:129-132, 346-349) restricted to the table rows a stage references:
 *   gather: out[e,c] = tab[inv[e],c]  (e < n, c < CH; tab/out of the call's dtype; inv int32 or int64)
 *   grad  : d_tab[r,c] += sum_{e: inv[e]=r} d_out[e,c], fp32, d_tab [U,CH] must be zeroed by the caller; d_out addressed
 *           as base + (e / n_per)*d_sb + (e % n_per)*d_se + c*d_sc (element strides) so permuted gradients need no copy.
 *   Replaces ATen index / index_put_(accumulate) on this path; the gradient uses fp32 atomics (summation order is not fixed). */
int clusten_table_gather(const void *tab, const void *inv, int inv_is_i64, void *out, int64_t n, int U, int CH,
                         int dtype, void *stream);
/* Rows of the reference's pre_table (aff.py:21-31) for table indices rows[0:n] (int64, row = (dy + 511) * 1023 + dx + 511):
 * out[e,:] = (dx, dy, dist, dy / dist, dx / dist), the 0 / 0 centre zeroed; fp32 [n,5].  Same IEEE operations as the torch formulation. */
int clusten_rel_pos_features(const int64_t *rows, float *out, int64_t n, void *stream);
/* U_dev (device int32 scalar or NULL): number of table rows actually referenced when U is only an upper bound (the count
 * clusten_stage_prepare leaves on the device) -- lets the caller skip the device->host read of U. */
int clusten_table_grad(const void *d_out, const void *inv, int inv_is_i64, float *d_tab, int64_t n, int U, const int32_t *U_dev,
                       int CH, int64_t n_per, int64_t d_sb, int64_t d_se, int64_t d_sc, int dtype, void *stream);

/* ---- LayerNorm over the channel dimension of [R, C] token rows, C <= 1024 (aff.py:196-199,258,617-620): one warp per row.
 * x / y (and d_y) may be fp32 / fp16 / bf16 independently; gamma, beta, mean, rstd, d_gamma, d_beta are fp32.  mean / rstd may
 * both be NULL in inference.  d_gamma / d_beta are accumulated into (caller zeroes them; fp32 atomics across CTAs). */
int clusten_layer_norm_fwd(const void *x, const float *gamma, const float *beta, void *y, float *mean, float *rstd,
                           int64_t R, int C, float eps, int x_dtype, int y_dtype, void *stream);
int clusten_layer_norm_bwd(const void *d_y, const void *x, const float *gamma, const float *mean, const float *rstd,
                           void *d_x /* x's dtype */, float *d_gamma, float *d_beta, int64_t R, int C, int x_dtype, int g_dtype,
                           void *stream);

/* ---- stage preparation (aff.py:475-485 in one pass + the restriction of the 1023^2-row position table to the referenced rows):
 *   member_idx[b,i,c*m+r] = member[b, nearest[b,i,c], r]   (int64 [B,n,nnc*m]);   mask64 / mask8 = the same gather of
 *   cluster_mask (either may be NULL; both ignored when cluster_mask is NULL);
 *   pe_idx[b,i,j] = rel.y*1023 + rel.x with rel = clamp(pos[member_idx] - (pos[i] - 511), 0, 1022)          (int32)
 *   uniq[0:U] = ascending distinct pe_idx values (what torch.unique returns), bias_idx = rank of pe_idx in uniq (int32),
 *   *count = U (device scalar; uniq is written up to uniq_cap entries).  No sort: presence map + scan. */
size_t clusten_prepare_workspace_bytes(void);
int clusten_stage_prepare(const int64_t *nearest /* [B,n,nnc] */, const int64_t *member /* [B,k,m] */,
                          const int64_t *cluster_mask /* [B,k,m] or NULL */, const float *pos /* [B,n,2] */,
                          int B, int n, int k, int m, int nnc,
                          int64_t *member_idx, int64_t *mask64, uint8_t *mask8, int32_t *pe_idx, int32_t *bias_idx,
                          int32_t *uniq, int uniq_cap, int32_t *count, void *workspace, size_t workspace_bytes, void *stream);
/* The table-row restriction alone, for index tensors that arrive as int64 table rows (PointConv, msdeformattn_pc.py:305-306):
 * uniq[0:count] = the ascending distinct rows of pe_idx (what torch.unique returns, without the sort), inverse[e] = rank of
 * pe_idx[e]; uniq is filled up to uniq_cap rows, the count stays on the device (no host read).  Same workspace as above. */
int clusten_table_rank(const int64_t *pe_idx, int64_t total, int32_t *inverse, int32_t *uniq, int uniq_cap, int32_t *count,
                       void *workspace, size_t workspace_bytes, void *stream);

/* ---- Linear(F -> H) over the rows of the relative-position feature table: out[r,h] = bias[h] + sum_j feat[r,j] * weight[h,j]
 * -- `self.pos_embed(pre_table)` of backbone/aff.py:101,129 on the rows a stage references.  fp32; F <= 8, H <= 32;
 * count = device int32 scalar (rows actually referenced, see clusten_stage_prepare) or NULL (= R): rows >= *count are written
 * as zeros forward and ignored backward.  Backward: d_weight[h,j] += sum_r d_out[r,h] * feat[r,j], d_bias[h] += sum_r
 * d_out[r,h] (accumulated INTO with fp32 atomics, one per CTA and weight; d_bias may be NULL); feat gets no gradient
 * (the table is a constant, aff.py:17-31). */
int clusten_table_linear_fwd(const float *feat, const float *weight, const float *bias, float *out, int R, int F, int H,
                             const int32_t *count, void *stream);
int clusten_table_linear_bwd(const float *d_out, const float *feat, float *d_weight, float *d_bias, int R, int F, int H,
                             const int32_t *count, void *stream);

/* ---- fp32 Linear layer on the tensor cores (3xTF32 split, fp32-level accuracy): y[r,n] = sum_k x[r,k] * weight[n,k] + bias[n]
 * -- the q / kv / proj / fc1 / fc2 layers of the block (aff.py:62-70,103-106) in fp32 inference.  x [R,K] with row stride ldx,
 * weight [N,K] contiguous, bias [N] or NULL, y [R,N] with row stride ldy; K % 32 == 0, N % 2 == 0, 16-byte aligned x / weight rows.
 * Opt-in in the Python layer (CLUSTEN_TC_LINEAR=1); see DESIGN.md section 7 for its validation status. */
int clusten_linear_f32(const float *x, const float *weight, const float *bias, float *y, int64_t R, int K, int N,
                       int64_t ldx, int64_t ldy, void *stream);

/* ---- fp32 Linear layer on the 5th-generation tensor cores (tcgen05.mma kind::tf32 with TMEM accumulators, TMA-fed, one persistent
 * CTA per SM; 3xTF32 split, fp32-level accuracy): y = epilogue(x . weight^T + bias) -- the q / kv / proj / fc1 / fc2 layers of the
 * block (backbone/aff.py:62-70,103-106,181-189) and the merge Linear (aff.py:282) in fp32 inference, with the element-wise line that
 * follows each of them in the reference folded into the epilogue:
 *   epi 0   y = acc + bias, columns < alpha_cols then multiplied by alpha     (`q = self.q(feat) * self.scale`, aff.py:107-108)
 *   epi 1   y = GELU(acc + bias), exact erf form                               (`self.act(self.fc1(x))`, aff.py:45-46)
 *   epi 2   y = res + gamma * (acc + bias), gamma [N] or NULL (= 1)            (`x = shortcut + gamma * x`, aff.py:230,236)
 * clusten_tf32_split writes the two weight operands once per weight: hi = w rounded to TF32, lo = (w - hi) rounded to TF32.
 * x [R,K] row stride ldx, w_hi / w_lo [N,K] contiguous, y [R,N] row stride ldy, res [R,N] row stride ldres; K % 32 == 0,
 * N % 4 == 0, 16-byte aligned rows (else CLUSTEN_EUNSUPPORTED).  chain: K chunks of 32 summed inside one tensor-memory
 * accumulator before it is added to the fp32 running sum in registers (<= 0: default).  ln_mean / ln_rstd [R], ln_gamma / ln_beta [K]
 * (all NULL: off): the `norm1` / `norm2` / `norm` LayerNorm in front of the layer (aff.py:196-199,258) applied to the rows of x while
 * they are staged, from the statistics clusten_layer_norm_fwd writes when called with y = NULL.
 * w_fp16 = 1: the fp16 form of the split -- clusten_f16_split writes hi / lo [N,K] fp16 of w[n,:] * s_n (s_n = the power of two that
 * brings amax[n] = max |w[n,:]| to [512, 1024)) and 1 / s_n to inv_scale[n]; the kernel splits x into fp16 hi / lo too (11 + 11
 * significand bits like the TF32 form) and multiplies with kind::f16: half the MMAs and weight bytes.  x is NOT rescaled: |x| beyond
 * 65504 saturates and elements below 2^-14 keep an absolute error of 2^-25, so this form is meant for inputs of known scale -- the
 * Python layer uses it for the LayerNorm-fused layers (normalised rows) and the TF32 form elsewhere. */
int clusten_tf32_split(const float *w, float *hi, float *lo, int64_t n, void *stream);
int clusten_f16_split(const float *w, void *hi, void *lo, int64_t N, int K, const float *amax, float *inv_scale, void *stream);
int clusten_linear_tc_f32(const float *x, const void *w_hi, const void *w_lo, const float *bias, const float *res,
                          const float *gamma, float *y, int64_t R, int K, int N, int64_t ldx, int64_t ldy, int64_t ldres,
                          int epi, float alpha, int alpha_cols, int chain, const float *ln_mean, const float *ln_rstd,
                          const float *ln_gamma, const float *ln_beta, int w_fp16, const float *w_inv_scale, void *stream);

/* ---- column sum: out[c] += sum_r x[r*ld + c] (fp32 accumulation INTO out; caller zeroes it).  The bias gradient of the
 * backbone's Linear layers (grad_bias = grad_output.sum(0)); x fp32 / fp16 / bf16, C and ld multiples of 16 bytes. */
int clusten_col_sum(const void *x, float *out, int64_t R, int C, int64_t ld, int dtype, void *stream);

/* ---- residual add with layer scale and stochastic depth: out[b,i,c] = res[b,i,c] + x[b,i,c] * gamma[c] * sample_scale[b]
 * -- the `x = shortcut + drop_path(gamma * x)` lines of the reference block (backbone/aff.py:230,236; DropPath = timm 0.6.12).
 * res / x / out contiguous [B, rows, C], C % 4 == 0, C <= 1024; gamma fp32 [C] or NULL (= 1), sample_scale fp32 [B] or
 * NULL (= 1).  dtype triples (res, x, out): (f32,f32,f32) (f32,h,f32) (h,h,h) (h,h,f32) for h = fp16 / bf16; products and
 * the sum are rounded one by one in fp32 (the fp32 case equals the op-by-op ATen result bit for bit).
 * Backward: d_x = d_out * gamma * sample_scale (x's dtype; d_x may be NULL), d_gamma[c] += sum d_out * x * sample_scale
 * (fp32, accumulated INTO d_gamma with one atomic per CTA and channel; NULL = not wanted, then x may be NULL too);
 * d_res = d_out needs no kernel.  (d_out, x) dtypes: (f32,f32) (f32,h) (h,h). */
int clusten_scale_residual_fwd(const void *res, const void *x, const float *gamma, const float *sample_scale, void *out,
                               int64_t B, int64_t rows, int C, int res_dtype, int x_dtype, int out_dtype, void *stream);
int clusten_scale_residual_bwd(const void *d_out, const void *x, const float *gamma, const float *sample_scale, void *d_x,
                               float *d_gamma, int64_t B, int64_t rows, int C, int g_dtype, int x_dtype, void *stream);

/* ---- WF plan (optional, 16-bit tensor-core kernels): per index tensor, built once and passed to clusten_wf_fwd / _bwd.
 * Holds the token processing order (kept tokens arrive in top-k order, aff.py:320-324; neighbouring tokens re-use rows
 * out of L1 when processed together) and, when M % 8 == 0, the per-octet reference lists that turn the d_f scatter of
 * clustenwf_cuda_kernel.cu:120-131 into dense per-octet products.  Opaque device buffer of clusten_wf_plan_bytes bytes.
 * When the plan's device-side flag routes d_f to the generic kernels, clusten_csr_build(..., pack = plan) builds the
 * inverse list, and skips it otherwise (same flag position as the tile pack). */
size_t clusten_wf_plan_bytes(int B, int Nq, int M, int Nk);
int clusten_wf_plan_build(const int64_t *nbhd_idx, int B, int Nq, int M, int Nk, void *plan, size_t plan_bytes, void *stream);

/* ---- WF: out[b,i,ic,c] = sum_j w[b,i,j,ic] * f[b,idx[b,i,j],c]             (clustenwf_cuda_kernel.cu:41-49)
 * w [B,Nq,M,IC] contiguous, f rows base + b*f_sb + n*f_sn + c, out [B,Nq,IC,C] contiguous.  IC in {1,2,4,8}. */
int clusten_wf_fwd(const void *w, const void *f, const int64_t *nbhd_idx, const void *plan /* or NULL */, void *out,
                   int B, int Nq, int Nk, int C, int M, int IC, int64_t f_sb, int64_t f_sn, int dtype, void *stream);
/* d_w[b,i,j,ic] = sum_c f[b,idx,c] d_out[b,i,ic,c];  d_f[b,r,:] = sum_{(i,j)->r} sum_ic w[b,i,j,ic] d_out[b,i,ic,:]
 *                                                                 (clustenwf_cuda_kernel.cu:120-131,161-165) */
int clusten_wf_bwd(const void *d_out, const void *w, const void *f,
                   const int64_t *nbhd_idx, const int32_t *csr_offsets, const uint32_t *csr_entries,
                   const void *plan /* or NULL */, void *d_w, void *d_f,
                   int B, int Nq, int Nk, int C, int M, int IC, int64_t f_sb, int64_t f_sn,
                   int64_t df_sb, int64_t df_sn, int dtype, void *stream);

/* ---- WEIGHTEDGATHER: out[b,i,c] = sum_k w[b,i,k] * f[b,idx[b,i,k],c]   (weighted_gather_cuda_kernel.cu:38-45);
 * argument order follows the reference (idx, weights, feat).  Same kernels as WF with IC = 1. */
int clusten_wg_fwd(const int64_t *nbhd_idx, const void *w, const void *f, void *out,
                   int B, int Nq, int Nk, int C, int K, int64_t f_sb, int64_t f_sn, int dtype, void *stream);
int clusten_wg_bwd(const void *d_out, const int64_t *nbhd_idx, const void *w, const void *f,
                   const int32_t *csr_offsets, const uint32_t *csr_entries, void *d_w, void *d_f,
                   int B, int Nq, int Nk, int C, int K, int64_t f_sb, int64_t f_sn,
                   int64_t df_sb, int64_t df_sn, int dtype, void *stream);

/* ---- MSDETRPC (point-cloud deformable attention, pixel decoder): feat[b,i,c] = sum_m attn[b,i,m] * sum_k nn_weight[b,i,m,k] *
 * val[b, nn_idx[b,i,m,k], c]   (clusten/src/msdetrpc_cuda.cpp:25-51, msdetrpc_cuda_kernel.cu:18-55); backward = d_nn_weight, d_attn
 * (msdetrpc_cuda_kernel.cu:138-181) and d_val by the inverse neighbour list of nn_idx viewed as [B,N,M*K] (clusten_csr_build)
 * instead of the reference's atomics (:113-131).  One pass each: the product weights are formed inside the kernels.
 * nn_idx int64 [B,N,M,K], nn_weight [B,N,M,K], attn [B,N,M], val strided [B,Nk,C] (unit inner stride), out [B,N,C]; M*K <= 256. */
int clusten_msdetrpc_fwd(const int64_t *nn_idx, const void *nn_weight, const void *attn, const void *val, void *out,
                         int B, int N, int Nk, int C, int M, int K, int64_t v_sb, int64_t v_sn, int dtype, void *stream);
int clusten_msdetrpc_bwd(const void *d_out, const int64_t *nn_idx, const void *nn_weight, const void *attn, const void *val,
                         const int32_t *csr_offsets, const uint32_t *csr_entries, void *d_weight, void *d_attn, void *d_val,
                         int B, int N, int Nk, int C, int M, int K, int64_t v_sb, int64_t v_sn, int64_t dv_sb, int64_t dv_sn,
                         int dtype, void *stream);

/* ---- kNN (2-D, fp32): the k nearest database points of each query, ascending distance, ties -> lowest index;
 * dist = sqrt_rn(fl(dx*dx) + fl(dy*dy)) without FMA contraction.  idx_out int64 [B,Nq,k]; dist_out fp32 [B,Nq,k] or NULL.
 * 1 <= k <= 16, k <= Ndb.                                                    (point_utils.py:41-60) */
int clusten_knn(const float *query /* [B,Nq,2] */, const float *database /* [B,Ndb,2] */,
                int B, int Nq, int Ndb, int k, int64_t *idx_out, float *dist_out, void *stream);

/* ---- balanced space-filling-curve clustering (point_utils.py:135-287, sf_type='', use_anchor=True, no_reorder=False)
 * pos fp32 [B,n,2]; outputs: pos_sorted fp32 [B,n,2], mean_pos fp32 [B,k,2], member_idx int64 [B,k,m],
 * cluster_mask int64 [B,k,m] (written only when k*m != n; may be NULL otherwise), pos_ranking int64 [B,n],
 * with k = ceil(n/m).  The sort is stable (ties -> lower original index). */
size_t clusten_sfc_workspace_bytes(int B, int n);
int clusten_sfc_cluster(const float *pos, int B, int n, int m, int h, int w,
                        float *pos_sorted, float *mean_pos, int64_t *member_idx, int64_t *cluster_mask,
                        int64_t *pos_ranking, void *workspace, size_t workspace_bytes, void *stream);

/* ---- selection primitives of the adaptive downsampling (aff.py:320-324)
 * topk_select: idx_out[b, 0:k] = first k of a stable DESCENDING sort of score[b,:] (fp32 [B,n]);
 *              written at idx_out + b*out_stride.
 * mask_select: idx_out[b, 0:count] = ascending indices i with mask[b,i] != 0 (fp32 [B,n]); exactly `count` slots are
 *              written per row (surplus indices dropped, missing slots filled with 0). */
size_t clusten_topk_workspace_bytes(int B, int n);
int clusten_topk_select(const float *score, int B, int n, int k, int64_t *idx_out, int64_t out_stride,
                        void *workspace, size_t workspace_bytes, void *stream);
int clusten_mask_select(const float *mask, int B, int n, int count, int64_t *idx_out, int64_t out_stride, void *stream);
/* merge_scores: the scores those two select from (aff.py:292-315) in one pass: grid_prob = all(pos.long() % s == 0) with s = stride, or
 * per token 2 ** (ceil(log2(min_dist[.., 1])) + 1) when min_dist (fp32 [B,n,dist_stride], the kNN-2 distances of aff.py:299) is given;
 * final_prob = grid_prob + learned_prob * alpha (learned_prob fp32 [B,n] or NULL) [+ reserve_mask * (-100)];
 * reserve_mask = all(pos.long() % (2 stride) == 0) (written when reserve_on).  fp32, every operation rounded separately. */
int clusten_merge_scores(const float *pos /* [B,n,2] */, const float *min_dist, int dist_stride, const float *learned_prob, float alpha,
                         int stride, int reserve_on, float *final_prob, float *reserve_mask, int B, int n, void *stream);

/* ---- row gather: out[b,i,:] = src[b, idx[b,i], :] for i < n_out -- the `x.gather(index=idx.expand(-1,-1,c), dim=1)` row
 * reorders / selections of backbone/aff.py:332,335,340,471 (features into cluster order, kept tokens' positions, member rows and
 * masks after a merge).  Rows are opaque: row_bytes bytes each (fp32 features, int64 member rows, uint8 masks alike), copied as
 * 16 / 8 / 4 / 1-byte pieces by alignment.  src [B,n_src,row_bytes], idx int64 [B,n_out], out [B,n_out,row_bytes], all contiguous.
 * An index outside [0, n_src) copies nothing and sets *bad (device int, or NULL) to 1. */
int clusten_gather_rows(const void *src, const int64_t *idx, void *out, int B, int n_src, int n_out, int row_bytes,
                        int *bad, void *stream);

/* ---- the stem, fp32 inference (PatchEmbed.forward, backbone/aff.py:527-530,549):
 * clusten_stem_conv_bn_gelu: y = GELU(BatchNorm_eval(Conv2d(3 -> OC, 3x3, stride 2, padding 1)(x))) -- `self.act1(self.bn(self.proj1(x)))`
 *   in one pass instead of four.  x [B,IC,H,W] NCHW, weight [OC,IC,3,3], bias [OC] or NULL, bn_mean / bn_var [OC] (running
 *   statistics), bn_weight / bn_bias [OC] or NULL; y [B,OC,(H+1)/2,(W+1)/2] NCHW, or pixel-major [B,(H+1)/2,(W+1)/2,OC] with
 *   channels_last = 1; contiguous fp32.  IC = 3 and OC in {16, 24, 32, 48, 64}; anything else returns CLUSTEN_EUNSUPPORTED (the
 *   caller keeps the four-pass formulation).
 * clusten_stem_im2col: the rows of the second convolution (3x3, stride 2, padding 1) over a pixel-major map mid [B,H,W,C]:
 *   A[(b,py,px), (ky*3+kx)*C + c] = mid[b, 2py-1+ky, 2px-1+kx, c] (0 outside), columns 9C .. Kp-1 zero; A [B*OH*OW, Kp] with
 *   OH = (H+1)/2, OW = (W+1)/2.  `self.proj2` then is clusten_linear_tc_f32 over A with the weight as [E, (ky,kx,c)] padded to Kp,
 *   and its output rows are the tokens [B, OH*OW, E] of `x.flatten(2).transpose(1, 2)`.  C % 4 == 0, Kp % 4 == 0, Kp >= 9C. */
int clusten_stem_conv_bn_gelu(const float *x, const float *weight, const float *bias, const float *bn_mean, const float *bn_var,
                              const float *bn_weight, const float *bn_bias, float eps, float *y, int B, int IC, int H, int W,
                              int OC, int channels_last, void *stream);
int clusten_stem_im2col(const float *mid, float *A, int B, int H, int W, int C, int Kp, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* CLUSTEN_B200_H */
