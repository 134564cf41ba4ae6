// Fused ClusterAttention core, BACKWARD (sm_100a, 16-bit types): the gradient of clusten_attn_fwd
// (mask2former/modeling/backbone/aff.py:114-155 run backwards) without ever materialising the fp32 [B,H,N,M+1] tensors
// autograd keeps for the reference's seven glue passes.  Flash-style: the probabilities are recomputed from the
// log-sum-exp the forward saved.
//
//   s_ij  = q_i . k_idx(i,j) + bias_tab[bias_idx(i,j), h] + mask term        p_ij = exp(s_ij - lse_i)            j < M
//   s_iM  = q_i . blank_k[h]                                                 p_iM = exp(s_iM - lse_i)
//   dp_ij = dO_i . v_idx(i,j),  dp_iM = dO_i . blank_v[h],  D_i = dO_i . O_i (= sum_j p_ij dp_ij)
//   ds_ij = p_ij (dp_ij - D_i)
//   d_q_i = sum_j ds_ij k_idx(i,j) + ds_iM blank_k[h]
//   d_k_r = sum_{(i,j)->r} ds_ij q_i        d_v_r = sum_{(i,j)->r} p_ij dO_i           (scat2 kernels on the P / dS tensors)
//   d_bias_tab[r,h] = sum_{(i,j): bias_idx = r} ds_ij                                  (table.cu on dS)
//   d_blank_k[h] = sum_i ds_iM q_i,  d_blank_v[h] = sum_i p_iM dO_i                    (caller, from the [B,H,N] outputs)
//
// This file holds the token-tile kernel: one warp per (16-token tile, head) walks the tile's union of key octets twice
// (tile.cuh): pass 1 runs the two dot shapes (q.K^T and dO.V^T) on the tensor cores straight from 128-bit row loads and
// parks the selected 16x8 blocks in shared memory; an elementwise phase turns them into P and dS (16-bit, written once to
// global memory as contiguous tiles for the scatter kernels); pass 2 runs the axpy shape dS.K with K staged key-major by
// cp.async and read back with ldmatrix.trans.  Impure tokens (tile.cuh) are computed one (token, head) per warp by the warps
// that finish first, exactly as in clusten_tile2.cu.
#include "posbias.cuh"
#include "t2.cuh"

namespace clusten {
namespace fb {

using namespace t2;

struct BwdArgs {
    const void *q, *k, *v, *dO, *O;
    const int64_t *idx;
    const float *bias_tab;
    const int32_t *bias_idx;
    const uint8_t *mask;
    const void *blank_k, *blank_v;
    const float *lse;
    void *dq, *P, *dS;
    float *Pb, *dSb;
    int B, H, Nq, Nk, C, M;
    int q_sh, q_sn, k_sh, k_sn, v_sh, v_sn, do_sh, do_sn, o_sh, o_sn, dq_sh, dq_sn;
    int64_t q_sb, k_sb, v_sb, do_sb, o_sb, dq_sb;
    int smem_per_warp;
};
// position-bias variant (PB kernels, clusten_attn_pos_bwd; posbias.cuh): bias from positions, pos_embed gradient into pe_parts.
// A separate type so that the kernels of the table variant keep their parameter block (and code) exactly as validated.
struct BwdArgsPB : BwdArgs {
    const float *pos_q, *pos_k, *pe_w, *pe_b;            // [B,Nq,2], [B,Nk,2], [H,5], [H] or NULL
    float *pe_parts;                                     // [PB_PARTS][H][6] fp32, accumulated into
};
template <bool PB> using ArgsOf = std::conditional_t<PB, BwdArgsPB, BwdArgs>;

template <typename T> __device__ __forceinline__ float pair_dot(uint32_t a, uint32_t b) {
    float2 fa, fb_;
    if constexpr (std::is_same<T, __half>::value) {
        fa = __half22float2(*reinterpret_cast<const __half2 *>(&a));
        fb_ = __half22float2(*reinterpret_cast<const __half2 *>(&b));
    } else {
        fa = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&a));
        fb_ = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&b));
    }
    return fmaf(fa.x, fb_.x, fa.y * fb_.y);
}
__device__ __forceinline__ float quad_sum(float x) {                 // over the 4 lanes (t) that share a fragment row
    x += __shfl_xor_sync(FULL, x, 1);
    x += __shfl_xor_sync(FULL, x, 2);
    return x;
}
template <typename T> __device__ __forceinline__ float2 unpack_pair(uint32_t a) {
    if constexpr (std::is_same<T, __half>::value) return __half22float2(*reinterpret_cast<const __half2 *>(&a));
    else return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&a));
}
__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(FULL, x, o);
    return x;
}

// One impure (token, head), one warp: everything of the header for this row, rows read as scalars / 16-byte chunks.
// `sc` = M floats of shared scratch.
template <typename T, bool PB = false>
__device__ __noinline__ void bwd_row(const ArgsOf<PB> &a, int b, int h, int i, float *sc, int lane) {
    const T *q = reinterpret_cast<const T *>(a.q) + b * a.q_sb + h * a.q_sh + (int64_t)i * a.q_sn;
    const T *dO = reinterpret_cast<const T *>(a.dO) + b * a.do_sb + h * a.do_sh + (int64_t)i * a.do_sn;
    const T *O = reinterpret_cast<const T *>(a.O) + b * a.o_sb + h * a.o_sh + (int64_t)i * a.o_sn;
    const T *kb = reinterpret_cast<const T *>(a.k) + b * a.k_sb + h * a.k_sh;
    const T *vb = reinterpret_cast<const T *>(a.v) + b * a.v_sb + h * a.v_sh;
    const T *bk = reinterpret_cast<const T *>(a.blank_k) + h * a.C;
    const T *bv = reinterpret_cast<const T *>(a.blank_v) + h * a.C;
    const int64_t row = ((int64_t)b * a.H + h) * a.Nq + i;
    const int64_t *irow = a.idx + ((int64_t)b * a.Nq + i) * a.M;
    const int32_t *bi = PB ? nullptr : a.bias_idx + ((int64_t)b * a.Nq + i) * a.M;
    const uint8_t *mk = a.mask ? a.mask + ((int64_t)b * a.Nq + i) * a.M : nullptr;
    const int M = a.M, C = a.C;
    PosBiasW pw = {};
    float2 pq = make_float2(0.f, 0.f);
    const float2 *PK = nullptr;
    float gacc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if constexpr (PB) {
        pw = pos_bias_load(a.pe_w, a.pe_b, h);
        pq = __ldg(reinterpret_cast<const float2 *>(a.pos_q) + (int64_t)b * a.Nq + i);
        PK = reinterpret_cast<const float2 *>(a.pos_k) + (int64_t)b * a.Nk;
    }
    const float lse = a.lse[row];
    float D = 0.f, sb = 0.f, dpb = 0.f;
    if (lane < C) {
        const float qc = to_f(q[lane]), dc = to_f(dO[lane]);
        D = dc * to_f(O[lane]);
        sb = qc * to_f(bk[lane]);
        dpb = dc * to_f(bv[lane]);
    }
    D = warp_sum(D); sb = warp_sum(sb); dpb = warp_sum(dpb);
    const float pb = __expf(sb - lse), dsb = pb * (dpb - D);
    if (lane == 0) { a.Pb[row] = pb; a.dSb[row] = dsb; }
    T *Prow = reinterpret_cast<T *>(a.P) + row * M, *dSrow = reinterpret_cast<T *>(a.dS) + row * M;
    for (int j = lane; j < M; j += 32) {
        const int64_t r = irow[j];
        const T *kr = kb + r * a.k_sn, *vr = vb + r * a.v_sn;
        float s = 0.f, dp = 0.f;
        for (int c = 0; c < C; ++c) {
            s = fmaf(to_f(q[c]), to_f(kr[c]), s);
            dp = fmaf(to_f(dO[c]), to_f(vr[c]), dp);
        }
        if constexpr (PB) s += pos_bias(pw, pq, __ldg(PK + r));
        else s += a.bias_tab[(int64_t)bi[j] * a.H + h];
        if (mk && !mk[j]) s += -100.f;
        const float p = __expf(s - lse);
        const T dsr = from_f<T>(p * (dp - D));
        Prow[j] = from_f<T>(p);
        dSrow[j] = dsr;
        sc[j] = to_f(dsr);                               // the rounded value: what the scatter kernels will multiply with
        if constexpr (PB) pos_bias_grad(gacc, pq, __ldg(PK + r), to_f(dsr));
    }
    if constexpr (PB) pos_bias_grad_flush(gacc, a.pe_parts, i, a.H, h, lane);
    __syncwarp();
    if (lane < C) {
        float acc = dsb * to_f(bk[lane]);
        for (int j = 0; j < M; ++j) acc = fmaf(sc[j], to_f(kb[irow[j] * a.k_sn + lane]), acc);
        reinterpret_cast<T *>(a.dq)[b * a.dq_sb + h * a.dq_sh + (int64_t)i * a.dq_sn + lane] = from_f<T>(acc);
    }
    __syncwarp();
}

template <typename T, int CH, int NT, bool PB>
__global__ void __launch_bounds__(128)
attn_bwd_tile_kernel(const ArgsOf<PB> a, const PackView pk) {
    extern __shared__ __align__(16) unsigned char dyn_fb[];
    if (pk.flags[0]) return;
    constexpr int NR = CH / 2;
    constexpr int ROWB = NT * 16 + 16;
    constexpr int KSTG = 16 * ROWB;                      // one stage: two octets of K, key-major
    constexpr int RPP = 32 / NT;
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int M = a.M, MP = M + 4, H = a.H, Nq = a.Nq, C = a.C;
    unsigned char *smw = dyn_fb + (size_t)(threadIdx.x >> 5) * a.smem_per_warp;
    float *S = reinterpret_cast<float *>(smw);           // [16][MP]: logits; [row][M] = D, [row][M+1] = lse
    float *DP = S + 16 * MP;                             // [16][MP]
    T *Pt = reinterpret_cast<T *>(DP + 16 * MP);         // [16][M]
    T *dSt = Pt + 16 * M;                                // [16][M]
    const uint32_t sK = (uint32_t)__cvta_generic_to_shared(dSt + 16 * M);
    const int tile = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (tile < pk.T) {
    const int b = blockIdx.y / H, h = blockIdx.y - b * H;
    const int bt = b * pk.T + tile, i0 = tile * TILE_TOK;
    const int U = pk.tile_u[bt];
    const int *octp = pk.tile_oct + bt * U_MAX;
    const int oc0 = octp[lane], oc1 = octp[32 + (lane & 15)];
    auto octet = [&](int u) { u = min(u, U - 1); return __shfl_sync(FULL, u < 32 ? oc0 : oc1, u & 31); };
    uint32_t impm = 0;
    {
        const uint4 v4 = __ldg(reinterpret_cast<const uint4 *>(pk.tok_imp + bt * TILE_TOK));
        if (v4.x | v4.y | v4.z | v4.w) {
            const uint32_t w[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4)
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) impm |= ((w[q4] >> (8 * k4)) & 1u) << (4 * q4 + k4);
        }
    }
    const bool cact = CH * t < C;
    const int cofs = cact ? CH * t : 0;
    const int ra = i0 + g, rb = ra + 8;
    const int rac = min(ra, Nq - 1), rbc = min(rb, Nq - 1);
    const bool va = ra < Nq && !((impm >> g) & 1u), vb = rb < Nq && !((impm >> (g + 8)) & 1u);
    const T *Qb = opaque(reinterpret_cast<const T *>(a.q) + b * a.q_sb + h * a.q_sh);
    const T *Kb = opaque(reinterpret_cast<const T *>(a.k) + b * a.k_sb + h * a.k_sh);
    const T *Vb = opaque(reinterpret_cast<const T *>(a.v) + b * a.v_sb + h * a.v_sh);
    const T *Gb = opaque(reinterpret_cast<const T *>(a.dO) + b * a.do_sb + h * a.do_sh);
    const T *Ob = opaque(reinterpret_cast<const T *>(a.O) + b * a.o_sb + h * a.o_sh);
    const int64_t row0 = ((int64_t)b * H + h) * Nq;
    // ---- fragments of q and dO (rows g, g+8), D = dO.O, blank-token terms ---------------------------------------------------
    uint32_t xa[NR], xb[NR], da[NR], db[NR];
    ld_chunk<CH * 2>(xa, at(Qb, rac * a.q_sn + cofs));
    ld_chunk<CH * 2>(xb, at(Qb, rbc * a.q_sn + cofs));
    ld_chunk<CH * 2>(da, at(Gb, rac * a.do_sn + cofs));
    ld_chunk<CH * 2>(db, at(Gb, rbc * a.do_sn + cofs));
    float dsb_a, dsb_b;                                  // ds of the blank token, rows g / g+8 (kept for the epilogue)
    {
        uint32_t oa[NR], ob[NR], bk[NR], bv[NR];
        ld_chunk<CH * 2>(oa, at(Ob, rac * a.o_sn + cofs));
        ld_chunk<CH * 2>(ob, at(Ob, rbc * a.o_sn + cofs));
        ld_chunk<CH * 2>(bk, reinterpret_cast<const T *>(a.blank_k) + h * C + cofs);
        ld_chunk<CH * 2>(bv, reinterpret_cast<const T *>(a.blank_v) + h * C + cofs);
        if (!cact) {
#pragma unroll
            for (int x = 0; x < NR; ++x) xa[x] = xb[x] = da[x] = db[x] = 0u;
        }
        float Da = 0.f, Db = 0.f, sa_ = 0.f, sb_ = 0.f, pa_ = 0.f, pb_ = 0.f;
#pragma unroll
        for (int x = 0; x < NR; ++x) {
            Da += pair_dot<T>(da[x], oa[x]); Db += pair_dot<T>(db[x], ob[x]);
            sa_ += pair_dot<T>(xa[x], bk[x]); sb_ += pair_dot<T>(xb[x], bk[x]);
            pa_ += pair_dot<T>(da[x], bv[x]); pb_ += pair_dot<T>(db[x], bv[x]);
        }
        Da = quad_sum(Da); Db = quad_sum(Db); sa_ = quad_sum(sa_); sb_ = quad_sum(sb_); pa_ = quad_sum(pa_); pb_ = quad_sum(pb_);
        const float la = a.lse[row0 + rac], lb = a.lse[row0 + rbc];
        const float pba = __expf(sa_ - la), pbb = __expf(sb_ - lb);
        dsb_a = pba * (pa_ - Da);
        dsb_b = pbb * (pb_ - Db);
        if (t == 0) {
            S[g * MP + M] = Da; S[g * MP + M + 1] = la;
            S[(g + 8) * MP + M] = Db; S[(g + 8) * MP + M + 1] = lb;
            if (va) { a.Pb[row0 + ra] = pba; a.dSb[row0 + ra] = dsb_a; }
            if (vb) { a.Pb[row0 + rb] = pbb; a.dSb[row0 + rb] = dsb_b; }
        }
    }
    const int8_t *sa = pk.slot_of + (bt * TILE_TOK + g) * U_MAX;
    const int8_t *sb = sa + 8 * U_MAX;
    // PB: the lane owns rows g / g+8 against keys 2t, 2t+1 of each octet in both passes (posbias.cuh)
    PosBiasW pw = {};
    float2 pqa = make_float2(0.f, 0.f), pqb = pqa;
    const float2 *PK = nullptr;
    if constexpr (PB) {
        pw = pos_bias_load(a.pe_w, a.pe_b, h);
        const float2 *PQ = reinterpret_cast<const float2 *>(a.pos_q) + (int64_t)b * Nq;
        pqa = __ldg(PQ + rac);
        pqb = __ldg(PQ + rbc);
        PK = reinterpret_cast<const float2 *>(a.pos_k) + (int64_t)b * a.Nk;
    }
    // ---- pass 1: S = q.K^T and DP = dO.V^T against every union octet, selected blocks -> shared memory ------------------------
    {
        const int klast = a.Nk - 1;                      // (mask-aware packs may hold the partial last octet: rows are clamped)
        float *Sa = S + g * MP + 2 * t, *Sb = S + (g + 8) * MP + 2 * t;
        float *Da_ = DP + g * MP + 2 * t, *Db_ = DP + (g + 8) * MP + 2 * t;
        for (int u0 = 0; u0 < U; u0 += 2) {
            const uint32_t s2a = __ldg(reinterpret_cast<const unsigned short *>(sa + u0));
            const uint32_t s2b = __ldg(reinterpret_cast<const unsigned short *>(sb + u0));
            uint32_t yk[2][NR], yv[2][NR];
            int oct8[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int o = octet(u0 + j);
                oct8[j] = o * 8;
                const int kr = min(o * 8 + g, klast);
                ld_chunk<CH * 2>(yk[j], at(Kb, kr * a.k_sn + cofs));
                ld_chunk<CH * 2>(yv[j], at(Vb, kr * a.v_sn + cofs));
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                int s0 = sbyte(s2a, j), s1 = sbyte(s2b, j);
                if (u0 + j >= U) s0 = s1 = -1;
                float sacc[4] = {0.f, 0.f, 0.f, 0.f}, dacc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int s = 0; s < CH / 4; ++s) {
                    mma16<T>(sacc, xa[2 * s], xb[2 * s], xa[2 * s + 1], xb[2 * s + 1], yk[j][2 * s], yk[j][2 * s + 1]);
                    mma16<T>(dacc, da[2 * s], db[2 * s], da[2 * s + 1], db[2 * s + 1], yv[j][2 * s], yv[j][2 * s + 1]);
                }
                if constexpr (PB) {
                    const float2 k0 = __ldg(PK + min(oct8[j] + 2 * t, klast)), k1 = __ldg(PK + min(oct8[j] + 2 * t + 1, klast));
                    sacc[0] += pos_bias(pw, pqa, k0); sacc[1] += pos_bias(pw, pqa, k1);
                    sacc[2] += pos_bias(pw, pqb, k0); sacc[3] += pos_bias(pw, pqb, k1);
                }
                if (s0 >= 0) {
                    *reinterpret_cast<float2 *>(Sa + 8 * s0) = make_float2(sacc[0], sacc[1]);
                    *reinterpret_cast<float2 *>(Da_ + 8 * s0) = make_float2(dacc[0], dacc[1]);
                }
                if (s1 >= 0) {
                    *reinterpret_cast<float2 *>(Sb + 8 * s1) = make_float2(sacc[2], sacc[3]);
                    *reinterpret_cast<float2 *>(Db_ + 8 * s1) = make_float2(dacc[2], dacc[3]);
                }
            }
        }
    }
    __syncwarp();
    // ---- elementwise: p = exp(s + bias + mask - lse), ds = p (dp - D); two lanes per token row; 16-bit tiles ---------------------
    {
        const int row = lane >> 1, half = lane & 1;
        const int i = i0 + row;
        const bool rvalid = i < Nq && !((impm >> row) & 1u);
        const float *Sr = S + row * MP, *Dr = DP + row * MP;
        T *Pr = Pt + row * M, *dSr = dSt + row * M;
        const int Mh = M >> 1, j0 = half * Mh, j1 = j0 + Mh;
        if (rvalid) {
            const float D = Sr[M], lse = Sr[M + 1];
            const int32_t *bi = PB ? nullptr : a.bias_idx + ((int64_t)b * Nq + i) * M;
            const uint8_t *mk = a.mask ? a.mask + ((int64_t)b * Nq + i) * M : nullptr;
            for (int j = j0; j < j1; j += 4) {
                float4 x = *reinterpret_cast<const float4 *>(Sr + j);
                const float4 d = *reinterpret_cast<const float4 *>(Dr + j);
                if constexpr (!PB) {                                 // (PB: the bias went in with the logits in pass 1)
                    const int4 bv4 = __ldg(reinterpret_cast<const int4 *>(bi + j));
                    x.x += __ldg(a.bias_tab + bv4.x * H + h);
                    x.y += __ldg(a.bias_tab + bv4.y * H + h);
                    x.z += __ldg(a.bias_tab + bv4.z * H + h);
                    x.w += __ldg(a.bias_tab + bv4.w * H + h);
                }
                if (mk) {
                    const uchar4 m4 = *reinterpret_cast<const uchar4 *>(mk + j);
                    if (!m4.x) x.x += -100.f;
                    if (!m4.y) x.y += -100.f;
                    if (!m4.z) x.z += -100.f;
                    if (!m4.w) x.w += -100.f;
                }
                const float p0 = __expf(x.x - lse), p1 = __expf(x.y - lse), p2 = __expf(x.z - lse), p3 = __expf(x.w - lse);
                uint2 pw, dw;
                pw.x = pack_pair<T>(p0, p1); pw.y = pack_pair<T>(p2, p3);
                dw.x = pack_pair<T>(p0 * (d.x - D), p1 * (d.y - D)); dw.y = pack_pair<T>(p2 * (d.z - D), p3 * (d.w - D));
                *reinterpret_cast<uint2 *>(Pr + j) = pw;
                *reinterpret_cast<uint2 *>(dSr + j) = dw;
            }
        } else {
            for (int j = j0; j < j1; j += 4) {
                *reinterpret_cast<uint2 *>(Pr + j) = make_uint2(0u, 0u);
                *reinterpret_cast<uint2 *>(dSr + j) = make_uint2(0u, 0u);
            }
        }
    }
    __syncwarp();
    // ---- P and dS tiles -> global memory (contiguous [rows][M] blocks of the [B,H,Nq,M] tensors), 16-byte stores ------------------
    {
        const int rows = min(TILE_TOK, Nq - i0);
        const int cpr = M >> 3;                          // 16-byte chunks per row
        const int nch = rows * cpr;
        uint4 *gP = reinterpret_cast<uint4 *>(reinterpret_cast<T *>(a.P) + (row0 + i0) * M);
        uint4 *gS = reinterpret_cast<uint4 *>(reinterpret_cast<T *>(a.dS) + (row0 + i0) * M);
        const uint4 *sP = reinterpret_cast<const uint4 *>(Pt), *sS = reinterpret_cast<const uint4 *>(dSt);
        if (impm == 0) {
            for (int c = lane; c < nch; c += 32) { gP[c] = sP[c]; gS[c] = sS[c]; }
        } else {
            for (int c = lane; c < nch; c += 32)
                if (!((impm >> (c / cpr)) & 1u)) { gP[c] = sP[c]; gS[c] = sS[c]; }
        }
    }
    // ---- pass 2: d_q = dS.K over the union octets (K staged key-major, two octets per stage, double buffered) ----------------------
    float acc[NT][4];
#pragma unroll
    for (int n = 0; n < NT; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;
    float gacc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};       // PB: sum of dS * [feat | 1] over this lane's (row, key) pairs
    {
        const int blk = lane % NT, prow = lane / NT;
        const bool act = 8 * blk < C;
        const uint32_t dst_lane = sK + prow * ROWB + blk * 16;
        auto stage = [&](int p, int which) {
            const int o0 = min(octet(2 * p) * 8 + (prow & 7), a.Nk - 1) * a.k_sn + 8 * blk;
            const int o1 = min(octet(2 * p + 1) * 8 + (prow & 7), a.Nk - 1) * a.k_sn + 8 * blk;
            if (act) {
                if constexpr (RPP == 8) {
                    cp16(dst_lane + which * KSTG, at(Kb, o0));
                    cp16(dst_lane + which * KSTG + 8 * ROWB, at(Kb, o1));
                } else {
                    cp16(dst_lane + which * KSTG, at(Kb, prow >= 8 ? o1 : o0));
                }
            }
            cp_commit();
        };
        if (8 * NT > C) {
            for (int x = lane; x < 2 * KSTG / 16; x += 32)
                asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(sK + 16 * x), "r"(0u));
            __syncwarp();
        }
        const int P2 = (U + 1) >> 1;
        const uint32_t wa = (uint32_t)__cvta_generic_to_shared(dSt) + (g * M + 2 * t) * 2, wb = wa + 16 * M;
        const int mi = lane >> 3;
        const uint32_t lrow = sK + (((mi & 1) << 3) + (lane & 7)) * ROWB + (mi >> 1) * 16;
        if (P2 > 0) stage(0, 0);
        for (int p = 0; p < P2; ++p) {
            if (p + 1 < P2) stage(p + 1, (p + 1) & 1);
            const uint32_t s2a = __ldg(reinterpret_cast<const unsigned short *>(sa + 2 * p));
            const uint32_t s2b = __ldg(reinterpret_cast<const unsigned short *>(sb + 2 * p));
            const int s00 = sbyte(s2a, 0), s10 = sbyte(s2b, 0);
            int s01 = sbyte(s2a, 1), s11 = sbyte(s2b, 1);
            if (2 * p + 1 >= U) s01 = s11 = -1;
            uint32_t af[4];
            af[0] = lds32(wa + 16 * max(s00, 0));
            af[1] = lds32(wb + 16 * max(s10, 0));
            af[2] = lds32(wa + 16 * max(s01, 0));
            af[3] = lds32(wb + 16 * max(s11, 0));
            if (s00 < 0) af[0] = 0u;
            if (s10 < 0) af[1] = 0u;
            if (s01 < 0) af[2] = 0u;
            if (s11 < 0) af[3] = 0u;
            if constexpr (PB) {                          // af[2jj] / af[2jj+1] = dS of rows g / g+8 at keys 2t, 2t+1 of octet 2p + jj
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) {
                    const int o8 = octet(2 * p + jj) * 8;
                    const float2 k0 = __ldg(PK + min(o8 + 2 * t, a.Nk - 1)), k1 = __ldg(PK + min(o8 + 2 * t + 1, a.Nk - 1));
                    const float2 dsa = unpack_pair<T>(af[2 * jj]), dsb2 = unpack_pair<T>(af[2 * jj + 1]);
                    pos_bias_grad(gacc, pqa, k0, dsa.x); pos_bias_grad(gacc, pqa, k1, dsa.y);
                    pos_bias_grad(gacc, pqb, k0, dsb2.x); pos_bias_grad(gacc, pqb, k1, dsb2.y);
                }
            }
            if (p + 1 < P2) cp_wait<1>(); else cp_wait<0>();
            __syncwarp();
            const uint32_t yst = lrow + (p & 1) * KSTG;
#pragma unroll
            for (int n = 0; n < NT; n += 2) {
                uint32_t bfr[4];
                ldsm4t(bfr, yst + n * 16);
                mma16<T>(acc[n], af[0], af[1], af[2], af[3], bfr[0], bfr[1]);
                mma16<T>(acc[n + 1], af[0], af[1], af[2], af[3], bfr[2], bfr[3]);
            }
            __syncwarp();
        }
    }
    if constexpr (PB) pos_bias_grad_flush(gacc, a.pe_parts, tile + 131 * b, H, h, lane);
    // ---- epilogue: d_q = acc + ds_blank * blank_k --------------------------------------------------------------------------------
    {
        T *Dq = opaque(reinterpret_cast<T *>(a.dq) + b * a.dq_sb + h * a.dq_sh);
        const T *bk = reinterpret_cast<const T *>(a.blank_k) + h * C;
        const int oa = ra * a.dq_sn + 2 * t, ob = rb * a.dq_sn + 2 * t;
#pragma unroll
        for (int n = 0; n < NT; ++n) {
            const int ch = 8 * n + 2 * t;
            const bool ok = ch < C;
            const float k0 = ok ? to_f(bk[ch]) : 0.f, k1 = ok ? to_f(bk[ch + 1]) : 0.f;
            st_pair_if<T>(at(Dq, oa + 8 * n), acc[n][0] + dsb_a * k0, acc[n][1] + dsb_a * k1, (va && ok) ? 0 : -1);
            st_pair_if<T>(at(Dq, ob + 8 * n), acc[n][2] + dsb_b * k0, acc[n][3] + dsb_b * k1, (vb && ok) ? 0 : -1);
        }
    }
    }
    __syncwarp();
    const SlowIter si = slow_items(pk, a.H);
    for (int it = si.first; it < si.n; it += si.stride) {
        const int gi = pk.imp_list[it / a.H], hh = it % a.H;
        const int bb = gi / a.Nq;
        bwd_row<T, PB>(a, bb, hh, gi - bb * a.Nq, S, lane);
    }
}

// generic path (index tensors without octet structure): one warp per (token, head)
template <typename T, bool PB>
__global__ void __launch_bounds__(256)
attn_bwd_generic_kernel(const ArgsOf<PB> a, const int *__restrict__ tile_flag) {
    extern __shared__ __align__(16) unsigned char dyn_fb[];
    if (tile_flag && tile_flag[0] == 0) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *sc = reinterpret_cast<float *>(dyn_fb) + warp * a.M;
    const int64_t total = (int64_t)a.B * a.Nq * a.H;
    for (int64_t it = (int64_t)blockIdx.x * 8 + warp; it < total; it += (int64_t)gridDim.x * 8) {
        const int h = (int)(it % a.H);
        const int64_t bi = it / a.H;
        bwd_row<T, PB>(a, (int)(bi / a.Nq), h, (int)(bi % a.Nq), sc, lane);
    }
}

}  // namespace fb

template <typename T> int launch_scat_tile2(const T *W, const T *X, const int32_t *csr_off, const uint32_t *csr_ent, const void *pack, T *out,
                                            int B, int H, int Nq, int Nk, int C, int M, Rows4 w, Rows4 x, Rows4 o, cudaStream_t st);

static inline bool fb_fits31(int64_t v) { return v >= 0 && v < (1LL << 31); }

template <typename T, bool PB>
static int launch_attn_bwd(const fb::ArgsOf<PB> &a0, const void *pack, cudaStream_t st) {
    fb::ArgsOf<PB> a = a0;
    const int M = a.M, C = a.C;
    const int *flag = nullptr;
    const int NT = C <= 16 ? 2 : 4;
    const size_t spw = (size_t)2 * 16 * (M + 4) * 4 + (size_t)2 * 16 * M * 2 + 2 * 16 * (NT * 16 + 16);
    const bool shape_ok = C % 8 == 0 && C >= 8 && C <= 32 && M % 8 == 0 && M <= 256;
    const int WPC = spw * 4 <= 100 * 1024 ? 4 : spw * 2 <= 100 * 1024 ? 2 : 1;
    auto al = [](const void *p, int64_t sb, int sh, int sn) { return aligned16(p) && sb % 8 == 0 && sh % 8 == 0 && sn % 8 == 0; };
    const bool align_ok = al(a.q, a.q_sb, a.q_sh, a.q_sn) && al(a.k, a.k_sb, a.k_sh, a.k_sn) && al(a.v, a.v_sb, a.v_sh, a.v_sn) &&
                          al(a.dO, a.do_sb, a.do_sh, a.do_sn) && al(a.O, a.o_sb, a.o_sh, a.o_sn) && al(a.dq, a.dq_sb, a.dq_sh, a.dq_sn) &&
                          aligned16(a.P) && aligned16(a.dS) && (PB || aligned16(a.bias_idx)) && aligned16(a.blank_k) && aligned16(a.blank_v) &&
                          (!a.mask || (reinterpret_cast<uintptr_t>(a.mask) & 3u) == 0);
    if (pack && shape_ok && align_ok && (int64_t)a.B * a.H <= 65535 && spw * WPC <= 200 * 1024) {
        const PackView pk = pack_view(const_cast<void *>(pack), a.B, a.Nq, a.Nk);
        a.smem_per_warp = (int)spw;
        const dim3 grid(ceil_div(pk.T, WPC), a.B * a.H);
        const size_t smem = spw * WPC;
#define FB_LAUNCH(CH_, NT_) do { auto kfn = fb::attn_bwd_tile_kernel<T, CH_, NT_, PB>; \
            if (smem > 48 * 1024) cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            kfn<<<grid, WPC * 32, smem, st>>>(a, pk); } while (0)
        if (C <= 16) FB_LAUNCH(4, 2); else FB_LAUNCH(8, 4);
#undef FB_LAUNCH
        note_launches(1);
        if (int e = check_launch("attn_bwd_tile")) return e;
        flag = reinterpret_cast<const int *>(pack);
    }
    const int64_t total = (int64_t)a.B * a.Nq * a.H;
    int grid = ceil_div(total, 8);
    if (flag && grid > 148 * 8) grid = 148 * 8;
    const size_t smem = (size_t)8 * M * sizeof(float);
    if (smem > 48 * 1024) return set_error(CLUSTEN_EUNSUPPORTED, "fused attention backward: M=%d too large", M);
    fb::attn_bwd_generic_kernel<T, PB><<<grid, 256, smem, st>>>(a, flag);
    note_launches(1);
    return check_launch("attn_bwd_generic");
}

}  // namespace clusten

using namespace clusten;

extern "C" int clusten_attn_bwd(const void *d_out, const void *out, const float *lse, const void *q, const void *k, const void *v,
                                const int64_t *nbhd_idx, const void *pack, const float *bias_tab, const int32_t *bias_idx,
                                const uint8_t *mask, const void *blank_k, const void *blank_v,
                                void *d_q, void *probs, void *d_logits, float *p_blank, float *ds_blank,
                                int B, int H, int Nq, int Nk, int C, int M,
                                int64_t q_sb, int64_t q_sh, int64_t q_sn, int64_t k_sb, int64_t k_sh, int64_t k_sn,
                                int64_t v_sb, int64_t v_sh, int64_t v_sn, int64_t do_sb, int64_t do_sh, int64_t do_sn,
                                int64_t o_sb, int64_t o_sh, int64_t o_sn, int64_t dq_sb, int64_t dq_sh, int64_t dq_sn,
                                int dtype, void *stream) {
    if (B < 0 || H <= 0 || Nq < 0 || Nk <= 0 || C <= 0 || M <= 0)
        return set_error(CLUSTEN_EINVAL, "bad sizes B=%d H=%d Nq=%d Nk=%d C=%d M=%d", B, H, Nq, Nk, C, M);
    if (!d_out || !out || !lse || !q || !k || !v || !nbhd_idx || !bias_tab || !bias_idx || !blank_k || !blank_v || !d_q || !probs ||
        !d_logits || !p_blank || !ds_blank)
        return set_error(CLUSTEN_EINVAL, "null pointer");
    if (dtype != CLUSTEN_F16 && dtype != CLUSTEN_BF16)
        return set_error(CLUSTEN_EUNSUPPORTED, "fused attention backward supports fp16 / bf16 only (dtype %d)", dtype);
    if ((int64_t)B * Nq == 0) return 0;
    const int64_t lim[] = {Nq * q_sn + H * q_sh, (int64_t)Nk * k_sn + H * k_sh, (int64_t)Nk * v_sn + H * v_sh, Nq * do_sn + H * do_sh,
                           Nq * o_sn + H * o_sh, Nq * dq_sn + H * dq_sh, (int64_t)H * Nq * M};
    for (int64_t x : lim)
        if (!fb_fits31(x)) return set_error(CLUSTEN_EUNSUPPORTED, "fused attention backward: per-sample extent exceeds 2^31 elements");
    fb::BwdArgs a{q, k, v, d_out, out, nbhd_idx, bias_tab, bias_idx, mask, blank_k, blank_v, lse, d_q, probs, d_logits, p_blank, ds_blank,
                  B, H, Nq, Nk, C, M, (int)q_sh, (int)q_sn, (int)k_sh, (int)k_sn, (int)v_sh, (int)v_sn, (int)do_sh, (int)do_sn,
                  (int)o_sh, (int)o_sn, (int)dq_sh, (int)dq_sn, q_sb, k_sb, v_sb, do_sb, o_sb, dq_sb, 0};
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == CLUSTEN_F16) return launch_attn_bwd<__half, false>(a, pack, st);
    return launch_attn_bwd<__nv_bfloat16, false>(a, pack, st);
}

// The same backward for the position-bias variant (clusten_attn_pos_fwd): no bias table; the gradient of pos_embed
// comes out as pe_grad_parts [1024][H][6] fp32 partial sums (accumulated INTO with atomics; caller zeroes the buffer and sums
// over the first dimension): [.., h, 0:5] = d_pe_weight[h, :], [.., h, 5] = d_pe_bias[h].
extern "C" int clusten_attn_pos_bwd(const void *d_out, const void *out, const float *lse, const void *q, const void *k, const void *v,
                                    const int64_t *nbhd_idx, const void *pack, const float *pos_q, const float *pos_k,
                                    const float *pe_weight, const float *pe_bias,
                                    const uint8_t *mask, const void *blank_k, const void *blank_v,
                                    void *d_q, void *probs, void *d_logits, float *p_blank, float *ds_blank, float *pe_grad_parts,
                                    int B, int H, int Nq, int Nk, int C, int M,
                                    int64_t q_sb, int64_t q_sh, int64_t q_sn, int64_t k_sb, int64_t k_sh, int64_t k_sn,
                                    int64_t v_sb, int64_t v_sh, int64_t v_sn, int64_t do_sb, int64_t do_sh, int64_t do_sn,
                                    int64_t o_sb, int64_t o_sh, int64_t o_sn, int64_t dq_sb, int64_t dq_sh, int64_t dq_sn,
                                    int dtype, void *stream) {
    if (B < 0 || H <= 0 || Nq < 0 || Nk <= 0 || C <= 0 || M <= 0)
        return set_error(CLUSTEN_EINVAL, "bad sizes B=%d H=%d Nq=%d Nk=%d C=%d M=%d", B, H, Nq, Nk, C, M);
    if (!d_out || !out || !lse || !q || !k || !v || !nbhd_idx || !pos_q || !pos_k || !pe_weight || !blank_k || !blank_v || !d_q ||
        !probs || !d_logits || !p_blank || !ds_blank || !pe_grad_parts)
        return set_error(CLUSTEN_EINVAL, "null pointer");
    if (dtype != CLUSTEN_F16 && dtype != CLUSTEN_BF16)
        return set_error(CLUSTEN_EUNSUPPORTED, "fused attention backward supports fp16 / bf16 only (dtype %d)", dtype);
    if ((reinterpret_cast<uintptr_t>(pos_q) | reinterpret_cast<uintptr_t>(pos_k)) & 7u)
        return set_error(CLUSTEN_EUNSUPPORTED, "positions must be 8-byte aligned");
    if ((int64_t)B * Nq == 0) return 0;
    const int64_t lim[] = {Nq * q_sn + H * q_sh, (int64_t)Nk * k_sn + H * k_sh, (int64_t)Nk * v_sn + H * v_sh, Nq * do_sn + H * do_sh,
                           Nq * o_sn + H * o_sh, Nq * dq_sn + H * dq_sh, (int64_t)H * Nq * M};
    for (int64_t x : lim)
        if (!fb_fits31(x)) return set_error(CLUSTEN_EUNSUPPORTED, "fused attention backward: per-sample extent exceeds 2^31 elements");
    fb::BwdArgsPB a;
    static_cast<fb::BwdArgs &>(a) = fb::BwdArgs{q, k, v, d_out, out, nbhd_idx, nullptr, nullptr, mask, blank_k, blank_v, lse, d_q, probs,
                                                d_logits, p_blank, ds_blank, B, H, Nq, Nk, C, M, (int)q_sh, (int)q_sn, (int)k_sh, (int)k_sn,
                                                (int)v_sh, (int)v_sn, (int)do_sh, (int)do_sn, (int)o_sh, (int)o_sn, (int)dq_sh, (int)dq_sn,
                                                q_sb, k_sb, v_sb, do_sb, o_sb, dq_sb, 0};
    a.pos_q = pos_q; a.pos_k = pos_k; a.pe_w = pe_weight; a.pe_b = pe_bias; a.pe_parts = pe_grad_parts;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == CLUSTEN_F16) return launch_attn_bwd<__half, true>(a, pack, st);
    return launch_attn_bwd<__nv_bfloat16, true>(a, pack, st);
}
