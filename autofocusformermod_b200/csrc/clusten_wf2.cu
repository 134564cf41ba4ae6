// Tensor-core WF / WEIGHTEDGATHER kernels for 16-bit types (sm_100a) + the WF plan (wf2.cuh).
//
//   fwd  out[b,i,ic,c]  = sum_j w[b,i,j,ic] f[b,idx[b,i,j],c]          one warp per token:  D[c][ic]  = F^T[c][j]  W[j][ic]
//   d_w  d_w[b,i,j,ic]  = sum_c f[b,idx[b,i,j],c] d_out[b,i,ic,c]      one warp per token:  D[j][ic]  = F[j][c]    dO^T[c][ic]
//   d_f  d_f[b,8o+r,c]  = sum_{(i,s)->o} sum_ic w[b,i,8s+r,ic] d_out[b,i,ic,c]
//                                                                      one warp per octet:  D[c][r]   = dO^T[c][(e,ic)] W[(e,ic)][r]
// All three are mma.sync.m16n8k16 products whose big operand is read STRAIGHT from global memory into fragments: a
// lane loads LW = 16 / 8 / 4 contiguous bytes of four rows (k = 2t, 2t+1, 2t+8, 2t+9) and two PRMTs per 32-bit word
// transpose them into the (k, k+1) pairs the A fragment wants; the m index of the tile is mapped onto the channels a
// lane holds (m = g -> even channel, m = g + 8 -> odd channel of word w), which is a free permutation because the
// accumulator rows come back to the same lane.  No shared memory, no index structure: any idx works for fwd / d_w.
// Row re-use between neighbouring tokens comes from L1 -- the plan's perm makes the tokens of a CTA spatial neighbours.
#include "t2.cuh"
#include "wf2.cuh"

namespace clusten {
namespace wf2 {

constexpr int WPC = 8;                      // warps per CTA

template <int LW> struct Chunk { uint32_t r[LW / 4]; };
template <int LW> __device__ __forceinline__ void ld_bytes(Chunk<LW> &c, const void *p) {
    if constexpr (LW == 16) { const uint4 v = t2::ldg16(p); c.r[0] = v.x; c.r[1] = v.y; c.r[2] = v.z; c.r[3] = v.w; }
    else if constexpr (LW == 8) { const uint2 v = t2::ldg8(p); c.r[0] = v.x; c.r[1] = v.y; }
    else c.r[0] = t2::ldg4(p);
}
template <int LW> __device__ __forceinline__ void st_bytes(void *p, const uint32_t (&r)[LW / 4]) {
    if constexpr (LW == 16) *reinterpret_cast<uint4 *>(p) = make_uint4(r[0], r[1], r[2], r[3]);
    else if constexpr (LW == 8) *reinterpret_cast<uint2 *>(p) = make_uint2(r[0], r[1]);
    else *reinterpret_cast<uint32_t *>(p) = r[0];
}
__device__ __forceinline__ uint32_t ldu16(const void *p) {
    return (uint32_t)__ldg(reinterpret_cast<const unsigned short *>(p));
}

// ---------------------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------------------
template <typename T, int LW, int NB>
__global__ void __launch_bounds__(WPC * 32)
fwd_kernel(const T *__restrict__ W, const T *__restrict__ F, const int64_t *__restrict__ idx, T *__restrict__ out,
           const int *__restrict__ perm, const int *__restrict__ list, const int *__restrict__ list_cnt,
           int B, int Nq, int C, int M, int IC, int64_t f_sb, int f_sn, int tpw) {
    constexpr int WPL = LW / 4, CB = 4 * LW, EPL = LW / 2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int64_t total = list ? (int64_t)min((int64_t)list_cnt[0], (int64_t)B * Nq) : (int64_t)B * Nq;
    // the warps of a CTA take NEIGHBOURING tokens of the plan's order at the same time: rows they share are fetched from
    // L2 once (L1 hit, or merged with the miss still in flight)
    const int64_t cta_first = (int64_t)blockIdx.x * WPC * tpw + warp;
    for (int it = 0; it < tpw; ++it) {
        const int64_t r = cta_first + (int64_t)it * WPC;
        if (r >= total) break;
        const int b = (int)((list ? (int64_t)list[r] : r) / Nq);
        const int64_t tok = list ? (int64_t)list[r] : perm ? (int64_t)b * Nq + perm[r] : r;
        const int64_t *irow = idx + tok * M;
        const T *wrow = W + tok * M * IC;
        const T *Fb = t2::opaque(F + b * f_sb + g * EPL);
        T *orow = out + tok * IC * C + g * EPL;
        for (int c0 = 0; c0 < C; c0 += NB * CB) {
            float acc[NB][WPL][4];
#pragma unroll
            for (int nb = 0; nb < NB; ++nb)
#pragma unroll
                for (int w = 0; w < WPL; ++w) acc[nb][w][0] = acc[nb][w][1] = acc[nb][w][2] = acc[nb][w][3] = 0.f;
#pragma unroll 1
            for (int k0 = 0; k0 < M; k0 += 16) {
                const int j0 = k0 + 2 * t, j1 = j0 + 1, j2 = j0 + 8, j3 = j0 + 9;
                const int r0 = (int)__ldg(irow + min(j0, M - 1)) * f_sn + c0, r1 = (int)__ldg(irow + min(j1, M - 1)) * f_sn + c0;
                const int r2 = (int)__ldg(irow + min(j2, M - 1)) * f_sn + c0, r3 = (int)__ldg(irow + min(j3, M - 1)) * f_sn + c0;
                // weights of ic = g; lanes g >= IC re-read the last ic: their accumulator columns are never stored
                const int gi = min(g, IC - 1);
                const uint32_t w0 = j0 < M ? ldu16(wrow + j0 * IC + gi) : 0u, w1 = j1 < M ? ldu16(wrow + j1 * IC + gi) : 0u;
                const uint32_t w2 = j2 < M ? ldu16(wrow + j2 * IC + gi) : 0u, w3 = j3 < M ? ldu16(wrow + j3 * IC + gi) : 0u;
                const uint32_t b0 = w0 | (w1 << 16), b1 = w2 | (w3 << 16);
                Chunk<LW> x0[NB], x1[NB], x2[NB], x3[NB];
#pragma unroll
                for (int nb = 0; nb < NB; ++nb) {
                    ld_bytes<LW>(x0[nb], t2::at(Fb, r0 + nb * CB));
                    ld_bytes<LW>(x1[nb], t2::at(Fb, r1 + nb * CB));
                    ld_bytes<LW>(x2[nb], t2::at(Fb, r2 + nb * CB));
                    ld_bytes<LW>(x3[nb], t2::at(Fb, r3 + nb * CB));
                }
#pragma unroll
                for (int nb = 0; nb < NB; ++nb)
#pragma unroll
                    for (int w = 0; w < WPL; ++w)
                        t2::mma16<T>(acc[nb][w], __byte_perm(x0[nb].r[w], x1[nb].r[w], 0x5410), __byte_perm(x0[nb].r[w], x1[nb].r[w], 0x7632),
                                     __byte_perm(x2[nb].r[w], x3[nb].r[w], 0x5410), __byte_perm(x2[nb].r[w], x3[nb].r[w], 0x7632), b0, b1);
            }
            // lane (g, t) holds channels c0 + nb*CB + g*EPL + [0, EPL) of ic = 2t (acc[..][0], [2]) and ic = 2t+1 ([1], [3])
            if (2 * t < IC) {
#pragma unroll
                for (int nb = 0; nb < NB; ++nb) {
                    uint32_t o0[WPL], o1[WPL];
#pragma unroll
                    for (int w = 0; w < WPL; ++w) {
                        o0[w] = t2::pack_pair<T>(acc[nb][w][0], acc[nb][w][2]);
                        o1[w] = t2::pack_pair<T>(acc[nb][w][1], acc[nb][w][3]);
                    }
                    st_bytes<LW>(orow + (int64_t)(2 * t) * C + c0 + nb * CB, o0);
                    if (2 * t + 1 < IC) st_bytes<LW>(orow + (int64_t)(2 * t + 1) * C + c0 + nb * CB, o1);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// d_w : D[j][ic] = sum_c F[idx_j][c] dO[ic][c]; lane loads LW = 16 (8) bytes = 8 (4) channels of rows g / g+8: one (half a)
// 32-channel block = 2 (1) k-steps, the k index permuted identically for both operands.
// ---------------------------------------------------------------------------------------------------------------------
template <typename T, int LW, int MT>
__global__ void __launch_bounds__(WPC * 32)
dw_kernel(const T *__restrict__ dO, const T *__restrict__ F, const int64_t *__restrict__ idx, T *__restrict__ dW,
          const int *__restrict__ perm, int B, int Nq, int C, int M, int IC, int64_t f_sb, int f_sn, int tpw) {
    constexpr int KS = LW / 8, CB = 2 * LW, EPL = LW / 2;       // k-steps / channels per block / channels per lane
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int64_t total = (int64_t)B * Nq;
    const int64_t cta_first = (int64_t)blockIdx.x * WPC * tpw + warp;
    for (int it = 0; it < tpw; ++it) {
        const int64_t r = cta_first + (int64_t)it * WPC;
        if (r >= total) break;
        const int b = (int)(r / Nq);
        const int64_t tok = perm ? (int64_t)b * Nq + perm[r] : r;
        const int64_t *irow = idx + tok * M;
        const T *Fb = t2::opaque(F + b * f_sb + t * EPL);
        // B operand: d_out row of ic = g; lanes g >= IC re-read the last ic -- their accumulator columns are never stored
        const T *drow = t2::opaque(dO + (tok * IC + min(g, IC - 1)) * C + t * EPL);
        T *wrow = dW + tok * M * IC;
        for (int m0 = 0; m0 < M; m0 += 16 * MT) {               // MT m-tiles of 16 neighbours at a time: 2*MT independent row loads
            int ra[MT], rb[MT];
#pragma unroll
            for (int x = 0; x < MT; ++x) {
                ra[x] = (int)__ldg(irow + min(m0 + 16 * x + g, M - 1)) * f_sn;
                rb[x] = (int)__ldg(irow + min(m0 + 16 * x + g + 8, M - 1)) * f_sn;
            }
            float acc[MT][4];
#pragma unroll
            for (int x = 0; x < MT; ++x) acc[x][0] = acc[x][1] = acc[x][2] = acc[x][3] = 0.f;
#pragma unroll 2
            for (int c0 = 0; c0 < C; c0 += CB) {
                Chunk<LW> d, xa[MT], xb[MT];
                ld_bytes<LW>(d, t2::at(drow, c0));
#pragma unroll
                for (int x = 0; x < MT; ++x) {
                    ld_bytes<LW>(xa[x], t2::at(Fb, ra[x] + c0));
                    ld_bytes<LW>(xb[x], t2::at(Fb, rb[x] + c0));
                }
#pragma unroll
                for (int x = 0; x < MT; ++x)
#pragma unroll
                    for (int s = 0; s < KS; ++s)
                        t2::mma16<T>(acc[x], xa[x].r[2 * s], xb[x].r[2 * s], xa[x].r[2 * s + 1], xb[x].r[2 * s + 1], d.r[2 * s], d.r[2 * s + 1]);
            }
            if (2 * t < IC) {
#pragma unroll
                for (int x = 0; x < MT; ++x) {
                    const int ja = m0 + 16 * x + g, jb = ja + 8;
                    if (IC >= 2) {
                        if (ja < M) *reinterpret_cast<uint32_t *>(wrow + ja * IC + 2 * t) = t2::pack_pair<T>(acc[x][0], acc[x][1]);
                        if (jb < M) *reinterpret_cast<uint32_t *>(wrow + jb * IC + 2 * t) = t2::pack_pair<T>(acc[x][2], acc[x][3]);
                    } else {
                        if (ja < M) wrow[ja] = from_f<T>(acc[x][0]);
                        if (jb < M) wrow[jb] = from_f<T>(acc[x][2]);
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// d_f, octet form (IC == 4, M % 8 == 0): one warp per feature octet; k = (entry, ic), 4 entries per k-step.
// ---------------------------------------------------------------------------------------------------------------------
template <typename T, int LW, int NB>
__global__ void __launch_bounds__(WPC * 32)
df_oct_kernel(const T *__restrict__ dO, const T *__restrict__ W, const WfPlanView pv, T *__restrict__ dF,
              int B, int Nq, int Nk, int C, int M, int64_t df_sb, int df_sn) {
    constexpr int WPL = LW / 4, CB = 4 * LW, EPL = LW / 2;
    if (pv.flags[0]) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int64_t oid = (int64_t)blockIdx.x * WPC + warp;
    if (oid >= (int64_t)B * pv.NO) return;
    const int b = (int)(oid / pv.NO), o = (int)(oid - (int64_t)b * pv.NO);
    const int lo = pv.oct_off[(int64_t)b * (pv.NO + 1) + o], hi = pv.oct_off[(int64_t)b * (pv.NO + 1) + o + 1];
    const uint32_t *ent = pv.oct_ent + (int64_t)b * Nq * pv.S;
    const T *dOb = t2::opaque(dO + (int64_t)b * Nq * 4 * C + g * EPL);
    const T *Wb = t2::opaque(W + (int64_t)b * Nq * M * 4);
    const int ic = 2 * (t & 1);
    const int ra = 8 * o + 2 * t, rb = ra + 1;
    for (int c0 = 0; c0 < C; c0 += NB * CB) {
        float acc[NB][WPL][4];
#pragma unroll
        for (int nb = 0; nb < NB; ++nb)
#pragma unroll
            for (int w = 0; w < WPL; ++w) acc[nb][w][0] = acc[nb][w][1] = acc[nb][w][2] = acc[nb][w][3] = 0.f;
#pragma unroll 1
        for (int e0 = lo; e0 < hi; e0 += 4) {
            const int ea = e0 + (t >> 1), eb = ea + 2;
            const uint32_t va = __ldg(ent + min(ea, hi - 1)), vb = __ldg(ent + min(eb, hi - 1));
            const int ia = (int)(va >> 5), ib = (int)(vb >> 5);
            const uint32_t b0 = ea < hi ? t2::ldg4(t2::at(Wb, ((ia * M + 8 * (int)(va & 31u) + g) << 2) + ic)) : 0u;
            const uint32_t b1 = eb < hi ? t2::ldg4(t2::at(Wb, ((ib * M + 8 * (int)(vb & 31u) + g) << 2) + ic)) : 0u;
            const int pa = (ia * 4 + ic) * C + c0, pb = (ib * 4 + ic) * C + c0;
            Chunk<LW> x0[NB], x1[NB], x2[NB], x3[NB];
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) {
                ld_bytes<LW>(x0[nb], t2::at(dOb, pa + nb * CB));
                ld_bytes<LW>(x1[nb], t2::at(dOb, pa + C + nb * CB));
                ld_bytes<LW>(x2[nb], t2::at(dOb, pb + nb * CB));
                ld_bytes<LW>(x3[nb], t2::at(dOb, pb + C + nb * CB));
            }
#pragma unroll
            for (int nb = 0; nb < NB; ++nb)
#pragma unroll
                for (int w = 0; w < WPL; ++w)
                    t2::mma16<T>(acc[nb][w], __byte_perm(x0[nb].r[w], x1[nb].r[w], 0x5410), __byte_perm(x0[nb].r[w], x1[nb].r[w], 0x7632),
                                 __byte_perm(x2[nb].r[w], x3[nb].r[w], 0x5410), __byte_perm(x2[nb].r[w], x3[nb].r[w], 0x7632), b0, b1);
        }
        T *da = dF + b * df_sb + (int64_t)ra * df_sn + g * EPL + c0;
        T *db = dF + b * df_sb + (int64_t)rb * df_sn + g * EPL + c0;
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) {
            uint32_t o0[WPL], o1[WPL];
#pragma unroll
            for (int w = 0; w < WPL; ++w) {
                o0[w] = t2::pack_pair<T>(acc[nb][w][0], acc[nb][w][2]);
                o1[w] = t2::pack_pair<T>(acc[nb][w][1], acc[nb][w][3]);
            }
            if (ra < Nk) st_bytes<LW>(da + nb * CB, o0);
            if (rb < Nk) st_bytes<LW>(db + nb * CB, o1);
        }
    }
}

// fp32 form of the same scatter (the merge under AMP runs fp32 weights: the reference's CLUSTENWF casts feat up, clusten.py:80-81).
// One warp per feature octet walks its sorted reference list; per entry (token i, slot s) the 8 x 4 weights w[i, 8s..8s+7, 0..3] are
// ONE coalesced 128-byte load (lane = 4 r + ic, broadcast by shuffles) and the four d_out rows of the token are read once for all 8
// feature rows -- the inverse-list kernel of clusten_wf.cu reads them once per (row, entry), 8x the L2 traffic.  Lane owns channels
// c0 + lane + 32 k, k < KC; plain FFMA (exact fp32), entries in ascending order (deterministic).
template <int KC>
__global__ void __launch_bounds__(WPC * 32)
df_oct32_kernel(const float *__restrict__ dO, const float *__restrict__ W, const WfPlanView pv, float *__restrict__ dF,
                int B, int Nq, int Nk, int C, int M, int64_t df_sb, int64_t df_sn) {
    if (pv.flags[0]) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t oid = (int64_t)blockIdx.x * WPC + warp;
    if (oid >= (int64_t)B * pv.NO) return;
    const int b = (int)(oid / pv.NO), o = (int)(oid - (int64_t)b * pv.NO);
    const int lo = pv.oct_off[(int64_t)b * (pv.NO + 1) + o], hi = pv.oct_off[(int64_t)b * (pv.NO + 1) + o + 1];
    const uint32_t *ent = pv.oct_ent + (int64_t)b * Nq * pv.S;
    const float *dOb = dO + (int64_t)b * Nq * 4 * C;
    const float *Wb = W + (int64_t)b * Nq * M * 4;
    for (int c0 = 0; c0 < C; c0 += 32 * KC) {
        float acc[8][KC];
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int k = 0; k < KC; ++k) acc[r][k] = 0.f;
#pragma unroll 2
        for (int e = lo; e < hi; ++e) {                      // (two entries per trip: the second one's loads overlap the first one's FMAs)
            const uint32_t v = __ldg(ent + e);
            const int64_t i = v >> 5;
            const float wl = __ldg(Wb + (i * M + 8 * (v & 31u)) * 4 + lane);           // w[i, 8s + lane/4, lane%4]
            float d[4][KC];
#pragma unroll
            for (int ic = 0; ic < 4; ++ic)
#pragma unroll
                for (int k = 0; k < KC; ++k) d[ic][k] = __ldg(dOb + (i * 4 + ic) * C + c0 + lane + 32 * k);
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int ic = 0; ic < 4; ++ic) {
                    const float w = __shfl_sync(FULL, wl, 4 * r + ic);
#pragma unroll
                    for (int k = 0; k < KC; ++k) acc[r][k] = fmaf(w, d[ic][k], acc[r][k]);
                }
        }
#pragma unroll
        for (int r = 0; r < 8; ++r)
            if (8 * o + r < Nk) {
#pragma unroll
                for (int k = 0; k < KC; ++k) dF[b * df_sb + (int64_t)(8 * o + r) * df_sn + c0 + lane + 32 * k] = acc[r][k];
            }
    }
}

// flagged rows (referenced by an impure slot): recomputed whole in fp32, one warp per row
template <typename T>
__global__ void __launch_bounds__(WPC * 32)
df_fix_kernel(const T *__restrict__ dO, const T *__restrict__ W, const int64_t *__restrict__ idx, const WfPlanView pv,
              T *__restrict__ dF, int Nq, int Nk, int C, int M, int64_t df_sb, int64_t df_sn) {
    if (pv.flags[0]) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y;
    const int nfr = min(pv.frow_cnt[b], WFP_IMP_CAP * 8);
    for (int fr = blockIdx.x * WPC + warp; fr < nfr; fr += gridDim.x * WPC) {
    const int r = pv.frow_list[(int64_t)b * WFP_IMP_CAP * 8 + fr];
    const int o = r >> 3, jr = r & 7;
    const int lo = pv.oct_off[(int64_t)b * (pv.NO + 1) + o], hi = pv.oct_off[(int64_t)b * (pv.NO + 1) + o + 1];
    const uint32_t *ent = pv.oct_ent + (int64_t)b * Nq * pv.S;
    const uint32_t *imp = pv.imp_list + (int64_t)b * WFP_IMP_CAP;
    const int nimp = min(pv.imp_cnt[b], WFP_IMP_CAP);
    const T *dOb = dO + (int64_t)b * Nq * 4 * C;
    const T *Wb = W + (int64_t)b * Nq * M * 4;
    const int64_t *ib = idx + (int64_t)b * Nq * M;
    for (int c = lane; c < C; c += 32) {
        float acc = 0.f;
        for (int e = lo; e < hi; ++e) {
            const uint32_t v = ent[e];
            const int64_t i = v >> 5;
            const T *wp = Wb + (i * M + 8 * (v & 31u) + jr) * 4;
#pragma unroll
            for (int ic = 0; ic < 4; ++ic) acc = fmaf(to_f(wp[ic]), to_f(dOb[(i * 4 + ic) * C + c]), acc);
        }
        for (int q = 0; q < nimp; ++q) {
            const uint32_t v = imp[q];
            const int64_t i = v >> 5;
            const int s = (int)(v & 31u);
            for (int j = 0; j < 8; ++j) {
                if (ib[i * M + 8 * s + j] != (int64_t)r) continue;
                const T *wp = Wb + (i * M + 8 * s + j) * 4;
#pragma unroll
                for (int ic = 0; ic < 4; ++ic) acc = fmaf(to_f(wp[ic]), to_f(dOb[(i * 4 + ic) * C + c]), acc);
            }
        }
        dF[b * df_sb + (int64_t)r * df_sn + c] = from_f<T>(acc);
    }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// plan build: count (one thread per slot) -> scan (one CTA per sample) -> fill (one thread per slot) -> sort (one warp
// per octet list).  idx is read exactly once; the per-slot octet ids are kept for the fill pass.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 1024;

// exclusive scan of h[0..n) in place (all threads of the CTA call it); returns the total
__device__ int block_excl_scan(int *h, int n, int *wsum) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const int per = (n + nt - 1) / nt;
    const int beg = min(tid * per, n), end = min(beg + per, n);
    int s = 0;
    for (int x = beg; x < end; ++x) s += h[x];
    int incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(FULL, incl, o);
        if ((tid & 31) >= o) incl += v;
    }
    if ((tid & 31) == 31) wsum[tid >> 5] = incl;
    __syncthreads();
    if (tid < 32) {
        const int nw = nt >> 5;
        const int v = tid < nw ? wsum[tid] : 0;
        int inc2 = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(FULL, inc2, o);
            if (tid >= o) inc2 += u;
        }
        wsum[tid] = inc2 - v;
        if (tid == 31) wsum[32] = inc2;
    }
    __syncthreads();
    int run = wsum[tid >> 5] + incl - s;
    for (int x = beg; x < end; ++x) { const int v = h[x]; h[x] = run; run += v; }
    __syncthreads();
    return wsum[32];
}

__global__ void __launch_bounds__(256)
plan_count_kernel(const int64_t *__restrict__ idx, int B, int Nq, int M, int Nk, WfPlanView pv) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int S = pv.S, NO = pv.NO;
    if (p >= (int64_t)B * Nq * S) return;
    const int64_t tok = p / S;
    const int s = (int)(p - tok * S);
    const int b = (int)(tok / Nq), i = (int)(tok - (int64_t)b * Nq);
    const int64_t *ip = idx + tok * M + 8 * s;
    const longlong2 *q = reinterpret_cast<const longlong2 *>(ip);
    const longlong2 a = __ldg(q), c = __ldg(q + 1), d = __ldg(q + 2), e = __ldg(q + 3);
    const int64_t base = a.x;
    const bool pure = base >= 0 && (base & 7) == 0 && base + 7 < (int64_t)Nk && a.y == base + 1 && c.x == base + 2 &&
                      c.y == base + 3 && d.x == base + 4 && d.y == base + 5 && e.x == base + 6 && e.y == base + 7;
    const int o = pure ? (int)(base >> 3) : -1;
    pv.slot_oct[p] = o;
    if (s == 0) {
        const int key = (int)min(max(base >> 3, (int64_t)0), (int64_t)NO - 1);
        pv.tok_key[tok] = key;
        atomicAdd(pv.hist_tok + (int64_t)b * NO + key, 1);
    }
    if (pure) { atomicAdd(pv.hist_oct + (int64_t)b * NO + o, 1); return; }
    const int pos = atomicAdd(pv.imp_cnt + b, 1);
    if (pos >= WFP_IMP_CAP) return;
    pv.imp_list[(int64_t)b * WFP_IMP_CAP + pos] = ((uint32_t)i << 5) | (uint32_t)s;
    const int64_t rows[8] = {a.x, a.y, c.x, c.y, d.x, d.y, e.x, e.y};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int64_t v = rows[j];
        if (v < 0 || v >= (int64_t)Nk) continue;
        const int64_t gb = (int64_t)b * Nk + v;             // byte-wide test-and-set through the aligned word of the flag
        unsigned *wp = reinterpret_cast<unsigned *>(pv.row_flag + (gb & ~(int64_t)3));
        const unsigned bit = 1u << (8 * (int)(gb & 3));
        if (!(atomicOr(wp, bit) & bit)) {
            const int fp = atomicAdd(pv.frow_cnt + b, 1);
            if (fp < WFP_IMP_CAP * 8) pv.frow_list[(int64_t)b * WFP_IMP_CAP * 8 + fp] = (int)v;
        }
    }
}

__global__ void __launch_bounds__(SCAN_THREADS)
plan_scan_kernel(WfPlanView pv) {
    extern __shared__ int hs[];                         // [NO] + [40]
    const int NO = pv.NO, tid = threadIdx.x, nt = blockDim.x, b = blockIdx.x;
    int *wsum = hs + NO;
    int *ho = pv.hist_oct + (int64_t)b * NO, *ht = pv.hist_tok + (int64_t)b * NO;
    int *off = pv.oct_off + (int64_t)b * (NO + 1);
    for (int x = tid; x < NO; x += nt) hs[x] = ho[x];
    __syncthreads();
    const int total = block_excl_scan(hs, NO, wsum);
    for (int x = tid; x < NO; x += nt) { off[x] = hs[x]; ho[x] = hs[x]; }       // ho becomes the fill cursor
    if (tid == 0) off[NO] = total;
    __syncthreads();
    for (int x = tid; x < NO; x += nt) hs[x] = ht[x];
    __syncthreads();
    block_excl_scan(hs, NO, wsum);
    for (int x = tid; x < NO; x += nt) ht[x] = hs[x];
    // impure slots ascending (fixes the summation order of the fix-up kernel)
    const int raw = pv.imp_cnt[b];
    const int n = min(raw, WFP_IMP_CAP);
    uint32_t *imp = pv.imp_list + (int64_t)b * WFP_IMP_CAP;
    uint32_t v = 0;
    int rk = 0;
    if (tid < n) {
        v = imp[tid];
        for (int k = 0; k < n; ++k) rk += imp[k] < v;
    }
    __syncthreads();
    if (tid < n) imp[rk] = v;
    if (tid == 0) {
        atomicAdd(pv.flags + 1, raw);
        if (raw > WFP_IMP_CAP) { atomicOr(pv.flags + 0, 1); atomicOr(pv.flags + 4, 1); }
    }
}

__global__ void __launch_bounds__(256)
plan_fill_kernel(int B, int Nq, WfPlanView pv) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int S = pv.S, NO = pv.NO;
    if (p >= (int64_t)B * Nq * S) return;
    const int64_t tok = p / S;
    const int s = (int)(p - tok * S);
    const int b = (int)(tok / Nq), i = (int)(tok - (int64_t)b * Nq);
    const int o = pv.slot_oct[p];
    if (o >= 0) pv.oct_ent[(int64_t)b * Nq * S + atomicAdd(pv.hist_oct + (int64_t)b * NO + o, 1)] = ((uint32_t)i << 5) | (uint32_t)s;
    if (s == 0) pv.perm[(int64_t)b * Nq + atomicAdd(pv.hist_tok + (int64_t)b * NO + pv.tok_key[tok], 1)] = i;
}

// every list ascending (fixes the summation order): rank sort in registers, <= 4 entries per lane
__global__ void __launch_bounds__(256)
plan_sort_kernel(int B, int Nq, WfPlanView pv) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t oid = (int64_t)blockIdx.x * 8 + warp;
    if (oid >= (int64_t)B * pv.NO) return;
    const int b = (int)(oid / pv.NO), o = (int)(oid - (int64_t)b * pv.NO);
    const int lo = pv.oct_off[(int64_t)b * (pv.NO + 1) + o], L = pv.oct_off[(int64_t)b * (pv.NO + 1) + o + 1] - lo;
    if (L <= 1) return;
    if (lane == 0 && L > 16) atomicMax(pv.flags + 2, L);
    if (L > WFP_LIST_MAX) { if (lane == 0) { atomicOr(pv.flags + 0, 1); atomicOr(pv.flags + 4, 1); } return; }
    uint32_t *ent = pv.oct_ent + (int64_t)b * Nq * pv.S + lo;
    if (L <= 32) {
        const uint32_t v = lane < L ? ent[lane] : 0xffffffffu;
        int rk = 0;
        for (int k = 0; k < L; ++k) rk += __shfl_sync(FULL, v, k) < v;
        if (lane < L) ent[rk] = v;
        return;
    }
    uint32_t v[4];
    int rk[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) { v[q] = lane + 32 * q < L ? ent[lane + 32 * q] : 0xffffffffu; rk[q] = 0; }
    for (int k = 0; k < L; ++k) {
        const uint32_t u = ent[k];
#pragma unroll
        for (int q = 0; q < 4; ++q) rk[q] += u < v[q];
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 4; ++q)
        if (lane + 32 * q < L) ent[rk[q]] = v[q];
}

// tiles of 16 consecutive tokens of perm: union of their octets + slot table (tile.cuh's structure, for a permuted token order)
constexpr int PT_WARPS = 4;
__global__ void __launch_bounds__(PT_WARPS * 32)
plan_tile_kernel(int B, int Nq, WfPlanView pv) {
    __shared__ int oct_rs[PT_WARPS][16][32];
    __shared__ __align__(16) int8_t slot_s[PT_WARPS][WF_UMAX][16];
    __shared__ int row_bad[PT_WARPS][16];
    __shared__ int row_tok[PT_WARPS][16];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t tb = (int64_t)blockIdx.x * PT_WARPS + warp;
    if (tb >= (int64_t)B * pv.T) return;
    const int b = (int)(tb / pv.T), tile = (int)(tb - (int64_t)b * pv.T);
    const int p0 = tile * 16, nrow = min(16, Nq - p0), S = pv.S;
    if (lane < 16) {
        row_tok[warp][lane] = lane < nrow ? pv.perm[(int64_t)b * Nq + p0 + lane] : -1;
        row_bad[warp][lane] = 0;
    }
    for (int x = lane; x < WF_UMAX * 16 / 4; x += 32) reinterpret_cast<int *>(&slot_s[warp][0][0])[x] = -1;
    __syncwarp();
    for (int item = lane; item < 16 * S; item += 32) {
        const int r = item / S, s = item - r * S;
        int o = -2;
        if (r < nrow) {
            o = pv.slot_oct[((int64_t)b * Nq + row_tok[warp][r]) * S + s];
            if (o < 0) row_bad[warp][r] = 1;               // benign race: all writers store 1
        }
        oct_rs[warp][r][s] = o;
    }
    __syncwarp();
    int my0 = -1, my1 = -1, U = 0;                        // lane u holds union position u / u + 32
    for (int r = 0; r < nrow; ++r) {
        if (row_bad[warp][r]) continue;
        for (int s = 0; s < S; ++s) {
            const int o = oct_rs[warp][r][s];
            const unsigned m0 = __ballot_sync(FULL, my0 == o), m1 = __ballot_sync(FULL, my1 == o);
            int pos;
            if (m0) pos = __ffs(m0) - 1;
            else if (m1) pos = 32 + __ffs(m1) - 1;
            else {
                pos = U;
                if (U < 32) { if (lane == U) my0 = o; }
                else if (U < 64) { if (lane == U - 32) my1 = o; }
                ++U;
            }
            if (pos < WF_UMAX) {
                const int prev = slot_s[warp][pos][r];
                __syncwarp();
                if (lane == 0) {
                    if (prev != -1) row_bad[warp][r] = 1;  // one octet twice in a neighbourhood: token by token
                    else slot_s[warp][pos][r] = (int8_t)s;
                }
                __syncwarp();
            }
        }
    }
    __syncwarp();
    const bool over = U > WF_UMAX;
    for (int r = 0; r < nrow; ++r) {
        if (!(over || row_bad[warp][r])) continue;
        for (int u = lane; u < WF_UMAX; u += 32) slot_s[warp][u][r] = -1;
        if (lane == 0) pv.timp_list[atomicAdd(pv.flags + 3, 1)] = b * Nq + row_tok[warp][r];
    }
    __syncwarp();
    const int Uc = over ? 0 : U;
    pv.tile_oct[tb * WF_UMAX + lane] = lane < Uc ? my0 : 0;
    pv.tile_oct[tb * WF_UMAX + 32 + lane] = lane + 32 < Uc ? my1 : 0;
    const int4 *src = reinterpret_cast<const int4 *>(&slot_s[warp][0][0]);
    int4 *dst = reinterpret_cast<int4 *>(pv.slot_t + tb * WF_UMAX * 16);
    for (int x = lane; x < WF_UMAX; x += 32) dst[x] = src[x];
    if (lane == 0) {
        pv.tile_u[tb] = Uc;
        atomicMax(pv.flags + 5, U);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
static inline int pick_nb(int nblk) { return nblk % 4 == 0 ? 4 : nblk % 3 == 0 ? 3 : nblk % 2 == 0 ? 2 : 1; }
static inline int pick_lw(int C) { return C % 64 == 0 ? 16 : C % 32 == 0 ? 8 : C % 16 == 0 ? 4 : 0; }
static inline bool fits31(int64_t v) { return v >= 0 && v < (1LL << 31); }

#define WF2_DISPATCH(LW_, NB_, ...)                                                                     \
    switch (LW_ * 8 + NB_) {                                                                            \
        case 16 * 8 + 1: { constexpr int LW = 16, NB = 1; __VA_ARGS__; break; }                         \
        case 16 * 8 + 2: { constexpr int LW = 16, NB = 2; __VA_ARGS__; break; }                         \
        case 16 * 8 + 3: { constexpr int LW = 16, NB = 3; __VA_ARGS__; break; }                         \
        case 16 * 8 + 4: { constexpr int LW = 16, NB = 4; __VA_ARGS__; break; }                         \
        case 8 * 8 + 1: { constexpr int LW = 8, NB = 1; __VA_ARGS__; break; }                           \
        case 8 * 8 + 2: { constexpr int LW = 8, NB = 2; __VA_ARGS__; break; }                           \
        case 8 * 8 + 3: { constexpr int LW = 8, NB = 3; __VA_ARGS__; break; }                           \
        case 8 * 8 + 4: { constexpr int LW = 8, NB = 4; __VA_ARGS__; break; }                           \
        case 4 * 8 + 1: { constexpr int LW = 4, NB = 1; __VA_ARGS__; break; }                           \
        case 4 * 8 + 2: { constexpr int LW = 4, NB = 2; __VA_ARGS__; break; }                           \
        case 4 * 8 + 3: { constexpr int LW = 4, NB = 3; __VA_ARGS__; break; }                           \
        default: { constexpr int LW = 4, NB = 4; __VA_ARGS__; break; }                                  \
    }

static inline int tokens_per_warp(int64_t tokens) {
    // enough CTAs to fill the machine several times over, while a warp still walks a run of neighbouring tokens
    int tpw = 8;
    while (tpw > 1 && tokens / (tpw * WPC) < 148 * 4) tpw >>= 1;
    return tpw;
}

template <typename T>
static int fwd_t(const T *w, const T *f, const int64_t *idx, T *out, const void *plan, int B, int Nq, int Nk, int C, int M, int IC,
                 int64_t f_sb, int64_t f_sn, cudaStream_t st, bool listed = false) {
    const int LW_ = pick_lw(C);
    if (!LW_ || IC > 8 || (IC > 1 && (IC & 1))) return 0;
    if ((reinterpret_cast<uintptr_t>(f) | reinterpret_cast<uintptr_t>(out)) % LW_ || f_sb % (LW_ / 2) || f_sn % (LW_ / 2)) return 0;
    if (!fits31((int64_t)Nk * f_sn + C) || !fits31((int64_t)M * IC)) return 0;
    const int NB_ = pick_nb(C / (4 * LW_));
    const int *perm = plan ? wf_plan_view(const_cast<void *>(plan), B, Nq, M, Nk).perm : nullptr;
    const int64_t tokens = (int64_t)B * Nq;
    int tpw = tokens_per_warp(tokens);
    int grid = ceil_div(tokens, (int64_t)tpw * WPC);
    const int *list = nullptr, *list_cnt = nullptr;
    if (listed) {                                        // only the tokens of plan.timp_list (count on the device): short grid
        if (!plan) return 0;
        const WfPlanView pv = wf_plan_view(const_cast<void *>(plan), B, Nq, M, Nk);
        list = pv.timp_list;
        list_cnt = pv.flags + 3;
        grid = (int)std::min<int64_t>(grid, 148 * 2);
        tpw = ceil_div(tokens, (int64_t)grid * WPC);
    }
    WF2_DISPATCH(LW_, NB_, (fwd_kernel<T, LW, NB><<<grid, WPC * 32, 0, st>>>(w, f, idx, out, perm, list, list_cnt, B, Nq, C, M, IC, f_sb, (int)f_sn, tpw)));
    note_launches(1);
    return 1;
}

template <typename T>
static int dw_t(const T *d_out, const T *f, const int64_t *idx, T *d_w, const void *plan, int B, int Nq, int Nk, int C, int M, int IC,
                int64_t f_sb, int64_t f_sn, cudaStream_t st) {
    const int LW_ = C % 32 == 0 ? 16 : C % 16 == 0 ? 8 : 0;
    if (!LW_ || IC > 8 || (IC > 1 && (IC & 1))) return 0;
    if ((reinterpret_cast<uintptr_t>(f) | reinterpret_cast<uintptr_t>(d_out)) % LW_ || f_sb % (LW_ / 2) || f_sn % (LW_ / 2)) return 0;
    if (reinterpret_cast<uintptr_t>(d_w) % 4 || !fits31((int64_t)Nk * f_sn + C) || !fits31((int64_t)M * IC)) return 0;
    const int *perm = plan ? wf_plan_view(const_cast<void *>(plan), B, Nq, M, Nk).perm : nullptr;
    const int64_t tokens = (int64_t)B * Nq;
    const int tpw = tokens_per_warp(tokens);
    const int grid = ceil_div(tokens, (int64_t)tpw * WPC);
    if (M <= 16) {
        if (LW_ == 16) dw_kernel<T, 16, 1><<<grid, WPC * 32, 0, st>>>(d_out, f, idx, d_w, perm, B, Nq, C, M, IC, f_sb, (int)f_sn, tpw);
        else dw_kernel<T, 8, 1><<<grid, WPC * 32, 0, st>>>(d_out, f, idx, d_w, perm, B, Nq, C, M, IC, f_sb, (int)f_sn, tpw);
    } else {
        if (LW_ == 16) dw_kernel<T, 16, 3><<<grid, WPC * 32, 0, st>>>(d_out, f, idx, d_w, perm, B, Nq, C, M, IC, f_sb, (int)f_sn, tpw);
        else dw_kernel<T, 8, 3><<<grid, WPC * 32, 0, st>>>(d_out, f, idx, d_w, perm, B, Nq, C, M, IC, f_sb, (int)f_sn, tpw);
    }
    note_launches(1);
    return 1;
}

template <typename T>
static int df_t(const T *d_out, const T *w, const int64_t *idx, T *d_f, const void *plan, int B, int Nq, int Nk, int C, int M, int IC,
                int64_t df_sb, int64_t df_sn, cudaStream_t st) {
    const int LW_ = pick_lw(C);
    if (!plan || !LW_ || IC != 4 || M % 8 || M > 256) return 0;
    if ((reinterpret_cast<uintptr_t>(d_f) | reinterpret_cast<uintptr_t>(d_out)) % LW_ || df_sb % (LW_ / 2) || df_sn % (LW_ / 2)) return 0;
    if (reinterpret_cast<uintptr_t>(w) % 4 || !fits31((int64_t)Nq * M * 4) || !fits31((int64_t)Nq * 4 * C) || Nq >= (1 << 27)) return 0;
    const WfPlanView pv = wf_plan_view(const_cast<void *>(plan), B, Nq, M, Nk);
    const int NB_ = pick_nb(C / (4 * LW_));
    const int grid = ceil_div((int64_t)B * pv.NO, WPC);
    WF2_DISPATCH(LW_, NB_, (df_oct_kernel<T, LW, NB><<<grid, WPC * 32, 0, st>>>(d_out, w, pv, d_f, B, Nq, Nk, C, M, df_sb, (int)df_sn)));
    df_fix_kernel<T><<<dim3(4, B), WPC * 32, 0, st>>>(d_out, w, idx, pv, d_f, Nq, Nk, C, M, df_sb, df_sn);
    note_launches(2);
    return 1;
}

static int df32_t(const float *d_out, const float *w, const int64_t *idx, float *d_f, const void *plan, int B, int Nq, int Nk, int C, int M,
                  int IC, int64_t df_sb, int64_t df_sn, cudaStream_t st) {
    if (!plan || IC != 4 || M % 8 || M > 256 || C % 32 || Nq >= (1 << 27)) return 0;
    const WfPlanView pv = wf_plan_view(const_cast<void *>(plan), B, Nq, M, Nk);
    const int blocks = C / 32;
    const int grid = ceil_div((int64_t)B * pv.NO, WPC);
    if (blocks % 4 == 0) df_oct32_kernel<4><<<grid, WPC * 32, 0, st>>>(d_out, w, pv, d_f, B, Nq, Nk, C, M, df_sb, df_sn);
    else if (blocks % 3 == 0) df_oct32_kernel<3><<<grid, WPC * 32, 0, st>>>(d_out, w, pv, d_f, B, Nq, Nk, C, M, df_sb, df_sn);
    else if (blocks % 2 == 0) df_oct32_kernel<2><<<grid, WPC * 32, 0, st>>>(d_out, w, pv, d_f, B, Nq, Nk, C, M, df_sb, df_sn);
    else df_oct32_kernel<1><<<grid, WPC * 32, 0, st>>>(d_out, w, pv, d_f, B, Nq, Nk, C, M, df_sb, df_sn);
    df_fix_kernel<float><<<dim3(4, B), WPC * 32, 0, st>>>(d_out, w, idx, pv, d_f, Nq, Nk, C, M, df_sb, df_sn);
    note_launches(2);
    return 1;
}

}  // namespace wf2

int wf2_fwd(const void *w, const void *f, const int64_t *idx, void *out, const void *plan, int B, int Nq, int Nk, int C, int M,
            int IC, int64_t f_sb, int64_t f_sn, int dtype, cudaStream_t st) {
    if (dtype == CLUSTEN_BF16)
        return wf2::fwd_t<__nv_bfloat16>((const __nv_bfloat16 *)w, (const __nv_bfloat16 *)f, idx, (__nv_bfloat16 *)out, plan, B, Nq, Nk, C, M, IC, f_sb, f_sn, st);
    if (dtype == CLUSTEN_F16)
        return wf2::fwd_t<__half>((const __half *)w, (const __half *)f, idx, (__half *)out, plan, B, Nq, Nk, C, M, IC, f_sb, f_sn, st);
    return 0;
}
int wf2_fwd_listed(const void *w, const void *f, const int64_t *idx, void *out, const void *plan, int B, int Nq, int Nk, int C, int M,
                   int IC, int64_t f_sb, int64_t f_sn, int dtype, cudaStream_t st) {
    if (dtype == CLUSTEN_BF16)
        return wf2::fwd_t<__nv_bfloat16>((const __nv_bfloat16 *)w, (const __nv_bfloat16 *)f, idx, (__nv_bfloat16 *)out, plan, B, Nq, Nk, C, M, IC, f_sb, f_sn, st, true);
    if (dtype == CLUSTEN_F16)
        return wf2::fwd_t<__half>((const __half *)w, (const __half *)f, idx, (__half *)out, plan, B, Nq, Nk, C, M, IC, f_sb, f_sn, st, true);
    return 0;
}
int wf2_dw(const void *d_out, const void *f, const int64_t *idx, void *d_w, const void *plan, int B, int Nq, int Nk, int C, int M,
           int IC, int64_t f_sb, int64_t f_sn, int dtype, cudaStream_t st) {
    if (dtype == CLUSTEN_BF16)
        return wf2::dw_t<__nv_bfloat16>((const __nv_bfloat16 *)d_out, (const __nv_bfloat16 *)f, idx, (__nv_bfloat16 *)d_w, plan, B, Nq, Nk, C, M, IC, f_sb, f_sn, st);
    if (dtype == CLUSTEN_F16)
        return wf2::dw_t<__half>((const __half *)d_out, (const __half *)f, idx, (__half *)d_w, plan, B, Nq, Nk, C, M, IC, f_sb, f_sn, st);
    return 0;
}
int wf2_df(const void *d_out, const void *w, const int64_t *idx, void *d_f, const void *plan, int B, int Nq, int Nk, int C, int M,
           int IC, int64_t df_sb, int64_t df_sn, int dtype, cudaStream_t st) {
    if (dtype == CLUSTEN_BF16)
        return wf2::df_t<__nv_bfloat16>((const __nv_bfloat16 *)d_out, (const __nv_bfloat16 *)w, idx, (__nv_bfloat16 *)d_f, plan, B, Nq, Nk, C, M, IC, df_sb, df_sn, st);
    if (dtype == CLUSTEN_F16)
        return wf2::df_t<__half>((const __half *)d_out, (const __half *)w, idx, (__half *)d_f, plan, B, Nq, Nk, C, M, IC, df_sb, df_sn, st);
    if (dtype == CLUSTEN_F32)
        return wf2::df32_t((const float *)d_out, (const float *)w, idx, (float *)d_f, plan, B, Nq, Nk, C, M, IC, df_sb, df_sn, st);
    return 0;
}

}  // namespace clusten

using namespace clusten;

extern "C" size_t clusten_wf_plan_bytes(int B, int Nq, int M, int Nk) {
    if (B <= 0 || Nq <= 0 || M <= 0 || Nk <= 0) return 256;
    return wf_plan_layout(B, Nq, M, Nk).total;
}

extern "C" int clusten_wf_plan_build(const int64_t *nbhd_idx, int B, int Nq, int M, int Nk, void *plan, size_t plan_bytes,
                                     void *stream) {
    if (B < 0 || Nq < 0 || M <= 0 || Nk <= 0) return set_error(CLUSTEN_EINVAL, "bad sizes B=%d Nq=%d M=%d Nk=%d", B, Nq, M, Nk);
    if (!nbhd_idx || !plan) return set_error(CLUSTEN_EINVAL, "null pointer");
    if (M % 8) return set_error(CLUSTEN_EUNSUPPORTED, "wf plan needs M %% 8 == 0 (got %d)", M);
    if (plan_bytes < clusten_wf_plan_bytes(B, Nq, M, Nk))
        return set_error(CLUSTEN_EWORKSPACE, "plan buffer too small: %zu < %zu", plan_bytes, clusten_wf_plan_bytes(B, Nq, M, Nk));
    cudaStream_t st = (cudaStream_t)stream;
    const WfPlanLayout L = wf_plan_layout(B, Nq, M, Nk);
    WfPlanView pv = wf_plan_view(plan, B, Nq, M, Nk);
    const size_t smem = ((size_t)pv.NO + 40) * sizeof(int);
    if (pv.S > 32 || Nq >= (1 << 27) || smem > 200 * 1024 || (int64_t)B * Nq * pv.S >= (1LL << 31))
        return set_error(CLUSTEN_EUNSUPPORTED, "wf plan: M=%d / Nq=%d / Nk=%d out of range", M, Nq, Nk);
    cudaMemsetAsync(plan, 0, 256, st);
    if (B == 0 || Nq == 0) return check_launch("wf plan memset");
    cudaMemsetAsync(reinterpret_cast<char *>(plan) + L.zero_begin, 0, L.zero_end - L.zero_begin, st);   // histograms, counters, row flags
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(wf2::plan_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr_set = true;
    }
    const int64_t slots = (int64_t)B * Nq * pv.S;
    wf2::plan_count_kernel<<<ceil_div(slots, 256), 256, 0, st>>>(nbhd_idx, B, Nq, M, Nk, pv);
    wf2::plan_scan_kernel<<<B, wf2::SCAN_THREADS, smem, st>>>(pv);
    wf2::plan_fill_kernel<<<ceil_div(slots, 256), 256, 0, st>>>(B, Nq, pv);
    wf2::plan_sort_kernel<<<ceil_div((int64_t)B * pv.NO, 8), 256, 0, st>>>(B, Nq, pv);
    wf2::plan_tile_kernel<<<ceil_div((int64_t)B * pv.T, wf2::PT_WARPS), wf2::PT_WARPS * 32, 0, st>>>(B, Nq, pv);
    note_launches(5);
    return check_launch("wf_plan_build");
}
