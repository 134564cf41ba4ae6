// CLUSTEN WF (PointConv weighted-feature merge) and WEIGHTEDGATHER (= WF with IC = 1), forward + backward, sm_100a.
//
//   fwd  : out[b,i,ic,c] = sum_j w[b,i,j,ic] * f[b,idx[b,i,j],c]     every neighbour row is gathered ONCE and feeds all
//                                                                    IC accumulators (the reference re-gathers it IC
//                                                                    times, clustenwf_cuda_kernel.cu:41-49)
//   d_w  : d_w[b,i,j,ic] = sum_c f[b,idx,c] * d_out[b,i,ic,c]
//   d_f  : d_f[b,r,:]    = sum_{(i,j)->r} sum_ic w[b,i,j,ic] * d_out[b,i,ic,:]   deterministic CSR gather instead of
//                                                                    the atomics of clustenwf_cuda_kernel.cu:129
// One warp owns one output token (fwd, d_w) or one feature row (d_f) and walks the channel dimension in blocks of
// 32 x 16 bytes; rows with fewer than 32 chunks are shared by lane groups exactly as in clusten_attn.cu.
#include <algorithm>
#include <initializer_list>

#include "common.cuh"
#include "wf2.cuh"

namespace clusten {

constexpr int WF_UNROLL = 2;

template <typename T, int G, int IC>
__global__ void __launch_bounds__(CTA_THREADS)
wf_fwd_kernel(const T *__restrict__ Wt, const T *__restrict__ F, const int64_t *__restrict__ idx, T *__restrict__ out,
              int B, int Nq, int nchunk, int M, int64_t f_sb, int64_t f_sn, const T *__restrict__ outer = nullptr, int Kin = 1) {
    constexpr int VPT = Vec<T>::VPT;
    constexpr int RPI = 32 / G;
    extern __shared__ int smem_i[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int *idx_s = smem_i + warp * (M + M * IC);
    float *w_s = reinterpret_cast<float *>(idx_s + M);
    const int64_t tok = (int64_t)blockIdx.x * WARPS_PER_CTA + warp;
    if (tok >= (int64_t)B * Nq) return;
    const int b = (int)(tok / Nq);
    const int64_t *irow = idx + tok * M;
    const T *wrow = Wt + tok * M * IC;
    for (int j = lane; j < M; j += 32) idx_s[j] = (int)irow[j];
    // MSDETRPC (IC = 1): the weight of gather t is the product attn[t / Kin] * nn_weight[t], formed here (msdetrpc_cuda_kernel.cu:40-47)
    if (outer) { for (int t = lane; t < M; t += 32) w_s[t] = to_f(wrow[t]) * to_f(outer[tok * (M / Kin) + t / Kin]); }
    else { for (int t = lane; t < M * IC; t += 32) w_s[t] = to_f(wrow[t]); }
    __syncwarp();
    const int grp = lane / G, lg = lane % G;
    const int C = nchunk * VPT;
    T *orow = out + tok * IC * C;
    for (int c0 = 0; c0 < nchunk; c0 += G) {          // channel blocks of G chunks
        const int ch = c0 + lg;
        const bool act = ch < nchunk;
        float acc[IC][VPT];
#pragma unroll
        for (int ic = 0; ic < IC; ++ic)
#pragma unroll
            for (int v = 0; v < VPT; ++v) acc[ic][v] = 0.f;
        const T *fbase = F + b * f_sb + ch * VPT;
        for (int j0 = 0; j0 < M; j0 += RPI * WF_UNROLL) {
            float ff[WF_UNROLL][VPT];
            int jj[WF_UNROLL];
#pragma unroll
            for (int u = 0; u < WF_UNROLL; ++u) {
                const int j = j0 + u * RPI + grp;
                jj[u] = (act && j < M) ? j : -1;
#pragma unroll
                for (int v = 0; v < VPT; ++v) ff[u][v] = 0.f;
                if (jj[u] >= 0) load16(fbase + (int64_t)idx_s[j] * f_sn, ff[u]);
            }
#pragma unroll
            for (int u = 0; u < WF_UNROLL; ++u) {
                if (jj[u] < 0) continue;
#pragma unroll
                for (int ic = 0; ic < IC; ++ic) {
                    const float a = w_s[jj[u] * IC + ic];
#pragma unroll
                    for (int v = 0; v < VPT; ++v) acc[ic][v] = fmaf(a, ff[u][v], acc[ic][v]);
                }
            }
        }
#pragma unroll
        for (int ic = 0; ic < IC; ++ic) {
#pragma unroll
            for (int v = 0; v < VPT; ++v) acc[ic][v] = cross_group_sum<G>(acc[ic][v]);
            if (grp == 0 && act) store16(orow + (int64_t)ic * C + ch * VPT, acc[ic]);
        }
    }
}

// Weighted gather with a handful of neighbours (Shepard upsampling, point_utils.py:103-114: K = 4): one THREAD per 16-byte channel
// chunk of an output row; the K rows of a token are all in flight at once, no shared memory, no shuffles; the threads of a token read
// its K indices / weights as broadcast loads and its rows as contiguous segments.
template <typename T, int KMAX>
__global__ void __launch_bounds__(256)
wg_small_kernel(const T *__restrict__ Wt, const T *__restrict__ F, const int64_t *__restrict__ idx, T *__restrict__ out,
                int64_t tokens, int Nq, int nchunk, int K, int64_t f_sb, int64_t f_sn) {
    constexpr int VPT = Vec<T>::VPT;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t tok = t / nchunk;
    if (tok >= tokens) return;
    const int ch = (int)(t - tok * nchunk);
    const int b = (int)(tok / Nq);
    const T *fb = F + b * f_sb + ch * VPT;
    float v[KMAX][VPT], w[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
        if (k < K) {
            w[k] = to_f(__ldg(Wt + tok * K + k));
            load16(fb + __ldg(idx + tok * K + k) * f_sn, v[k]);
        }
    }
    float acc[VPT];
#pragma unroll
    for (int x = 0; x < VPT; ++x) acc[x] = 0.f;
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
        if (k < K) {
#pragma unroll
            for (int x = 0; x < VPT; ++x) acc[x] = fmaf(w[k], v[k][x], acc[x]);
        }
    store16(out + tok * (int64_t)nchunk * VPT + ch * VPT, acc);
}

template <typename T, int G, int IC>
__global__ void __launch_bounds__(CTA_THREADS)
wf_dw_kernel(const T *__restrict__ dO, const T *__restrict__ F, const int64_t *__restrict__ idx, T *__restrict__ dW,
             int B, int Nq, int nchunk, int M, int64_t f_sb, int64_t f_sn, const T *__restrict__ outer = nullptr, int Kin = 1,
             const T *__restrict__ Wt = nullptr, T *__restrict__ dOuter = nullptr) {
    constexpr int VPT = Vec<T>::VPT;
    constexpr int RPI = 32 / G;
    extern __shared__ int smem_i[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int *idx_s = smem_i + warp * (M + M * IC);
    float *dw_s = reinterpret_cast<float *>(idx_s + M);
    const int64_t tok = (int64_t)blockIdx.x * WARPS_PER_CTA + warp;
    if (tok >= (int64_t)B * Nq) return;
    const int b = (int)(tok / Nq);
    const int64_t *irow = idx + tok * M;
    for (int j = lane; j < M; j += 32) idx_s[j] = (int)irow[j];
    for (int t = lane; t < M * IC; t += 32) dw_s[t] = 0.f;
    __syncwarp();
    const int grp = lane / G, lg = lane % G;
    const int C = nchunk * VPT;
    const T *drow = dO + tok * IC * C;
    for (int c0 = 0; c0 < nchunk; c0 += G) {
        const int ch = c0 + lg;
        const bool act = ch < nchunk;
        float df[IC][VPT];
#pragma unroll
        for (int ic = 0; ic < IC; ++ic) {
#pragma unroll
            for (int v = 0; v < VPT; ++v) df[ic][v] = 0.f;
            if (act) load16(drow + (int64_t)ic * C + ch * VPT, df[ic]);
        }
        const T *fbase = F + b * f_sb + ch * VPT;
        for (int j0 = 0; j0 < M; j0 += RPI * WF_UNROLL) {
            float ff[WF_UNROLL][VPT];
#pragma unroll
            for (int u = 0; u < WF_UNROLL; ++u) {
                const int j = j0 + u * RPI + grp;
#pragma unroll
                for (int v = 0; v < VPT; ++v) ff[u][v] = 0.f;
                if (act && j < M) load16(fbase + (int64_t)idx_s[j] * f_sn, ff[u]);
            }
#pragma unroll
            for (int u = 0; u < WF_UNROLL; ++u) {
                const int j = j0 + u * RPI + grp;
#pragma unroll
                for (int ic = 0; ic < IC; ++ic) {
                    float s = 0.f;
#pragma unroll
                    for (int v = 0; v < VPT; ++v) s = fmaf(df[ic][v], ff[u][v], s);
                    s = group_sum<G>(s);
                    if (lg == 0 && j < M) dw_s[j * IC + ic] += s;       // one writer per (j, ic): lane 0 of group grp
                }
            }
        }
        __syncwarp();
    }
    T *wrow = dW + tok * M * IC;
    if (outer) {
        // MSDETRPC: dw_s holds the gradient of the PRODUCT weights; d_nn_weight = it * attn, d_attn = sum_k it * nn_weight
        // (msdetrpc_cuda_kernel.cu:160-177)
        const int Mo = M / Kin;
        for (int t = lane; t < M; t += 32) wrow[t] = from_f<T>(dw_s[t] * to_f(outer[tok * Mo + t / Kin]));
        for (int mo = lane; mo < Mo; mo += 32) {
            float s = 0.f;
            for (int k = 0; k < Kin; ++k) s = fmaf(dw_s[mo * Kin + k], to_f(Wt[tok * M + mo * Kin + k]), s);
            dOuter[tok * Mo + mo] = from_f<T>(s);
        }
    } else {
        for (int t = lane; t < M * IC; t += 32) wrow[t] = from_f<T>(dw_s[t]);
    }
}

template <typename T, int G, int IC>
__global__ void __launch_bounds__(CTA_THREADS)
wf_df_kernel(const T *__restrict__ dO, const T *__restrict__ Wt, const int32_t *__restrict__ offsets,
             const uint32_t *__restrict__ entries, T *__restrict__ dF,
             int B, int Nq, int Nk, int nchunk, int M, int64_t df_sb, int64_t df_sn, const int *__restrict__ run_if,
             const T *__restrict__ outer = nullptr, int Kin = 1) {
    constexpr int VPT = Vec<T>::VPT;
    if (run_if && run_if[0] == 0) return;            // the octet kernels of clusten_wf2.cu own this call (wf2.cuh flags[0])
    constexpr int RPI = 32 / G;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = lane / G, lg = lane % G;
    const int C = nchunk * VPT;
    for (int64_t row = (int64_t)blockIdx.x * WARPS_PER_CTA + warp; row < (int64_t)B * Nk; row += (int64_t)gridDim.x * WARPS_PER_CTA) {
    const int b = (int)(row / Nk), r = (int)(row - (int64_t)b * Nk);
    const int lo = offsets[(int64_t)b * (Nk + 1) + r], hi = offsets[(int64_t)b * (Nk + 1) + r + 1];
    const uint32_t *ent = entries + (int64_t)b * Nq * M;
    const T *wb = Wt + (int64_t)b * Nq * M * IC;
    const T *dob = dO + (int64_t)b * Nq * IC * C;
    for (int c0 = 0; c0 < nchunk; c0 += G) {
        const int ch = c0 + lg;
        const bool act = ch < nchunk;
        float acc[VPT];
#pragma unroll
        for (int v = 0; v < VPT; ++v) acc[v] = 0.f;
        for (int e0 = lo; e0 < hi; e0 += RPI) {
            const int e = e0 + grp;
            if (act && e < hi) {
                const uint32_t pk = __ldg(ent + e);
                const int64_t qi = pk >> 8;
                const T *wp = wb + (qi * M + (pk & 255u)) * IC;
                const T *dp = dob + qi * IC * C + ch * VPT;
                float d[IC][VPT];
#pragma unroll
                for (int ic = 0; ic < IC; ++ic) load16(dp + (int64_t)ic * C, d[ic]);
#pragma unroll
                for (int ic = 0; ic < IC; ++ic) {
                    float a = to_f(wp[ic]);
                    if (outer) a *= to_f(outer[((int64_t)b * Nq + qi) * (M / Kin) + (pk & 255u) / Kin]);      // MSDETRPC product weight
#pragma unroll
                    for (int v = 0; v < VPT; ++v) acc[v] = fmaf(a, d[ic][v], acc[v]);
                }
            }
        }
#pragma unroll
        for (int v = 0; v < VPT; ++v) acc[v] = cross_group_sum<G>(acc[v]);
        if (grp == 0 && act) store16(dF + b * df_sb + (int64_t)r * df_sn + ch * VPT, acc);
    }
    }
}

// ---- scalar fallbacks ----------------------------------------------------------------------------------------
template <typename T>
__global__ void wf_fwd_scalar(const T *Wt, const T *F, const int64_t *idx, T *out, int B, int Nq, int C, int M, int IC,
                              int64_t f_sb, int64_t f_sn) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)B * Nq * IC * C) return;
    const int c = (int)(t % C);
    const int ic = (int)((t / C) % IC);
    const int64_t tok = t / ((int64_t)C * IC);
    const int b = (int)(tok / Nq);
    float s = 0.f;
    for (int j = 0; j < M; ++j)
        s = fmaf(to_f(Wt[(tok * M + j) * IC + ic]), to_f(F[b * f_sb + idx[tok * M + j] * f_sn + c]), s);
    out[t] = from_f<T>(s);
}
template <typename T>
__global__ void wf_dw_scalar(const T *dO, const T *F, const int64_t *idx, T *dW, int B, int Nq, int C, int M, int IC,
                             int64_t f_sb, int64_t f_sn) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)B * Nq * M * IC) return;
    const int ic = (int)(t % IC);
    const int64_t tj = t / IC;
    const int64_t tok = tj / M;
    const int b = (int)(tok / Nq);
    const T *f = F + b * f_sb + idx[tj] * f_sn;
    const T *d = dO + (tok * IC + ic) * C;
    float s = 0.f;
    for (int c = 0; c < C; ++c) s = fmaf(to_f(f[c]), to_f(d[c]), s);
    dW[t] = from_f<T>(s);
}
template <typename T>
__global__ void wf_df_scalar(const T *dO, const T *Wt, const int32_t *offsets, const uint32_t *entries, T *dF,
                             int B, int Nq, int Nk, int C, int M, int IC, int64_t df_sb, int64_t df_sn, const int *run_if) {
    if (run_if && run_if[0] == 0) return;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)B * Nk * C) return;
    const int c = (int)(t % C);
    const int64_t row = t / C;
    const int b = (int)(row / Nk), r = (int)(row - (int64_t)b * Nk);
    const int lo = offsets[(int64_t)b * (Nk + 1) + r], hi = offsets[(int64_t)b * (Nk + 1) + r + 1];
    const uint32_t *ent = entries + (int64_t)b * Nq * M;
    float s = 0.f;
    for (int e = lo; e < hi; ++e) {
        const uint32_t pk = ent[e];
        const int64_t qi = pk >> 8;
        for (int ic = 0; ic < IC; ++ic)
            s = fmaf(to_f(Wt[(((int64_t)b * Nq + qi) * M + (pk & 255u)) * IC + ic]),
                     to_f(dO[(((int64_t)b * Nq + qi) * IC + ic) * C + c]), s);
    }
    dF[b * df_sb + (int64_t)r * df_sn + c] = from_f<T>(s);
}

template <typename T> static bool wf_vec_ok(int C, int IC, const void *f, int64_t f_sb, int64_t f_sn,
                                            std::initializer_list<const void *> contig) {
    constexpr int VPT = Vec<T>::VPT;
    if (C % VPT != 0) return false;
    if (!(IC == 1 || IC == 2 || IC == 4 || IC == 8)) return false;
    if (!aligned16(f) || f_sb % VPT || f_sn % VPT) return false;
    for (const void *p : contig)
        if (!aligned16(p)) return false;
    return true;
}

#define CLUSTEN_DISPATCH_IC(IC_, ...)                              \
    switch (IC_) {                                                 \
        case 1: { constexpr int IC = 1; __VA_ARGS__; break; }      \
        case 2: { constexpr int IC = 2; __VA_ARGS__; break; }      \
        case 4: { constexpr int IC = 4; __VA_ARGS__; break; }      \
        default: { constexpr int IC = 8; __VA_ARGS__; break; }     \
    }

template <typename T>
static int wf_fwd_impl(const T *w, const T *f, const int64_t *idx, T *out, const void *plan, int B, int Nq, int Nk, int C, int M, int IC_,
                       int64_t f_sb, int64_t f_sn, int dtype, cudaStream_t st) {
    if ((int64_t)B * Nq == 0) return 0;
    if (plan && wf3_fwd(w, f, idx, out, plan, B, Nq, Nk, C, M, IC_, f_sb, f_sn, dtype, st) > 0) return check_launch("wf3_fwd");
    if (IC_ == 1 && M <= 8 && wf_vec_ok<T>(C, 1, f, f_sb, f_sn, {out})) {
        const int nchunk = C / Vec<T>::VPT;
        const int64_t threads = (int64_t)B * Nq * nchunk;
        if (M <= 4) wg_small_kernel<T, 4><<<ceil_div(threads, 256), 256, 0, st>>>(w, f, idx, out, (int64_t)B * Nq, Nq, nchunk, M, f_sb, f_sn);
        else wg_small_kernel<T, 8><<<ceil_div(threads, 256), 256, 0, st>>>(w, f, idx, out, (int64_t)B * Nq, Nq, nchunk, M, f_sb, f_sn);
        note_launches(1);
        return check_launch("wg_small");
    }
    if (wf2_fwd(w, f, idx, out, plan, B, Nq, Nk, C, M, IC_, f_sb, f_sn, dtype, st)) return check_launch("wf2_fwd");
    if (wf_vec_ok<T>(C, IC_, f, f_sb, f_sn, {out})) {
        const int nchunk = C / Vec<T>::VPT;
        const int grid = ceil_div((int64_t)B * Nq, WARPS_PER_CTA);
        const size_t smem = (size_t)WARPS_PER_CTA * (M + M * IC_) * sizeof(int);
        CLUSTEN_DISPATCH_IC(IC_, CLUSTEN_DISPATCH_GROUP(pick_group(nchunk),
            (wf_fwd_kernel<T, G, IC><<<grid, CTA_THREADS, smem, st>>>(w, f, idx, out, B, Nq, nchunk, M, f_sb, f_sn))));
    } else {
        const int64_t total = (int64_t)B * Nq * IC_ * C;
        wf_fwd_scalar<T><<<ceil_div(total, 256), 256, 0, st>>>(w, f, idx, out, B, Nq, C, M, IC_, f_sb, f_sn);
    }
    note_launches(1);
    return check_launch("wf_fwd");
}

template <typename T>
static int wf_bwd_impl(const T *d_out, const T *w, const T *f, const int64_t *idx, const int32_t *off,
                       const uint32_t *ent, const void *plan, T *d_w, T *d_f, int B, int Nq, int Nk, int C, int M, int IC_,
                       int64_t f_sb, int64_t f_sn, int64_t df_sb, int64_t df_sn, int dtype, cudaStream_t st) {
    const int *run_if = nullptr;
    if ((int64_t)B * Nq > 0) {
        if (wf2_dw(d_out, f, idx, d_w, plan, B, Nq, Nk, C, M, IC_, f_sb, f_sn, dtype, st)) {
            if (int e = check_launch("wf2_dw")) return e;
        } else {
        if (wf_vec_ok<T>(C, IC_, f, f_sb, f_sn, {d_out})) {
            const int nchunk = C / Vec<T>::VPT;
            const int grid = ceil_div((int64_t)B * Nq, WARPS_PER_CTA);
            const size_t smem = (size_t)WARPS_PER_CTA * (M + M * IC_) * sizeof(int);
            CLUSTEN_DISPATCH_IC(IC_, CLUSTEN_DISPATCH_GROUP(pick_group(nchunk),
                (wf_dw_kernel<T, G, IC><<<grid, CTA_THREADS, smem, st>>>(d_out, f, idx, d_w, B, Nq, nchunk, M, f_sb, f_sn))));
        } else {
            const int64_t total = (int64_t)B * Nq * M * IC_;
            wf_dw_scalar<T><<<ceil_div(total, 256), 256, 0, st>>>(d_out, f, idx, d_w, B, Nq, C, M, IC_, f_sb, f_sn);
        }
        note_launches(1);
        if (int e = check_launch("wf_dw")) return e;
        }
        // octet form of d_f; the generic kernels below then only run when the plan says so (device-side flag)
        if (wf2_df(d_out, w, idx, d_f, plan, B, Nq, Nk, C, M, IC_, df_sb, df_sn, dtype, st)) {
            if (int e = check_launch("wf2_df")) return e;
            run_if = reinterpret_cast<const int *>(plan);
        }
    }
    if ((int64_t)B * Nk > 0) {
        if (wf_vec_ok<T>(C, IC_, d_f, df_sb, df_sn, {d_out})) {
            const int nchunk = C / Vec<T>::VPT;
            // behind a device-side flag the kernel usually exits at once: a short grid (it strides over the rows) keeps that cheap
            const int grid = run_if ? std::min(ceil_div((int64_t)B * Nk, WARPS_PER_CTA), 148 * 8) : ceil_div((int64_t)B * Nk, WARPS_PER_CTA);
            CLUSTEN_DISPATCH_IC(IC_, CLUSTEN_DISPATCH_GROUP(pick_group(nchunk),
                (wf_df_kernel<T, G, IC><<<grid, CTA_THREADS, 0, st>>>(d_out, w, off, ent, d_f, B, Nq, Nk, nchunk, M,
                                                                      df_sb, df_sn, run_if))));
        } else {
            const int64_t total = (int64_t)B * Nk * C;
            wf_df_scalar<T><<<ceil_div(total, 256), 256, 0, st>>>(d_out, w, off, ent, d_f, B, Nq, Nk, C, M, IC_, df_sb, df_sn, run_if);
        }
        note_launches(1);
        if (int e = check_launch("wf_df")) return e;
    }
    return 0;
}

static int wf_check(int B, int Nq, int Nk, int C, int M, int IC) {
    if (B < 0 || Nq < 0 || Nk <= 0 || C <= 0 || M <= 0 || IC <= 0)
        return set_error(CLUSTEN_EINVAL, "bad sizes B=%d Nq=%d Nk=%d C=%d M=%d IC=%d", B, Nq, Nk, C, M, IC);
    if (M > 4096 || IC > 64) return set_error(CLUSTEN_EUNSUPPORTED, "M=%d / IC=%d too large", M, IC);
    return 0;
}

}  // namespace clusten

using namespace clusten;

extern "C" int clusten_wf_fwd(const void *w, const void *f, const int64_t *nbhd_idx, const void *plan, void *out,
                              int B, int Nq, int Nk, int C, int M, int IC, int64_t f_sb, int64_t f_sn,
                              int dtype, void *stream) {
    if (int e = wf_check(B, Nq, Nk, C, M, IC)) return e;
    if (!w || !f || !nbhd_idx || !out) return set_error(CLUSTEN_EINVAL, "null pointer");
    CLUSTEN_DISPATCH_DTYPE(dtype, return wf_fwd_impl<T>((const T *)w, (const T *)f, nbhd_idx, (T *)out, plan, B, Nq, Nk, C, M, IC,
                                                        f_sb, f_sn, dtype, (cudaStream_t)stream));
    return 0;
}

extern "C" int clusten_wf_bwd(const void *d_out, const void *w, const void *f, const int64_t *nbhd_idx,
                              const int32_t *csr_offsets, const uint32_t *csr_entries, const void *plan, void *d_w, void *d_f,
                              int B, int Nq, int Nk, int C, int M, int IC, int64_t f_sb, int64_t f_sn,
                              int64_t df_sb, int64_t df_sn, int dtype, void *stream) {
    if (int e = wf_check(B, Nq, Nk, C, M, IC)) return e;
    if (!d_out || !w || !f || !nbhd_idx || !csr_offsets || !csr_entries || !d_w || !d_f)
        return set_error(CLUSTEN_EINVAL, "null pointer");
    if (M > 256) return set_error(CLUSTEN_EUNSUPPORTED, "backward needs M <= 256 (got %d)", M);
    CLUSTEN_DISPATCH_DTYPE(dtype, return wf_bwd_impl<T>((const T *)d_out, (const T *)w, (const T *)f, nbhd_idx, csr_offsets,
                                                        csr_entries, plan, (T *)d_w, (T *)d_f, B, Nq, Nk, C, M, IC, f_sb, f_sn,
                                                        df_sb, df_sn, dtype, (cudaStream_t)stream));
    return 0;
}

// ---- MSDETRPC: feat[b,i,c] = sum_m attn[b,i,m] * sum_k nn_weight[b,i,m,k] * val[b, nn_idx[b,i,m,k], c]   (msdetrpc_cuda_kernel.cu:18-55)
// One pass: the weighted-gather kernels above with the product weight attn[m] * nn_weight[m,k] formed in shared memory / registers
// (forward, d_val) and the two-term gradient taken from the product's gradient in the same kernel (d_nn_weight, d_attn).
template <typename T>
static int msd_fwd_impl(const int64_t *idx, const T *w, const T *attn, const T *val, T *out, int B, int N, int Nk, int C, int M, int K,
                        int64_t v_sb, int64_t v_sn, cudaStream_t st) {
    if ((int64_t)B * N == 0) return 0;
    if (!wf_vec_ok<T>(C, 1, val, v_sb, v_sn, {out})) return set_error(CLUSTEN_EUNSUPPORTED, "MSDETRPC: C=%d / alignment outside the vector path", C);
    const int nchunk = C / Vec<T>::VPT, MK = M * K;
    const int grid = ceil_div((int64_t)B * N, WARPS_PER_CTA);
    const size_t smem = (size_t)WARPS_PER_CTA * 2 * MK * sizeof(int);
    CLUSTEN_DISPATCH_GROUP(pick_group(nchunk), (wf_fwd_kernel<T, G, 1><<<grid, CTA_THREADS, smem, st>>>(w, val, idx, out, B, N, nchunk, MK, v_sb, v_sn, attn, K)));
    note_launches(1);
    return check_launch("msdetrpc_fwd");
}

template <typename T>
static int msd_bwd_impl(const T *d_out, const int64_t *idx, const T *w, const T *attn, const T *val, const int32_t *off, const uint32_t *ent,
                        T *d_w, T *d_attn, T *d_val, int B, int N, int Nk, int C, int M, int K, int64_t v_sb, int64_t v_sn, int64_t dv_sb,
                        int64_t dv_sn, cudaStream_t st) {
    if (!wf_vec_ok<T>(C, 1, val, v_sb, v_sn, {d_out}) || !wf_vec_ok<T>(C, 1, d_val, dv_sb, dv_sn, {d_out}))
        return set_error(CLUSTEN_EUNSUPPORTED, "MSDETRPC: C=%d / alignment outside the vector path", C);
    const int nchunk = C / Vec<T>::VPT, MK = M * K;
    if ((int64_t)B * N > 0) {
        const int grid = ceil_div((int64_t)B * N, WARPS_PER_CTA);
        const size_t smem = (size_t)WARPS_PER_CTA * 2 * MK * sizeof(int);
        CLUSTEN_DISPATCH_GROUP(pick_group(nchunk), (wf_dw_kernel<T, G, 1><<<grid, CTA_THREADS, smem, st>>>(d_out, val, idx, d_w, B, N, nchunk, MK, v_sb, v_sn,
                                                                                                   attn, K, w, d_attn)));
        note_launches(1);
        if (int e = check_launch("msdetrpc_dw")) return e;
    }
    if ((int64_t)B * Nk > 0) {
        const int grid = ceil_div((int64_t)B * Nk, WARPS_PER_CTA);
        CLUSTEN_DISPATCH_GROUP(pick_group(nchunk), (wf_df_kernel<T, G, 1><<<grid, CTA_THREADS, 0, st>>>(d_out, w, off, ent, d_val, B, N, Nk, nchunk, MK, dv_sb, dv_sn,
                                                                                                nullptr, attn, K)));
        note_launches(1);
        if (int e = check_launch("msdetrpc_dval")) return e;
    }
    return 0;
}

extern "C" int clusten_msdetrpc_fwd(const int64_t *nn_idx, const void *nn_weight, const void *attn, const void *val, void *out,
                                    int B, int N, int Nk, int C, int M, int K, int64_t v_sb, int64_t v_sn, int dtype, void *stream) {
    if (int e = wf_check(B, N, Nk, C, M * (K > 0 ? K : 1), 1)) return e;
    if (K <= 0 || M * K > 256) return set_error(CLUSTEN_EUNSUPPORTED, "MSDETRPC: M*K = %d outside 1..256", M * K);
    if (!nn_idx || !nn_weight || !attn || !val || !out) return set_error(CLUSTEN_EINVAL, "null pointer");
    CLUSTEN_DISPATCH_DTYPE(dtype, return msd_fwd_impl<T>(nn_idx, (const T *)nn_weight, (const T *)attn, (const T *)val, (T *)out, B, N, Nk, C, M, K,
                                                         v_sb, v_sn, (cudaStream_t)stream));
    return 0;
}

extern "C" int clusten_msdetrpc_bwd(const void *d_out, const int64_t *nn_idx, const void *nn_weight, const void *attn, const void *val,
                                    const int32_t *csr_offsets, const uint32_t *csr_entries, void *d_weight, void *d_attn, void *d_val,
                                    int B, int N, int Nk, int C, int M, int K, int64_t v_sb, int64_t v_sn, int64_t dv_sb, int64_t dv_sn,
                                    int dtype, void *stream) {
    if (int e = wf_check(B, N, Nk, C, M * (K > 0 ? K : 1), 1)) return e;
    if (K <= 0 || M * K > 256) return set_error(CLUSTEN_EUNSUPPORTED, "MSDETRPC: M*K = %d outside 1..256", M * K);
    if (!d_out || !nn_idx || !nn_weight || !attn || !val || !csr_offsets || !csr_entries || !d_weight || !d_attn || !d_val)
        return set_error(CLUSTEN_EINVAL, "null pointer");
    CLUSTEN_DISPATCH_DTYPE(dtype, return msd_bwd_impl<T>((const T *)d_out, nn_idx, (const T *)nn_weight, (const T *)attn, (const T *)val, csr_offsets,
                                                         csr_entries, (T *)d_weight, (T *)d_attn, (T *)d_val, B, N, Nk, C, M, K, v_sb, v_sn,
                                                         dv_sb, dv_sn, (cudaStream_t)stream));
    return 0;
}

extern "C" int clusten_wg_fwd(const int64_t *nbhd_idx, const void *w, const void *f, void *out,
                              int B, int Nq, int Nk, int C, int K, int64_t f_sb, int64_t f_sn, int dtype, void *stream) {
    return clusten_wf_fwd(w, f, nbhd_idx, nullptr, out, B, Nq, Nk, C, K, 1, f_sb, f_sn, dtype, stream);
}

extern "C" int clusten_wg_bwd(const void *d_out, const int64_t *nbhd_idx, const void *w, const void *f,
                              const int32_t *csr_offsets, const uint32_t *csr_entries, void *d_w, void *d_f,
                              int B, int Nq, int Nk, int C, int K, int64_t f_sb, int64_t f_sn,
                              int64_t df_sb, int64_t df_sn, int dtype, void *stream) {
    return clusten_wf_bwd(d_out, w, f, nbhd_idx, csr_offsets, csr_entries, nullptr, d_w, d_f, B, Nq, Nk, C, K, 1, f_sb, f_sn,
                          df_sb, df_sn, dtype, stream);
}
