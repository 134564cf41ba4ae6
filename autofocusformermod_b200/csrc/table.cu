// Relative-position table lookup and its gradient (sm_100a).
//
// The reference evaluates pos_embed / weight_net on the whole 1023^2-row pre_table and gathers it with pe_idx
// (mask2former/modeling/backbone/aff.py:129-132, 346-349).  The caller here evaluates the network on the U table rows
// a stage actually references and gathers with the inverse map:
//   fwd : out[e, c]  = tab[inv[e], c]                     e < n, c < CH        (a plain gather; output dtype = table dtype)
//   bwd : d_tab[r,c] = sum_{e : inv[e] = r} d_out[e, c]   r < U                (fp32 accumulation; d_out addressed as
//         base + (e / n_per)*d_sb + (e % n_per)*d_se + c*d_sc, so the permuted [B,H,N,M] gradient of aff.py:132 needs no copy)
// ATen's backward of `tab[inv]` (index_put_ with accumulate) sorts the n indices and then walks each unique row's
// duplicates serially -- with n = B*N*M ~ 25 M lookups into U ~ 10^3 rows that is hundreds of milliseconds per block.
// Here every CTA accumulates its slice of e into a shared-memory copy of the table (fp32 shared atomics; lanes of a warp
// hold consecutive neighbours of one token = distinct rows, so intra-warp conflicts are rare) and flushes the non-zero
// entries with one global fp32 atomic each.  Large tables (U*CH floats > 48 KB) use global atomics directly.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace clusten {

constexpr int TG_SMEM_FLOATS = 50 * 1024;        // 200 KB of the 227 KB a CTA may use

// blockIdx.y = channel split (channels [c_lo, c_lo + chs) of the table live in this CTA's shared memory).  U_dev (optional,
// device scalar): the number of table rows actually referenced (<= U, the row count the host knows): only that many rows
// are cleared and flushed, so callers may size the table by an upper bound without reading U back to the host.  A slice
// that does not fit the shared-memory budget after all is accumulated with global atomics by the same CTA.
template <typename T, typename I>
__global__ void __launch_bounds__(512)
table_grad_smem_kernel(const T *__restrict__ dout, const I *__restrict__ inv, float *__restrict__ dtab,
                       int64_t n, int U_host, const int *__restrict__ U_dev, int smem_floats, int CH, int chs, int64_t n_per,
                       int64_t d_sb, int64_t d_se, int64_t d_sc, int64_t per_cta) {
    extern __shared__ float tab_s[];
    const int U = U_dev ? min(U_dev[0], U_host) : U_host;
    const int c_lo = blockIdx.y * chs, cw = min(chs, CH - c_lo);
    const int64_t e0 = (int64_t)blockIdx.x * per_cta, e1 = min(n, e0 + per_cta);
    if ((int64_t)U * cw > smem_floats) {
        for (int64_t e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
            const int64_t r = (int64_t)inv[e];
            const int64_t b = e / n_per;
            const T *dp = dout + b * d_sb + (e - b * n_per) * d_se + c_lo * d_sc;
            for (int c = 0; c < cw; ++c) atomicAdd(dtab + r * CH + c_lo + c, to_f(dp[c * d_sc]));
        }
        return;
    }
    const int tot = U * cw;
    for (int x = threadIdx.x; x < tot; x += blockDim.x) tab_s[x] = 0.f;
    __syncthreads();
    {
        // (sample, position in sample) of this thread's entry, advanced incrementally: no 64-bit division per entry
        int64_t e = e0 + threadIdx.x;
        int64_t b = e < e1 ? e / n_per : 0;
        int64_t el = e - b * n_per;
        const T *base = dout + c_lo * d_sc;
        for (; e < e1; e += blockDim.x) {
            const int r = (int)inv[e];
            const T *dp = base + b * d_sb + el * d_se;
            // all channel loads of the entry first, then the shared-memory atomics: issued one by one behind each atomic they
            // serialised on global latency (ncu: 24 long-scoreboard stalls per issue, 122 us for 20 MB)
            for (int cb = 0; cb < cw; cb += 8) {
                float v[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) v[c] = cb + c < cw ? to_f(dp[(int64_t)(cb + c) * d_sc]) : 0.f;
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    if (cb + c < cw) atomicAdd(tab_s + r * cw + cb + c, v[c]);
            }
            el += blockDim.x;
            while (el >= n_per) { el -= n_per; ++b; }
        }
    }
    __syncthreads();
    for (int x = threadIdx.x; x < tot; x += blockDim.x) {
        const float v = tab_s[x];
        if (v != 0.f) { const int r = x / cw; atomicAdd(dtab + r * CH + c_lo + (x - r * cw), v); }
    }
}

// Sorted variant: shared-memory fp32 atomicAdd is a compare-and-swap loop on this architecture (ATOMS.CAST.SPIN in the SASS;
// ncu: 122 us for a 20 MB call), 32-bit integer atomics are native.  So a CTA counting-sorts its entries by table row --
// histogram and scatter cursors with integer atomics, the VALUES of its channel slice scattered into row order in shared
// memory -- and then sums every row's segment without any atomic; one global fp32 RED per non-empty (row, channel).
constexpr int TGS_THREADS = 512;
constexpr int TGS_HIST_CAP = 8192;

template <typename T, typename I>
__global__ void __launch_bounds__(TGS_THREADS)
table_grad_sort_kernel(const T *__restrict__ dout, const I *__restrict__ inv, float *__restrict__ dtab,
                       int64_t n, int U_host, const int *__restrict__ U_dev, int hist_cap, int CH, int chs, int64_t n_per,
                       int64_t d_sb, int64_t d_se, int64_t d_sc, int per_cta) {
    extern __shared__ __align__(16) unsigned char tgs_smem[];
    __shared__ int wsum[34];
    int *hist = reinterpret_cast<int *>(tgs_smem);
    T *vals = reinterpret_cast<T *>(tgs_smem + (((size_t)hist_cap * 4 + 15) & ~(size_t)15));
    const int tid = threadIdx.x;
    const int U = U_dev ? min(U_dev[0], U_host) : U_host;
    const int c_lo = blockIdx.y * chs, cw = min(chs, CH - c_lo);
    const int64_t e0 = (int64_t)blockIdx.x * per_cta, e1 = min(n, e0 + per_cta);
    const T *base = dout + c_lo * d_sc;
    int64_t b0 = e0 / n_per, el0 = e0 - b0 * n_per;                  // (sample, position in sample) of this thread's first entry
    el0 += tid;
    while (el0 >= n_per) { el0 -= n_per; ++b0; }
    if (U > hist_cap) {                                              // table larger than the histogram: plain global REDs
        int64_t b = b0, el = el0;
        for (int64_t e = e0 + tid; e < e1; e += TGS_THREADS) {
            const int64_t r = (int64_t)inv[e];
            const T *dp = base + b * d_sb + el * d_se;
            for (int c = 0; c < cw; ++c) atomicAdd(dtab + r * CH + c_lo + c, to_f(dp[c * d_sc]));
            el += TGS_THREADS;
            while (el >= n_per) { el -= n_per; ++b; }
        }
        return;
    }
    for (int x = tid; x < U; x += TGS_THREADS) hist[x] = 0;
    __syncthreads();
    for (int64_t e = e0 + tid; e < e1; e += TGS_THREADS) atomicAdd(&hist[(int)inv[e]], 1);
    __syncthreads();
    {   // exclusive scan of hist[0, U) in place
        const int per = (U + TGS_THREADS - 1) / TGS_THREADS;
        const int beg = min(tid * per, U), end = min(beg + per, U);
        int s = 0;
        for (int x = beg; x < end; ++x) s += hist[x];
        int incl = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(FULL, incl, o);
            if ((tid & 31) >= o) incl += v;
        }
        if ((tid & 31) == 31) wsum[tid >> 5] = incl;
        __syncthreads();
        if (tid < 32) {
            const int v = tid < TGS_THREADS / 32 ? wsum[tid] : 0;
            int inc2 = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(FULL, inc2, o);
                if (tid >= o) inc2 += u;
            }
            wsum[tid] = inc2 - v;
        }
        __syncthreads();
        int run = wsum[tid >> 5] + incl - s;
        for (int x = beg; x < end; ++x) { const int v = hist[x]; hist[x] = run; run += v; }
    }
    __syncthreads();
    {   // scatter the values of the channel slice into row order; afterwards hist[r] = end of row r's segment
        int64_t b = b0, el = el0;
        for (int64_t e = e0 + tid; e < e1; e += TGS_THREADS) {
            const int r = (int)inv[e];
            const T *dp = base + b * d_sb + el * d_se;
            T v[8];
#pragma unroll
            for (int c = 0; c < 8; ++c)
                if (c < cw) v[c] = dp[(int64_t)c * d_sc];
            const int pos = atomicAdd(&hist[r], 1);
#pragma unroll
            for (int c = 0; c < 8; ++c)
                if (c < cw) vals[pos * cw + c] = v[c];
            el += TGS_THREADS;
            while (el >= n_per) { el -= n_per; ++b; }
        }
    }
    __syncthreads();
    for (int t = tid; t < U * cw; t += TGS_THREADS) {
        const int r = t / cw, c = t - r * cw;
        const int s0 = r ? hist[r - 1] : 0, s1 = hist[r];
        if (s1 > s0) {
            float sum = 0.f;
            for (int k = s0; k < s1; ++k) sum += to_f(vals[k * cw + c]);
            atomicAdd(dtab + (int64_t)r * CH + c_lo + c, sum);
        }
    }
}

template <typename T, typename I>
__global__ void __launch_bounds__(256)
table_grad_global_kernel(const T *__restrict__ dout, const I *__restrict__ inv, float *__restrict__ dtab,
                         int64_t n, int CH, int64_t n_per, int64_t d_sb, int64_t d_se, int64_t d_sc) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = (int64_t)inv[e];
        const int64_t b = e / n_per;
        const T *dp = dout + b * d_sb + (e - b * n_per) * d_se;
        for (int c = 0; c < CH; ++c) atomicAdd(dtab + r * CH + c, to_f(dp[c * d_sc]));
    }
}

template <typename T, typename I>
__global__ void __launch_bounds__(256)
table_gather_kernel(const T *__restrict__ tab, const I *__restrict__ inv, T *__restrict__ out, int64_t n, int CH) {
    // one thread per (e, c); consecutive threads -> consecutive c of one e: coalesced writes, table reads hit L1/L2
    const int64_t tot = n * CH;
    for (int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; x < tot; x += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = x / CH;
        const int c = (int)(x - e * CH);
        out[x] = tab[(int64_t)inv[e] * CH + c];
    }
}

template <typename T, typename I>
static int launch_table_grad(const T *dout, const I *inv, float *dtab, int64_t n, int U, const int *U_dev, int CH, int64_t n_per,
                             int64_t d_sb, int64_t d_se, int64_t d_sc, cudaStream_t st) {
    if (n == 0) return 0;
    static const bool old_path = getenv("CLUSTEN_TG_OLD") != nullptr;        // diagnostics: the shared-memory-atomic variant
    if (!old_path && (U_dev || U <= TGS_HIST_CAP)) {
        // counting-sort variant (no floating-point shared-memory atomics): histogram of up to TGS_HIST_CAP rows, <= 8 channels
        // of <= 8192 entries in row order; tables that turn out larger on the device fall back to global REDs inside the kernel
        const int hist_cap = std::min(U, TGS_HIST_CAP);
        const int chs = std::min(CH, sizeof(T) == 2 ? 4 : 2);
        const int splits = (CH + chs - 1) / chs;
        int per_cta = (int)std::min<int64_t>(8192, std::max<int64_t>(2048, (n + 295) / 296));
        const size_t smem = (((size_t)hist_cap * 4 + 15) & ~(size_t)15) + (size_t)per_cta * chs * sizeof(T);
        static bool attr2 = false;
        if (!attr2) {
            cudaFuncSetAttribute(table_grad_sort_kernel<T, I>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
            attr2 = true;
        }
        const int64_t gx = (n + per_cta - 1) / per_cta;
        table_grad_sort_kernel<T, I><<<dim3((unsigned)gx, splits), TGS_THREADS, smem, st>>>(dout, inv, dtab, n, U, U_dev, hist_cap, CH, chs, n_per,
                                                                                            d_sb, d_se, d_sc, per_cta);
        note_launches(1);
        return check_launch("table_grad_sort");
    }
    // shared-memory accumulation of a channel slice of the table per CTA.  With a device-side row count the host only knows
    // an upper bound of U: slices are sized for the tables AFF stages really produce (a few thousand rows) within a 48 KB
    // budget (4 CTAs per SM); an oversized table falls back to global atomics inside the kernel.
    const int64_t Ug = U_dev ? std::min<int64_t>(U, 6000) : U;
    const int budget = U_dev ? 12 * 1024 : TG_SMEM_FLOATS;
    int splits = 1;
    while (splits < CH && Ug * ((CH + splits - 1) / splits) > budget) ++splits;
    const int chs = (CH + splits - 1) / splits;
    const int smem_floats = (int)std::min<int64_t>((int64_t)U * chs, budget);
    if (Ug * chs <= budget) {
        const size_t smem = (size_t)smem_floats * sizeof(float);
        const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(4, (size_t)(220 * 1024) / (smem + 1024)));
        int64_t grid = std::min<int64_t>((int64_t)148 * per_sm, std::max<int64_t>(1, n / ((int64_t)4 * std::max<int64_t>(Ug, 64))));
        grid = std::min<int64_t>(grid, (n + 2047) / 2048);
        grid = std::max<int64_t>(grid, 1);
        const int64_t per_cta = (n + grid - 1) / grid;
        static bool attr = false;                        // (one flag per (T, I): this function is a template)
        if (!attr) {
            cudaFuncSetAttribute(table_grad_smem_kernel<T, I>, cudaFuncAttributeMaxDynamicSharedMemorySize, TG_SMEM_FLOATS * 4);
            attr = true;
        }
        table_grad_smem_kernel<T, I><<<dim3((unsigned)grid, splits), 512, smem, st>>>(dout, inv, dtab, n, U, U_dev, smem_floats, CH, chs, n_per,
                                                                                     d_sb, d_se, d_sc, per_cta);
    } else {
        const int grid = (int)std::min<int64_t>(148 * 8, (n + 255) / 256);
        table_grad_global_kernel<T, I><<<grid, 256, 0, st>>>(dout, inv, dtab, n, CH, n_per, d_sb, d_se, d_sc);
    }
    note_launches(1);
    return check_launch("table_grad");
}

// Rows of the reference's pre_table (aff.py:21-31) for the given table indices: (dx, dy, dist, dy / dist, dx / dist) with the 0 / 0 centre
// zeroed, dx = idx % 1023 - 511, dy = idx / 1023 - 511.  The torch formulation (floor-divide, remainder, two casts, two squares, add,
// sqrt, two divisions, stack, isfinite, zeros_like, where) is ~12 launches per stage for a few thousand rows; every fp32 operation
// here is the same IEEE operation, rounded separately.
__global__ void __launch_bounds__(256)
rel_pos_features_kernel(const int64_t *__restrict__ rows, float *__restrict__ out, int64_t n) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const int64_t r = rows[e];
    int64_t qy = r / 1023, rx = r - qy * 1023;
    if (rx < 0) { rx += 1023; --qy; }                                  // floor-divide / remainder of ATen (indices are >= 0 on this path)
    const float ys = (float)(qy - 511), xs = (float)(rx - 511);
    const float dis = __fsqrt_rn(__fadd_rn(__fmul_rn(ys, ys), __fmul_rn(xs, xs)));
    float a = __fdiv_rn(ys, dis), b = __fdiv_rn(xs, dis);
    if (!isfinite(a)) a = 0.f;
    if (!isfinite(b)) b = 0.f;
    float *o = out + e * 5;
    o[0] = xs; o[1] = ys; o[2] = dis; o[3] = a; o[4] = b;
}

}  // namespace clusten

using namespace clusten;

extern "C" int clusten_table_gather(const void *tab, const void *inv, int inv_is_i64, void *out, int64_t n, int U, int CH,
                                    int dtype, void *stream) {
    if (n < 0 || U <= 0 || CH <= 0) return set_error(CLUSTEN_EINVAL, "bad sizes n=%lld U=%d CH=%d", (long long)n, U, CH);
    if (!tab || !inv || !out) return set_error(CLUSTEN_EINVAL, "null pointer");
    if (n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = (int)std::min<int64_t>(148 * 16, (n * CH + 255) / 256);
    CLUSTEN_DISPATCH_DTYPE(dtype, {
        if (inv_is_i64) table_gather_kernel<T, int64_t><<<grid, 256, 0, st>>>((const T *)tab, (const int64_t *)inv, (T *)out, n, CH);
        else table_gather_kernel<T, int32_t><<<grid, 256, 0, st>>>((const T *)tab, (const int32_t *)inv, (T *)out, n, CH);
    });
    note_launches(1);
    return check_launch("table_gather");
}

extern "C" int clusten_table_grad(const void *d_out, const void *inv, int inv_is_i64, float *d_tab, int64_t n, int U,
                                  const int32_t *U_dev, int CH, int64_t n_per, int64_t d_sb, int64_t d_se, int64_t d_sc, int dtype,
                                  void *stream) {
    if (n < 0 || U <= 0 || CH <= 0 || n_per <= 0) return set_error(CLUSTEN_EINVAL, "bad sizes n=%lld U=%d CH=%d", (long long)n, U, CH);
    if (!d_out || !inv || !d_tab) return set_error(CLUSTEN_EINVAL, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    CLUSTEN_DISPATCH_DTYPE(dtype, {
        if (inv_is_i64) return launch_table_grad<T, int64_t>((const T *)d_out, (const int64_t *)inv, d_tab, n, U, U_dev, CH, n_per, d_sb, d_se, d_sc, st);
        return launch_table_grad<T, int32_t>((const T *)d_out, (const int32_t *)inv, d_tab, n, U, U_dev, CH, n_per, d_sb, d_se, d_sc, st);
    });
    return 0;
}

extern "C" int clusten_rel_pos_features(const int64_t *rows, float *out, int64_t n, void *stream) {
    if (n < 0 || (n > 0 && (!rows || !out))) return set_error(CLUSTEN_EINVAL, "rel_pos_features: bad arguments");
    if (n == 0) return 0;
    rel_pos_features_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rows, out, n);
    note_launches(1);
    return check_launch("rel_pos_features");
}
