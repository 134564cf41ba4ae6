// CLUSTEN QK / AV forward + backward for sm_100a.
//
// Three warp-level kernel shapes cover the four entry points (one warp owns one token and loops over heads, so the
// 8-byte neighbour indices are read from HBM once per token, not once per head / per channel as in the reference):
//   dot   : out[b,h,i,j] = sum_c X[b,h,i,c] * Y[b,h,idx[b,i,j],c]        QK fwd (X=q,Y=k); AV bwd d_attn (X=d_feat,Y=v)
//   axpy  : out[b,h,i,:] = sum_j W[b,h,i,j] * Y[b,h,idx[b,i,j],:]        AV fwd (W=attn,Y=v); QK bwd d_q (W=d_attn,Y=k)
//   csr   : out[b,h,r,:] = sum_{(i,j)->r} W[b,h,i,j] * X[b,h,i,:]        QK bwd d_k (W=d_attn,X=q); AV bwd d_v (W=attn,X=d_feat)
// Neighbour rows are fetched with 128-bit loads: a group of G lanes covers one row (G*16 bytes), so one warp-wide load
// instruction brings in 32/G complete rows; reductions are warp shuffles.  The csr shape replaces the reference's
// B*H*N*M*C global atomics (clustenqk_cuda_kernel.cu:125, clustenav_cuda_kernel.cu:121) by a deterministic gather
// over the inverse neighbour list built by clusten_csr_build.
#include <initializer_list>

#include "tile.cuh"

namespace clusten {

constexpr int UNROLL = 4;   // independent 128-bit row loads in flight per lane

template <typename T, int G>
__global__ void __launch_bounds__(CTA_THREADS)
dot_rows_kernel(const T *__restrict__ X, const T *__restrict__ Y, const int64_t *__restrict__ idx, T *__restrict__ out,
                int B, int H, int Nq, int nchunk, int M,
                int64_t x_sb, int64_t x_sh, int64_t x_sn, int64_t y_sb, int64_t y_sh, int64_t y_sn,
                const int *__restrict__ tile_flag) {
    if (tile_flag && tile_flag[0] == 0) return;        // the tile-union kernel (clusten_tile.cu) takes this call
    constexpr int VPT = Vec<T>::VPT;
    constexpr int RPI = 32 / G;
    extern __shared__ int smem_i[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int *idx_s = smem_i + warp * 2 * M;
    float *out_s = reinterpret_cast<float *>(idx_s + M);
    const int grp = lane / G, lg = lane % G;
    const bool act = lg < nchunk;
    // grid-stride over tokens (b * Nq + i): one pass when the grid covers them all, a short grid when this launch is
    // only the fallback of the tile-union kernel
    for (int64_t tok = (int64_t)blockIdx.x * WARPS_PER_CTA + warp; tok < (int64_t)B * Nq; tok += (int64_t)gridDim.x * WARPS_PER_CTA) {
    const int b = (int)(tok / Nq), i = (int)(tok - (int64_t)b * Nq);
    const int64_t *irow = idx + tok * M;
    __syncwarp();
    for (int j = lane; j < M; j += 32) idx_s[j] = (int)irow[j];
    __syncwarp();
    for (int h = 0; h < H; ++h) {
        float xf[VPT];
#pragma unroll
        for (int v = 0; v < VPT; ++v) xf[v] = 0.f;
        if (act) load16(X + b * x_sb + h * x_sh + (int64_t)i * x_sn + lg * VPT, xf);
        const T *ybase = Y + b * y_sb + h * y_sh + lg * VPT;
        for (int j0 = 0; j0 < M; j0 += RPI * UNROLL) {
            float yf[UNROLL][VPT];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int j = j0 + u * RPI + grp;
#pragma unroll
                for (int v = 0; v < VPT; ++v) yf[u][v] = 0.f;
                if (act && j < M) load16(ybase + (int64_t)idx_s[j] * y_sn, yf[u]);
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int j = j0 + u * RPI + grp;
                float s = 0.f;
#pragma unroll
                for (int v = 0; v < VPT; ++v) s = fmaf(xf[v], yf[u][v], s);
                s = group_sum<G>(s);
                if (lg == 0 && j < M) out_s[j] = s;
            }
        }
        __syncwarp();
        T *orow = out + ((int64_t)(b * H + h) * Nq + i) * M;
        for (int j = lane; j < M; j += 32) orow[j] = from_f<T>(out_s[j]);
        __syncwarp();
    }
    }
}

template <typename T, int G>
__global__ void __launch_bounds__(CTA_THREADS)
axpy_rows_kernel(const T *__restrict__ W, const T *__restrict__ Y, const int64_t *__restrict__ idx, T *__restrict__ out,
                 int B, int H, int Nq, int nchunk, int M,
                 int64_t w_sb, int64_t w_sh, int64_t w_sn, int64_t y_sb, int64_t y_sh, int64_t y_sn,
                 int64_t o_sb, int64_t o_sh, int64_t o_sn, const int *__restrict__ tile_flag) {
    if (tile_flag && tile_flag[0] == 0) return;
    constexpr int VPT = Vec<T>::VPT;
    constexpr int RPI = 32 / G;
    extern __shared__ int smem_i[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int *idx_s = smem_i + warp * 2 * M;
    float *w_s = reinterpret_cast<float *>(idx_s + M);
    const int grp = lane / G, lg = lane % G;
    const bool act = lg < nchunk;
    for (int64_t tok = (int64_t)blockIdx.x * WARPS_PER_CTA + warp; tok < (int64_t)B * Nq; tok += (int64_t)gridDim.x * WARPS_PER_CTA) {
    const int b = (int)(tok / Nq), i = (int)(tok - (int64_t)b * Nq);
    const int64_t *irow = idx + tok * M;
    __syncwarp();
    for (int j = lane; j < M; j += 32) idx_s[j] = (int)irow[j];
    for (int h = 0; h < H; ++h) {
        __syncwarp();
        const T *wrow = W + b * w_sb + h * w_sh + (int64_t)i * w_sn;
        for (int j = lane; j < M; j += 32) w_s[j] = to_f(wrow[j]);
        __syncwarp();
        float acc[VPT];
#pragma unroll
        for (int v = 0; v < VPT; ++v) acc[v] = 0.f;
        const T *ybase = Y + b * y_sb + h * y_sh + lg * VPT;
        for (int j0 = 0; j0 < M; j0 += RPI * UNROLL) {
            float yf[UNROLL][VPT];
            float a[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int j = j0 + u * RPI + grp;
                a[u] = 0.f;
#pragma unroll
                for (int v = 0; v < VPT; ++v) yf[u][v] = 0.f;
                if (act && j < M) {
                    a[u] = w_s[j];
                    load16(ybase + (int64_t)idx_s[j] * y_sn, yf[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
#pragma unroll
                for (int v = 0; v < VPT; ++v) acc[v] = fmaf(a[u], yf[u][v], acc[v]);
        }
#pragma unroll
        for (int v = 0; v < VPT; ++v) acc[v] = cross_group_sum<G>(acc[v]);
        if (grp == 0 && act) store16(out + b * o_sb + h * o_sh + (int64_t)i * o_sn + lg * VPT, acc);
    }
    }
}

template <typename T, int G>
__global__ void __launch_bounds__(CTA_THREADS)
csr_rows_kernel(const T *__restrict__ W, const T *__restrict__ X, const int32_t *__restrict__ offsets,
                const uint32_t *__restrict__ entries, T *__restrict__ out,
                int B, int H, int Nq, int Nk, int nchunk, int M,
                int64_t w_sb, int64_t w_sh, int64_t w_sn, int64_t x_sb, int64_t x_sh, int64_t x_sn,
                int64_t o_sb, int64_t o_sh, int64_t o_sn, const int *__restrict__ tile_flag) {
    if (tile_flag && tile_flag[0] == 0) return;
    constexpr int VPT = Vec<T>::VPT;
    constexpr int RPI = 32 / G;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = lane / G, lg = lane % G;
    const bool act = lg < nchunk;
    for (int64_t row = (int64_t)blockIdx.x * WARPS_PER_CTA + warp; row < (int64_t)B * Nk; row += (int64_t)gridDim.x * WARPS_PER_CTA) {   // b * Nk + r
    const int b = (int)(row / Nk), r = (int)(row - (int64_t)b * Nk);
    const int lo = offsets[(int64_t)b * (Nk + 1) + r], hi = offsets[(int64_t)b * (Nk + 1) + r + 1];
    const uint32_t *ent = entries + (int64_t)b * Nq * M;
    for (int h = 0; h < H; ++h) {
        float acc[VPT];
#pragma unroll
        for (int v = 0; v < VPT; ++v) acc[v] = 0.f;
        const T *wbase = W + b * w_sb + h * w_sh;
        const T *xbase = X + b * x_sb + h * x_sh + lg * VPT;
        for (int e0 = lo; e0 < hi; e0 += RPI * UNROLL) {
            float xf[UNROLL][VPT];
            float a[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int e = e0 + u * RPI + grp;
                a[u] = 0.f;
#pragma unroll
                for (int v = 0; v < VPT; ++v) xf[u][v] = 0.f;
                if (act && e < hi) {
                    const uint32_t pk = __ldg(ent + e);
                    const int64_t qi = pk >> 8;
                    a[u] = to_f(wbase[qi * w_sn + (pk & 255u)]);
                    load16(xbase + qi * x_sn, xf[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
#pragma unroll
                for (int v = 0; v < VPT; ++v) acc[v] = fmaf(a[u], xf[u][v], acc[v]);
        }
#pragma unroll
        for (int v = 0; v < VPT; ++v) acc[v] = cross_group_sum<G>(acc[v]);
        if (grp == 0 && act) store16(out + b * o_sb + h * o_sh + (int64_t)r * o_sn + lg * VPT, acc);
    }
    }
}

// ---- scalar fallbacks: any C / stride / alignment; one thread per output element -------------------------------
template <typename T>
__global__ void dot_rows_scalar(const T *X, const T *Y, const int64_t *idx, T *out, int B, int H, int Nq, int C, int M,
                                int64_t x_sb, int64_t x_sh, int64_t x_sn, int64_t y_sb, int64_t y_sh, int64_t y_sn,
                                const int *tile_flag) {
    if (tile_flag && tile_flag[0] == 0) return;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < (int64_t)B * H * Nq * M; t += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(t % M);
    const int64_t u = t / M;
    const int i = (int)(u % Nq);
    const int h = (int)((u / Nq) % H), b = (int)(u / ((int64_t)Nq * H));
    const T *x = X + b * x_sb + h * x_sh + (int64_t)i * x_sn;
    const T *y = Y + b * y_sb + h * y_sh + idx[((int64_t)b * Nq + i) * M + j] * y_sn;
    float s = 0.f;
    for (int c = 0; c < C; ++c) s = fmaf(to_f(x[c]), to_f(y[c]), s);
    out[t] = from_f<T>(s);
    }
}
template <typename T>
__global__ void axpy_rows_scalar(const T *W, const T *Y, const int64_t *idx, T *out, int B, int H, int Nq, int C, int M,
                                 int64_t w_sb, int64_t w_sh, int64_t w_sn, int64_t y_sb, int64_t y_sh, int64_t y_sn,
                                 int64_t o_sb, int64_t o_sh, int64_t o_sn, const int *tile_flag) {
    if (tile_flag && tile_flag[0] == 0) return;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < (int64_t)B * H * Nq * C; t += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(t % C);
    const int64_t u = t / C;
    const int i = (int)(u % Nq);
    const int h = (int)((u / Nq) % H), b = (int)(u / ((int64_t)Nq * H));
    const T *w = W + b * w_sb + h * w_sh + (int64_t)i * w_sn;
    const int64_t *ir = idx + ((int64_t)b * Nq + i) * M;
    const T *y = Y + b * y_sb + h * y_sh + c;
    float s = 0.f;
    for (int j = 0; j < M; ++j) s = fmaf(to_f(w[j]), to_f(y[ir[j] * y_sn]), s);
    out[b * o_sb + h * o_sh + (int64_t)i * o_sn + c] = from_f<T>(s);
    }
}
template <typename T>
__global__ void csr_rows_scalar(const T *W, const T *X, const int32_t *offsets, const uint32_t *entries, T *out,
                                int B, int H, int Nq, int Nk, int C, int M,
                                int64_t w_sb, int64_t w_sh, int64_t w_sn, int64_t x_sb, int64_t x_sh, int64_t x_sn,
                                int64_t o_sb, int64_t o_sh, int64_t o_sn, const int *tile_flag) {
    if (tile_flag && tile_flag[0] == 0) return;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < (int64_t)B * H * Nk * C; t += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(t % C);
    const int64_t u = t / C;
    const int r = (int)(u % Nk);
    const int h = (int)((u / Nk) % H), b = (int)(u / ((int64_t)Nk * H));
    const int lo = offsets[(int64_t)b * (Nk + 1) + r], hi = offsets[(int64_t)b * (Nk + 1) + r + 1];
    const uint32_t *ent = entries + (int64_t)b * Nq * M;
    float s = 0.f;
    for (int e = lo; e < hi; ++e) {
        const uint32_t pk = ent[e];
        const int64_t qi = pk >> 8;
        s = fmaf(to_f(W[b * w_sb + h * w_sh + qi * w_sn + (pk & 255u)]), to_f(X[b * x_sb + h * x_sh + qi * x_sn + c]), s);
    }
    out[b * o_sb + h * o_sh + (int64_t)r * o_sn + c] = from_f<T>(s);
    }
}

// ---- host-side launchers --------------------------------------------------------------------------------------
struct Rows { const void *p; int64_t sb, sh, sn; };

template <typename T> bool tile_dot_eligible(int C, int M, Rows4 x, Rows4 y);
template <typename T> bool tile_axpy_eligible(int C, int M, Rows4 w, Rows4 y, Rows4 o);
template <typename T> bool tile_scat_eligible(int C, int M, Rows4 w, Rows4 x, Rows4 o);
template <typename T> int launch_dot_tile(const T *X, const T *Y, const int64_t *idx, const void *pack, T *out, int B, int H, int Nq,
                                          int Nk, int C, int M, Rows4 x, Rows4 y, cudaStream_t st);
template <typename T> int launch_axpy_tile(const T *W, const T *Y, const int64_t *idx, const void *pack, T *out, int B, int H, int Nq,
                                           int Nk, int C, int M, Rows4 w, Rows4 y, Rows4 o, cudaStream_t st);
template <typename T> int launch_scat_tile(const T *W, const T *X, const int32_t *csr_off, const uint32_t *csr_ent, const void *pack,
                                           T *out, int B, int H, int Nq, int Nk, int C, int M, Rows4 w, Rows4 x, Rows4 o, cudaStream_t st);

template <typename T> int launch_dot_tile2(const T *X, const T *Y, const int64_t *idx, const void *pack, T *out, int B, int H, int Nq,
                                           int Nk, int C, int M, Rows4 x, Rows4 y, cudaStream_t st);
template <typename T> int launch_axpy_tile2(const T *W, const T *Y, const int64_t *idx, const void *pack, T *out, int B, int H, int Nq,
                                            int Nk, int C, int M, Rows4 w, Rows4 y, Rows4 o, cudaStream_t st);

template <typename T> int launch_scat_tile2(const T *W, const T *X, const int32_t *csr_off, const uint32_t *csr_ent, const void *pack,
                                            T *out, int B, int H, int Nq, int Nk, int C, int M, Rows4 w, Rows4 x, Rows4 o, cudaStream_t st);

static inline Rows4 r4(const Rows &r) { return Rows4{r.p, r.sb, r.sh, r.sn}; }
static inline const int *tile_flag_of(const void *pack) { return reinterpret_cast<const int *>(pack); }   // PackView.flags is at offset 0
// grid of a generic kernel: everything when it is the only kernel, a short grid-stride grid when it is the (usually
// idle) fallback of a tile-union launch, so that skipping it costs microseconds
static inline int fallback_grid(int full, const int *flag) { return flag ? (full < 148 * 8 ? full : 148 * 8) : full; }

template <typename T> static bool vec_ok(int C, std::initializer_list<Rows> rows) {
    constexpr int VPT = Vec<T>::VPT;
    if (C % VPT != 0 || C / VPT > 32) return false;
    for (const Rows &r : rows)
        if (!aligned16(r.p) || r.sb % VPT || r.sh % VPT || r.sn % VPT) return false;
    return true;
}

static int check_common(int B, int H, int Nq, int Nk, int C, int M) {
    if (B < 0 || H <= 0 || Nq < 0 || Nk <= 0 || C <= 0 || M <= 0)
        return set_error(CLUSTEN_EINVAL, "bad sizes B=%d H=%d Nq=%d Nk=%d C=%d M=%d", B, H, Nq, Nk, C, M);
    if ((int64_t)B * H * Nq * M >= (1LL << 40) || M > 4096)
        return set_error(CLUSTEN_EUNSUPPORTED, "problem too large (M=%d)", M);
    return 0;
}

// Every launcher enqueues the tile-union kernel (when a pack is given and the shape qualifies) AND the generic kernel;
// the pack's device-side flag lets exactly one of them do the work, so no host synchronisation is needed.
template <typename T>
static int launch_dot(const T *X, const T *Y, const int64_t *idx, const void *pack, T *out, int B, int H, int Nq, int Nk,
                      int C, int M, Rows x, Rows y, cudaStream_t st) {
    if ((int64_t)B * Nq == 0) return 0;
    const int *flag = nullptr;
    if (pack && tile_dot_eligible<T>(C, M, r4(x), r4(y))) {
        int e = launch_dot_tile2<T>(X, Y, idx, pack, out, B, H, Nq, Nk, C, M, r4(x), r4(y), st);      // second generation first
        if (e < 0) e = launch_dot_tile<T>(X, Y, idx, pack, out, B, H, Nq, Nk, C, M, r4(x), r4(y), st);
        if (e) return e;
        flag = tile_flag_of(pack);
    }
    if (vec_ok<T>(C, {x, y})) {
        const int nchunk = C / Vec<T>::VPT;
        const int grid = fallback_grid(ceil_div((int64_t)B * Nq, WARPS_PER_CTA), flag);
        const size_t smem = (size_t)WARPS_PER_CTA * 2 * M * sizeof(int);
        CLUSTEN_DISPATCH_GROUP(pick_group(nchunk),
            (dot_rows_kernel<T, G><<<grid, CTA_THREADS, smem, st>>>(X, Y, idx, out, B, H, Nq, nchunk, M,
                                                                     x.sb, x.sh, x.sn, y.sb, y.sh, y.sn, flag)));
    } else {
        const int64_t total = (int64_t)B * H * Nq * M;
        dot_rows_scalar<T><<<fallback_grid(ceil_div(total, 256), flag), 256, 0, st>>>(X, Y, idx, out, B, H, Nq, C, M,
                                                                 x.sb, x.sh, x.sn, y.sb, y.sh, y.sn, flag);
    }
    note_launches(1);
    return check_launch("dot_rows");
}

template <typename T>
static int launch_axpy(const T *W, const T *Y, const int64_t *idx, const void *pack, T *out, int B, int H, int Nq, int Nk,
                       int C, int M, Rows w, Rows y, Rows o, cudaStream_t st) {
    if ((int64_t)B * Nq == 0) return 0;
    const int *flag = nullptr;
    if (pack && tile_axpy_eligible<T>(C, M, r4(w), r4(y), r4(o))) {
        int e = launch_axpy_tile2<T>(W, Y, idx, pack, out, B, H, Nq, Nk, C, M, r4(w), r4(y), r4(o), st);
        if (e < 0) e = launch_axpy_tile<T>(W, Y, idx, pack, out, B, H, Nq, Nk, C, M, r4(w), r4(y), r4(o), st);
        if (e) return e;
        flag = tile_flag_of(pack);
    }
    if (vec_ok<T>(C, {y, o})) {
        const int nchunk = C / Vec<T>::VPT;
        const int grid = fallback_grid(ceil_div((int64_t)B * Nq, WARPS_PER_CTA), flag);
        const size_t smem = (size_t)WARPS_PER_CTA * 2 * M * sizeof(int);
        CLUSTEN_DISPATCH_GROUP(pick_group(nchunk),
            (axpy_rows_kernel<T, G><<<grid, CTA_THREADS, smem, st>>>(W, Y, idx, out, B, H, Nq, nchunk, M,
                                                                      w.sb, w.sh, w.sn, y.sb, y.sh, y.sn,
                                                                      o.sb, o.sh, o.sn, flag)));
    } else {
        const int64_t total = (int64_t)B * H * Nq * C;
        axpy_rows_scalar<T><<<fallback_grid(ceil_div(total, 256), flag), 256, 0, st>>>(W, Y, idx, out, B, H, Nq, C, M, w.sb, w.sh, w.sn,
                                                                  y.sb, y.sh, y.sn, o.sb, o.sh, o.sn, flag);
    }
    note_launches(1);
    return check_launch("axpy_rows");
}

template <typename T>
static int launch_csr(const T *W, const T *X, const int32_t *off, const uint32_t *ent, const void *pack, T *out,
                      int B, int H, int Nq, int Nk, int C, int M, Rows w, Rows x, Rows o, cudaStream_t st) {
    if ((int64_t)B * Nk == 0) return 0;
    const int *flag = nullptr;
    if (pack && tile_scat_eligible<T>(C, M, r4(w), r4(x), r4(o))) {
        int e = launch_scat_tile2<T>(W, X, off, ent, pack, out, B, H, Nq, Nk, C, M, r4(w), r4(x), r4(o), st);
        if (e < 0) e = launch_scat_tile<T>(W, X, off, ent, pack, out, B, H, Nq, Nk, C, M, r4(w), r4(x), r4(o), st);
        if (e) return e;
        flag = tile_flag_of(pack);
    }
    if (vec_ok<T>(C, {x, o})) {
        const int nchunk = C / Vec<T>::VPT;
        const int grid = fallback_grid(ceil_div((int64_t)B * Nk, WARPS_PER_CTA), flag);
        CLUSTEN_DISPATCH_GROUP(pick_group(nchunk),
            (csr_rows_kernel<T, G><<<grid, CTA_THREADS, 0, st>>>(W, X, off, ent, out, B, H, Nq, Nk, nchunk, M,
                                                                  w.sb, w.sh, w.sn, x.sb, x.sh, x.sn,
                                                                  o.sb, o.sh, o.sn, flag)));
    } else {
        const int64_t total = (int64_t)B * H * Nk * C;
        csr_rows_scalar<T><<<fallback_grid(ceil_div(total, 256), flag), 256, 0, st>>>(W, X, off, ent, out, B, H, Nq, Nk, C, M, w.sb, w.sh,
                                                                 w.sn, x.sb, x.sh, x.sn, o.sb, o.sh, o.sn, flag);
    }
    note_launches(1);
    return check_launch("csr_rows");
}

}  // namespace clusten

using namespace clusten;

extern "C" int clusten_qk_fwd(const void *q, const void *k, const int64_t *nbhd_idx, const void *pack, void *attn,
                              int B, int H, int Nq, int Nk, int C, int M,
                              int64_t q_sb, int64_t q_sh, int64_t q_sn, int64_t k_sb, int64_t k_sh, int64_t k_sn,
                              int dtype, void *stream) {
    if (int e = check_common(B, H, Nq, Nk, C, M)) return e;
    if (!q || !k || !nbhd_idx || !attn) return set_error(CLUSTEN_EINVAL, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    CLUSTEN_DISPATCH_DTYPE(dtype, return launch_dot<T>((const T *)q, (const T *)k, nbhd_idx, pack, (T *)attn, B, H, Nq, Nk, C, M,
                                                       Rows{q, q_sb, q_sh, q_sn}, Rows{k, k_sb, k_sh, k_sn}, st));
    return 0;
}

extern "C" int clusten_qk_bwd(const void *d_attn, const void *q, const void *k, const int64_t *nbhd_idx,
                              const int32_t *csr_offsets, const uint32_t *csr_entries, const void *pack, void *d_q, void *d_k,
                              int B, int H, int Nq, int Nk, int C, int M,
                              int64_t q_sb, int64_t q_sh, int64_t q_sn, int64_t k_sb, int64_t k_sh, int64_t k_sn,
                              int64_t dq_sb, int64_t dq_sh, int64_t dq_sn, int64_t dk_sb, int64_t dk_sh, int64_t dk_sn,
                              int dtype, void *stream) {
    if (int e = check_common(B, H, Nq, Nk, C, M)) return e;
    if (!d_attn || !q || !k || !nbhd_idx || !csr_offsets || !csr_entries || !d_q || !d_k)
        return set_error(CLUSTEN_EINVAL, "null pointer");
    if (M > 256) return set_error(CLUSTEN_EUNSUPPORTED, "backward needs M <= 256 (got %d)", M);
    cudaStream_t st = (cudaStream_t)stream;
    const Rows da{d_attn, (int64_t)H * Nq * M, (int64_t)Nq * M, (int64_t)M};
    CLUSTEN_DISPATCH_DTYPE(dtype, {
        if (int e = launch_axpy<T>((const T *)d_attn, (const T *)k, nbhd_idx, pack, (T *)d_q, B, H, Nq, Nk, C, M, da,
                                   Rows{k, k_sb, k_sh, k_sn}, Rows{d_q, dq_sb, dq_sh, dq_sn}, st)) return e;
        return launch_csr<T>((const T *)d_attn, (const T *)q, csr_offsets, csr_entries, pack, (T *)d_k, B, H, Nq, Nk, C, M, da,
                             Rows{q, q_sb, q_sh, q_sn}, Rows{d_k, dk_sb, dk_sh, dk_sn}, st);
    });
    return 0;
}

extern "C" int clusten_av_fwd(const void *attn, const void *v, const int64_t *nbhd_idx, const void *pack, void *feat,
                              int B, int H, int Nq, int Nk, int C, int M,
                              int64_t a_sb, int64_t a_sh, int64_t a_sn, int64_t v_sb, int64_t v_sh, int64_t v_sn,
                              int64_t f_sb, int64_t f_sh, int64_t f_sn, int dtype, void *stream) {
    if (int e = check_common(B, H, Nq, Nk, C, M)) return e;
    if (!attn || !v || !nbhd_idx || !feat) return set_error(CLUSTEN_EINVAL, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    CLUSTEN_DISPATCH_DTYPE(dtype, return launch_axpy<T>((const T *)attn, (const T *)v, nbhd_idx, pack, (T *)feat, B, H, Nq, Nk, C, M,
                                                        Rows{attn, a_sb, a_sh, a_sn}, Rows{v, v_sb, v_sh, v_sn},
                                                        Rows{feat, f_sb, f_sh, f_sn}, st));
    return 0;
}

extern "C" int clusten_av_bwd(const void *d_feat, const void *attn, const void *v, const int64_t *nbhd_idx,
                              const int32_t *csr_offsets, const uint32_t *csr_entries, const void *pack, void *d_attn, void *d_v,
                              int B, int H, int Nq, int Nk, int C, int M,
                              int64_t df_sb, int64_t df_sh, int64_t df_sn, int64_t a_sb, int64_t a_sh, int64_t a_sn,
                              int64_t v_sb, int64_t v_sh, int64_t v_sn, int64_t dv_sb, int64_t dv_sh, int64_t dv_sn,
                              int dtype, void *stream) {
    if (int e = check_common(B, H, Nq, Nk, C, M)) return e;
    if (!d_feat || !attn || !v || !nbhd_idx || !csr_offsets || !csr_entries || !d_attn || !d_v)
        return set_error(CLUSTEN_EINVAL, "null pointer");
    if (M > 256) return set_error(CLUSTEN_EUNSUPPORTED, "backward needs M <= 256 (got %d)", M);
    cudaStream_t st = (cudaStream_t)stream;
    CLUSTEN_DISPATCH_DTYPE(dtype, {
        if (int e = launch_dot<T>((const T *)d_feat, (const T *)v, nbhd_idx, pack, (T *)d_attn, B, H, Nq, Nk, C, M,
                                  Rows{d_feat, df_sb, df_sh, df_sn}, Rows{v, v_sb, v_sh, v_sn}, st)) return e;
        return launch_csr<T>((const T *)attn, (const T *)d_feat, csr_offsets, csr_entries, pack, (T *)d_v, B, H, Nq, Nk, C, M,
                             Rows{attn, a_sb, a_sh, a_sn}, Rows{d_feat, df_sb, df_sh, df_sn},
                             Rows{d_v, dv_sb, dv_sh, dv_sn}, st);
    });
    return 0;
}

// out[b,h,r,:] = sum_{(i,j): idx[b,i,j] = r} w[b,h,i,j] * x[b,h,i,:] -- the scatter shape on its own (d_k / d_v of the fused
// attention backward, which produces its own w tensors).
extern "C" int clusten_scatter_rows(const void *w, const void *x, const int32_t *csr_offsets, const uint32_t *csr_entries,
                                    const void *pack, void *out, int B, int H, int Nq, int Nk, int C, int M,
                                    int64_t w_sb, int64_t w_sh, int64_t w_sn, int64_t x_sb, int64_t x_sh, int64_t x_sn,
                                    int64_t o_sb, int64_t o_sh, int64_t o_sn, int dtype, void *stream) {
    if (int e = check_common(B, H, Nq, Nk, C, M)) return e;
    if (!w || !x || !csr_offsets || !csr_entries || !out) return set_error(CLUSTEN_EINVAL, "null pointer");
    if (M > 256) return set_error(CLUSTEN_EUNSUPPORTED, "scatter needs M <= 256 (got %d)", M);
    cudaStream_t st = (cudaStream_t)stream;
    CLUSTEN_DISPATCH_DTYPE(dtype, return launch_csr<T>((const T *)w, (const T *)x, csr_offsets, csr_entries, pack, (T *)out, B, H, Nq, Nk, C, M,
                                                       Rows{w, w_sb, w_sh, w_sn}, Rows{x, x_sb, x_sh, x_sn}, Rows{out, o_sb, o_sh, o_sn}, st));
    return 0;
}
