// Brute-force 2-D kNN for sm_100a -- replaces the pykeops argKmin / Kmin_argKmin call of knn_keops
// (mask2former/modeling/backbone/point_utils.py:28-60).
//
// One thread per query keeps its k best (distance, index) pairs sorted in registers; database points stream through
// shared memory in tiles.  Arithmetic is pinned for bit-exactness against the oracle:
//     dx = qx - px; dy = qy - py; s = fl(fl(dx*dx) + fl(dy*dy)); r = sqrt_rn(s)          (no FMA contraction)
// and candidates are compared on r (NOT on s: sqrt merges distinct s into equal r, which changes tie groups).
// The database is scanned in ascending index and insertion uses strict '<', so equal distances keep the lower index
// first: the canonical tie rule (kept under the out-of-order scan below by ordering on (distance, index)).  The square root is only taken for the few candidates that can still enter the list:
// sqrt_rn is monotone, so s > T^2 (1 + 2^-20) (T = current k-th distance; the factor covers the roundings of T = sqrt_rn(s_k)
// and of T*T) implies r >= T, which the strict '<' rejects anyway.
#include "common.cuh"

namespace clusten {

constexpr int KNN_THREADS = 128;
constexpr int KNN_TILE = 1024;                          // database points staged per chunk
constexpr int KNN_SUB = 32;                             // points per bounding box

// Bounding boxes make the scan sub-quadratic when the database is spatially coherent in index order (cluster centres along
// the space-filling curve, point_utils.py:203-212; stem-grid points in raster order, msdeformattn_pc.py:295): every chunk of
// 1024 staged points carries 32 boxes of 32 consecutive points (computed on the fly by the CTA) and one box of the whole
// chunk.  A box whose squared distance to the query exceeds thr cannot hold a candidate (the 1 - 1e-6 factor covers the
// roundings of the bound against those of s), so the warp skips it when all its lanes agree, the CTA when all its threads
// agree.  The chunk the CTA's own queries map to is scanned first so that thr is tight before the sweep; because the scan
// is then no longer in ascending index, insertion orders candidates by (distance, index) -- the same canonical result.
template <int K>
__global__ void __launch_bounds__(KNN_THREADS)
knn_kernel(const float2 *__restrict__ query, const float2 *__restrict__ db, int Nq, int Ndb,
           int64_t *__restrict__ idx_out, float *__restrict__ dist_out) {
    __shared__ float2 tile[KNN_TILE];
    __shared__ float4 box[KNN_TILE / KNN_SUB + 1];      // (xmin, ymin, xmax, ymax) per 32 points; last = the whole chunk
    const int b = blockIdx.y;
    const int qi = blockIdx.x * KNN_THREADS + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool valid = qi < Nq;
    const float2 q = query[(int64_t)b * Nq + min(qi, Nq - 1)];      // surplus threads shadow the last query (uniform control flow)
    float bd[K];
    int bi[K];
#pragma unroll
    for (int t = 0; t < K; ++t) { bd[t] = __int_as_float(0x7f800000); bi[t] = 0x7fffffff; }
    float thr = __int_as_float(0x7f800000);             // candidates with s > thr cannot enter the list
    const float2 *dbb = db + (int64_t)b * Ndb;
    const int nchunk = (Ndb + KNN_TILE - 1) / KNN_TILE;
    const int home = min((int)(((int64_t)blockIdx.x * KNN_THREADS * Ndb) / Nq) / KNN_TILE, nchunk - 1);
    for (int it = 0; it < nchunk; ++it) {
        const int ch = it == 0 ? home : (it <= home ? it - 1 : it);          // home first, then 0 .. nchunk-1 without it
        const int t0 = ch * KNN_TILE;
        const int cnt = min(KNN_TILE, Ndb - t0);
        __syncthreads();
        for (int x = threadIdx.x; x < cnt; x += KNN_THREADS) tile[x] = dbb[t0 + x];
        __syncthreads();
        const int nsub = (cnt + KNN_SUB - 1) / KNN_SUB;
        for (int sb = warp; sb < nsub; sb += KNN_THREADS / 32) {
            const int x = sb * KNN_SUB + lane;
            const float2 p = tile[min(x, cnt - 1)];
            float x0 = p.x, x1 = p.x, y0 = p.y, y1 = p.y;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                x0 = fminf(x0, __shfl_xor_sync(FULL, x0, o)); x1 = fmaxf(x1, __shfl_xor_sync(FULL, x1, o));
                y0 = fminf(y0, __shfl_xor_sync(FULL, y0, o)); y1 = fmaxf(y1, __shfl_xor_sync(FULL, y1, o));
            }
            if (lane == 0) box[sb] = make_float4(x0, y0, x1, y1);
        }
        __syncthreads();
        if (warp == 0) {
            const float4 bx = box[min(lane, nsub - 1)];
            float x0 = bx.x, y0 = bx.y, x1 = bx.z, y1 = bx.w;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                x0 = fminf(x0, __shfl_xor_sync(FULL, x0, o)); x1 = fmaxf(x1, __shfl_xor_sync(FULL, x1, o));
                y0 = fminf(y0, __shfl_xor_sync(FULL, y0, o)); y1 = fmaxf(y1, __shfl_xor_sync(FULL, y1, o));
            }
            if (lane == 0) box[KNN_TILE / KNN_SUB] = make_float4(x0, y0, x1, y1);
        }
        __syncthreads();
        auto far = [&](const float4 bx) {               // true: no point of the box can enter this query's list
            const float ex = fmaxf(fmaxf(bx.x - q.x, q.x - bx.z), 0.f), ey = fmaxf(fmaxf(bx.y - q.y, q.y - bx.w), 0.f);
            return (ex * ex + ey * ey) * 0.999999f > thr;
        };
        if (__syncthreads_and(far(box[KNN_TILE / KNN_SUB]))) continue;
        // in the home chunk the warp starts at the box its own queries map to and wraps around: thr is tight after one box
        int sb0 = 0;
        if (it == 0) sb0 = min(max((int)(((int64_t)(blockIdx.x * KNN_THREADS + warp * 32) * Ndb) / Nq) - t0, 0) / KNN_SUB, nsub - 1);
        for (int j = 0; j < nsub; ++j) {
            const int sb = j + sb0 < nsub ? j + sb0 : j + sb0 - nsub;
            if (__all_sync(FULL, far(box[sb]))) continue;
            const int xe = min((sb + 1) * KNN_SUB, cnt);
            for (int x = sb * KNN_SUB; x < xe; ++x) {
                const float2 p = tile[x];
                const float dx = __fsub_rn(q.x, p.x), dy = __fsub_rn(q.y, p.y);
                const float s = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
                if (s > thr) continue;
                const float r = __fsqrt_rn(s);
                const int pi = t0 + x;
                if (r < bd[K - 1] || (r == bd[K - 1] && pi < bi[K - 1])) {
#pragma unroll
                    for (int t = K - 1; t > 0; --t) {
                        const bool in = r < bd[t] || (r == bd[t] && pi < bi[t]);                 // lands at or before slot t
                        const bool sh = r < bd[t - 1] || (r == bd[t - 1] && pi < bi[t - 1]);     // ... strictly before it
                        const float nd = sh ? bd[t - 1] : r;
                        const int ni = sh ? bi[t - 1] : pi;
                        bd[t] = in ? nd : bd[t];
                        bi[t] = in ? ni : bi[t];
                    }
                    if (r < bd[0] || (r == bd[0] && pi < bi[0])) { bd[0] = r; bi[0] = pi; }
                    thr = __fmul_rn(__fmul_rn(bd[K - 1], bd[K - 1]), 1.00000095367431640625f);   // T^2 (1 + 2^-20); inf stays inf
                }
            }
        }
    }
    if (valid) {
        int64_t *io = idx_out + ((int64_t)b * Nq + qi) * K;
#pragma unroll
        for (int t = 0; t < K; ++t) io[t] = min(bi[t], Ndb - 1);      // (unfilled slots only with NaN coordinates: keep the index valid)
        if (dist_out) {
            float *dd = dist_out + ((int64_t)b * Nq + qi) * K;
#pragma unroll
            for (int t = 0; t < K; ++t) dd[t] = bd[t];
        }
    }
}

}  // namespace clusten

using namespace clusten;

extern "C" int clusten_knn(const float *query, const float *database, int B, int Nq, int Ndb, int k,
                           int64_t *idx_out, float *dist_out, void *stream) {
    if (B < 0 || Nq < 0 || Ndb <= 0) return set_error(CLUSTEN_EINVAL, "bad sizes B=%d Nq=%d Ndb=%d", B, Nq, Ndb);
    if (k < 1 || k > 16 || k > Ndb) return set_error(CLUSTEN_EUNSUPPORTED, "knn needs 1 <= k <= min(16, Ndb) (k=%d Ndb=%d)", k, Ndb);
    if (!query || !database || !idx_out) return set_error(CLUSTEN_EINVAL, "null pointer");
    if (B == 0 || Nq == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 grid(ceil_div(Nq, KNN_THREADS), B);
    const float2 *q = reinterpret_cast<const float2 *>(query);
    const float2 *d = reinterpret_cast<const float2 *>(database);
#define KNN_CASE(KK) case KK: knn_kernel<KK><<<grid, KNN_THREADS, 0, st>>>(q, d, Nq, Ndb, idx_out, dist_out); break;
    switch (k) {
        KNN_CASE(1) KNN_CASE(2) KNN_CASE(3) KNN_CASE(4) KNN_CASE(5) KNN_CASE(6) KNN_CASE(7) KNN_CASE(8)
        KNN_CASE(9) KNN_CASE(10) KNN_CASE(11) KNN_CASE(12) KNN_CASE(13) KNN_CASE(14) KNN_CASE(15) KNN_CASE(16)
    }
#undef KNN_CASE
    note_launches(1);
    return check_launch("knn");
}
