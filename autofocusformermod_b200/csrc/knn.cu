// Brute-force 2-D kNN for sm_100a -- replaces the pykeops argKmin / Kmin_argKmin call of knn_keops
// (mask2former/modeling/backbone/point_utils.py:28-60).
//
// One thread per query keeps its k best (distance, index) pairs sorted in registers; database points stream through
// shared memory in tiles.  Arithmetic is pinned for bit-exactness against the oracle:
//     dx = qx - px; dy = qy - py; s = fl(fl(dx*dx) + fl(dy*dy)); r = sqrt_rn(s)          (no FMA contraction)
// and candidates are compared on r (NOT on s: sqrt merges distinct s into equal r, which changes tie groups).
// The database is scanned in ascending index and insertion uses strict '<', so equal distances keep the lower index
// first: the canonical tie rule.  The square root is only taken for the few candidates that can still enter the list:
// sqrt_rn is monotone, so s > T^2 (1 + 2^-20) (T = current k-th distance; the factor covers the roundings of T = sqrt_rn(s_k)
// and of T*T) implies r >= T, which the strict '<' rejects anyway.
#include "common.cuh"

namespace clusten {

constexpr int KNN_THREADS = 128;
constexpr int KNN_TILE = 1024;

template <int K>
__global__ void __launch_bounds__(KNN_THREADS)
knn_kernel(const float2 *__restrict__ query, const float2 *__restrict__ db, int Nq, int Ndb,
           int64_t *__restrict__ idx_out, float *__restrict__ dist_out) {
    __shared__ float2 tile[KNN_TILE];
    const int b = blockIdx.y;
    const int qi = blockIdx.x * KNN_THREADS + threadIdx.x;
    const bool valid = qi < Nq;
    const float2 q = valid ? query[(int64_t)b * Nq + qi] : make_float2(0.f, 0.f);
    float bd[K];
    int bi[K];
#pragma unroll
    for (int t = 0; t < K; ++t) { bd[t] = __int_as_float(0x7f800000); bi[t] = 0; }
    float thr = __int_as_float(0x7f800000);             // candidates with s > thr cannot enter the list
    const float2 *dbb = db + (int64_t)b * Ndb;
    for (int t0 = 0; t0 < Ndb; t0 += KNN_TILE) {
        const int cnt = min(KNN_TILE, Ndb - t0);
        __syncthreads();
        for (int x = threadIdx.x; x < cnt; x += KNN_THREADS) tile[x] = dbb[t0 + x];
        __syncthreads();
        for (int x = 0; x < cnt; ++x) {
            const float2 p = tile[x];
            const float dx = __fsub_rn(q.x, p.x), dy = __fsub_rn(q.y, p.y);
            const float s = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
            if (s > thr) continue;
            const float r = __fsqrt_rn(s);
            if (r < bd[K - 1]) {
                const int pi = t0 + x;
#pragma unroll
                for (int t = K - 1; t > 0; --t) {
                    const bool in = r < bd[t];          // new element lands at or before slot t
                    const bool sh = r < bd[t - 1];      // ... strictly before: slot t takes its left neighbour
                    const float nd = sh ? bd[t - 1] : r;
                    const int ni = sh ? bi[t - 1] : pi;
                    bd[t] = in ? nd : bd[t];
                    bi[t] = in ? ni : bi[t];
                }
                if (r < bd[0]) { bd[0] = r; bi[0] = pi; }
                thr = __fmul_rn(__fmul_rn(bd[K - 1], bd[K - 1]), 1.00000095367431640625f);      // T^2 (1 + 2^-20); inf stays inf
            }
        }
    }
    if (valid) {
        int64_t *io = idx_out + ((int64_t)b * Nq + qi) * K;
#pragma unroll
        for (int t = 0; t < K; ++t) io[t] = bi[t];
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
