// Row gather  out[b, i, :] = src[b, idx[b, i], :]  (sm_100a).
//
// The AFF stages reorder / subsample token rows with `x.gather(index=idx.expand(-1, -1, c), dim=1)`
// (mask2former/modeling/backbone/aff.py:332,335,340,471): features into cluster order, the kept tokens' positions, member rows and
// masks after a merge.  ATen answers the expanded index with its element-wise gather (one int64 index load and one 4-byte copy per
// ELEMENT): 16 launches and 4 % of the AFF-Mini forward (profiles/r2_torch_profiler_mini_fwd_graph.txt).  Here a row is what it is:
// one index load per row, 16-byte copies.  Rows are opaque bytes, so the same kernel moves fp32 features, int64 member rows and
// uint8 masks.
#include <algorithm>

#include "common.cuh"

namespace clusten {

// one thread per V-sized piece of an output row; consecutive threads -> consecutive pieces of one row (coalesced both ways)
template <typename V>
__global__ void __launch_bounds__(256)
gather_rows_kernel(const V *__restrict__ src, const int64_t *__restrict__ idx, V *__restrict__ out, int64_t rows_out, int n_src,
                   int n_out, int ppr /* pieces per row */, int *__restrict__ bad) {
    const int64_t total = rows_out * ppr;
    for (int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = x / ppr;
        const int p = (int)(x - row * ppr);
        const int64_t b = row / n_out;
        const int64_t r = __ldg(idx + row);
        if (r < 0 || r >= n_src) {                       // torch.gather raises; here: flag it, copy nothing
            if (bad) *bad = 1;
            continue;
        }
        out[x] = __ldg(src + (b * n_src + r) * ppr + p);
    }
}

}  // namespace clusten

using namespace clusten;

extern "C" int clusten_gather_rows(const void *src, const int64_t *idx, void *out, int B, int n_src, int n_out, int row_bytes,
                                   int *bad, void *stream) {
    if (B < 0 || n_src <= 0 || n_out < 0 || row_bytes <= 0) return set_error(CLUSTEN_EINVAL, "gather_rows: bad sizes B=%d n_src=%d n_out=%d row_bytes=%d", B, n_src, n_out, row_bytes);
    if (B == 0 || n_out == 0) return 0;
    if (!src || !idx || !out) return set_error(CLUSTEN_EINVAL, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t rows = (int64_t)B * n_out;
    const uintptr_t al = reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(out) | (uintptr_t)row_bytes;
#define GR_LAUNCH(V_)                                                                                                      \
    {                                                                                                                      \
        const int ppr = row_bytes / (int)sizeof(V_);                                                                       \
        const int grid = (int)std::min<int64_t>((rows * ppr + 255) / 256, 148 * 32);                                       \
        gather_rows_kernel<V_><<<grid, 256, 0, st>>>((const V_ *)src, idx, (V_ *)out, rows, n_src, n_out, ppr, bad);       \
    }
    if ((al & 15u) == 0) GR_LAUNCH(int4)
    else if ((al & 7u) == 0) GR_LAUNCH(int2)
    else if ((al & 3u) == 0) GR_LAUNCH(int)
    else GR_LAUNCH(unsigned char)
#undef GR_LAUNCH
    note_launches(1);
    return check_launch("gather_rows");
}
