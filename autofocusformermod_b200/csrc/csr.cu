// Inverse neighbour list ("CSR over key rows") + library-wide error plumbing.
//
// For an index tensor idx[B,Nq,M] the backward of every CLUSTEN op scatters into the rows idx points at.  The
// reference does that with one global atomic per (i, j, c) (clustenqk_cuda_kernel.cu:125 etc.).  Here the scatter is
// turned into a gather: entries (i<<8 | j) are grouped by target row r = idx[b,i,j] with a stable radix sort, so the
// list of every row is in ascending (i, j) order -> the summation order is fixed -> gradients are deterministic.
// The list depends only on idx, which is constant across all blocks of an AFF stage (aff.py:487-493), so one build is
// amortised over 2*depth backward kernels.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace clusten {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void note_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int set_error(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int check_launch(const char *what) {
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

__global__ void csr_keys_kernel(const int64_t *__restrict__ idx, uint32_t *__restrict__ keys, int64_t total, const int *skip) {
    if (skip && skip[0] == 0) return;
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < total) keys[p] = (uint32_t)idx[p];
}

// sorted keys/vals of one segment -> row offsets + packed entries
__global__ void csr_finalize_kernel(const uint32_t *__restrict__ skeys, const uint32_t *__restrict__ svals,
                                    int32_t *__restrict__ offsets, uint32_t *__restrict__ entries,
                                    int nseg, int Nk, int M, const int *skip) {
    if (skip && skip[0] == 0) return;
    const int b = blockIdx.y;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nseg) return;
    const uint32_t *k = skeys + (int64_t)b * nseg;
    int32_t *off = offsets + (int64_t)b * (Nk + 1);
    const int key = (int)min(k[p], (uint32_t)Nk);                 // out-of-range indices (UB in the reference) are clamped
    const int prev = p ? (int)min(k[p - 1], (uint32_t)Nk) : -1;
    for (int r = prev + 1; r <= key; ++r) off[r] = p;
    if (p == nseg - 1)
        for (int r = key + 1; r <= Nk; ++r) off[r] = nseg;
    const uint32_t v = svals[(int64_t)b * nseg + p];              // = i*M + j
    const uint32_t i = v / (uint32_t)M;
    entries[(int64_t)b * nseg + p] = (i << 8) | (v - i * (uint32_t)M);
}

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace clusten

using namespace clusten;

extern "C" int clusten_abi_version(void) { return CLUSTEN_ABI_VERSION; }
extern "C" const char *clusten_last_error(void) { return g_err; }
extern "C" long long clusten_kernel_launches(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" size_t clusten_csr_workspace_bytes(int B, int Nq, int M, int Nk) {
    (void)Nk;
    const size_t tot = (size_t)B * Nq * M;
    return 4 * align256(tot * sizeof(uint32_t)) + radix_sort_workspace_bytes(B, Nq * M) + 256;
}

extern "C" int clusten_csr_build(const int64_t *nbhd_idx, int B, int Nq, int M, int Nk, int32_t *offsets,
                                 uint32_t *entries, void *workspace, size_t workspace_bytes, const void *pack, void *stream) {
    if (B < 0 || Nq < 0 || M <= 0 || Nk <= 0) return set_error(CLUSTEN_EINVAL, "bad sizes B=%d Nq=%d M=%d Nk=%d", B, Nq, M, Nk);
    if (M > 256 || Nq >= (1 << 24) || (int64_t)Nq * M >= (1LL << 31))
        return set_error(CLUSTEN_EUNSUPPORTED, "csr needs M <= 256, Nq < 2^24, Nq*M < 2^31 (M=%d Nq=%d)", M, Nq);
    if (!nbhd_idx || !offsets || !entries || !workspace) return set_error(CLUSTEN_EINVAL, "null pointer");
    if (workspace_bytes < clusten_csr_workspace_bytes(B, Nq, M, Nk))
        return set_error(CLUSTEN_EWORKSPACE, "workspace too small: %zu < %zu", workspace_bytes,
                         clusten_csr_workspace_bytes(B, Nq, M, Nk));
    cudaStream_t st = (cudaStream_t)stream;
    if (B == 0) return 0;
    // with a tile pack the list is only needed when the pack routes to the generic kernels or has impure tokens: every kernel
    // below exits at once otherwise -- decided on the device, no host synchronisation
    const int *skip = pack ? reinterpret_cast<const int *>(pack) + 4 : nullptr;     // PackView.flags[4]: list needed
    const int nseg = Nq * M;
    if (nseg == 0) {
        cudaMemsetAsync(offsets, 0, (size_t)B * (Nk + 1) * sizeof(int32_t), st);
        return check_launch("csr memset");
    }
    const size_t tot = (size_t)B * nseg;
    const size_t stride = align256(tot * sizeof(uint32_t));
    char *ws = reinterpret_cast<char *>(workspace);
    uint32_t *kA = reinterpret_cast<uint32_t *>(ws);
    uint32_t *kB = reinterpret_cast<uint32_t *>(ws + stride);
    uint32_t *vA = reinterpret_cast<uint32_t *>(ws + 2 * stride);
    uint32_t *vB = reinterpret_cast<uint32_t *>(ws + 3 * stride);
    void *hist = ws + 4 * stride;
    csr_keys_kernel<<<ceil_div((int64_t)tot, 256), 256, 0, st>>>(nbhd_idx, kA, (int64_t)tot, skip);
    // an EVEN number of passes lets the sorted keys land back in kA (see radix_sort_pairs ping-pong)
    int bits = 1;
    while ((1LL << bits) < Nk) ++bits;
    const int passes = bits <= 16 ? 2 : 4;
    if (int e = radix_sort_pairs(kA, nullptr, kB, vB, kA, vA, B, nseg, passes * 8, hist, st, skip)) return e;
    const dim3 grid(ceil_div(nseg, 256), B);
    csr_finalize_kernel<<<grid, 256, 0, st>>>(kA, vA, offsets, entries, nseg, Nk, M, skip);
    note_launches(2);
    return check_launch("csr_build");
}
