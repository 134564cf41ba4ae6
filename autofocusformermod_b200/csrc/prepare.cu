// Stage preparation: neighbourhood assembly + relative-position index of BasicLayer.forward
// (mask2former/modeling/backbone/aff.py:475-485) in one kernel, and the restriction of the 1023^2-row position table to the
// rows the stage references without sorting.
//
//   nearest [B,n,nnc] (kNN of the tokens among the cluster centres) , member [B,k,m] , cluster_mask [B,k,m] or NULL , pos [B,n,2]
//   member_idx[b,i,c*m+r] = member[b, nearest[b,i,c], r]                                                   aff.py:478
//   mask[b,i,c*m+r]       = cluster_mask[b, nearest[b,i,c], r]                                             aff.py:480
//   rel = pos[b, member_idx] - (pos[b,i] - 511), clamped to [0, 1022];  pe = rel.y * 1023 + rel.x          aff.py:481-485
//
// The reference then evaluates pos_embed / weight_net on ALL 1023^2 table rows per block; the torch-level restriction used
// before (torch.unique over the B*n*M int64 indices) is a 25 M-key radix sort per stage.  Here the prepare kernel marks the
// referenced table rows in a 1023^2-byte presence map, a single-CTA-per-chunk scan ranks them (ascending row id = the order
// torch.unique returns), and a second pass rewrites pe -> rank.  Three launches, no sort, one host read (the row count U).
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace clusten {

constexpr int PE_W = 1023;
constexpr int PE_ROWS = PE_W * PE_W;
constexpr float PE_HALF = 511.f;

// one thread per (token, neighbour cluster, group of 4 members): the cluster lookup and the token's position are shared by the
// group, member_idx / mask / pe leave as 32- / 16- / 4-byte vector stores (m % 4 == 0; a scalar tail covers other m)
__global__ void __launch_bounds__(256)
prepare_kernel(const int64_t *__restrict__ nearest, const int64_t *__restrict__ member, const int64_t *__restrict__ cmask,
               const float *__restrict__ pos, int B, int n, int k, int m, int nnc,
               int64_t *__restrict__ member_idx, int64_t *__restrict__ mask64, uint8_t *__restrict__ mask8,
               int32_t *__restrict__ pe, uint8_t *__restrict__ present, int *__restrict__ range) {
    const int M = nnc * m;
    const int G = (m & 3) == 0 ? 4 : 1;                  // members per thread
    const int gpc = m / G;                               // groups per cluster
    const int64_t total = (int64_t)B * n * nnc * gpc;
    int pmin = 0x7fffffff, pmax = -1;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int gq = (int)(t % gpc);
        const int64_t tc = t / gpc;                      // (b * n + i) * nnc + c
        const int c = (int)(tc % nnc);
        const int64_t bi = tc / nnc;                     // b * n + i
        const int b = (int)(bi / n);
        const int64_t cl = nearest[tc];
        const int64_t src = ((int64_t)b * k + cl) * m + gq * G;
        const int64_t e = bi * M + c * m + gq * G;
        const float2 pi = *reinterpret_cast<const float2 *>(pos + bi * 2);
        const float qx = __fsub_rn(pi.x, PE_HALF), qy = __fsub_rn(pi.y, PE_HALF);
        int64_t mi[4], mk[4] = {1, 1, 1, 1};
        int pv[4];
        for (int x = 0; x < G; ++x) mi[x] = member[src + x];
        if (cmask)
            for (int x = 0; x < G; ++x) mk[x] = cmask[src + x];
        for (int x = 0; x < G; ++x) {
            const float2 pn = *reinterpret_cast<const float2 *>(pos + ((int64_t)b * n + mi[x]) * 2);
            float rx = __fsub_rn(pn.x, qx), ry = __fsub_rn(pn.y, qy);
            rx = fminf(fmaxf(rx, 0.f), (float)(PE_W - 1));
            ry = fminf(fmaxf(ry, 0.f), (float)(PE_W - 1));
            pv[x] = (int)__fadd_rn(__fmul_rn(ry, (float)PE_W), rx);       // (rel.y * 1023 + rel.x).long(), fp32 arithmetic
            // benign race: every writer stores 1.  Read first: the B*n*M entries reference only a few thousand distinct rows, and
            // unconditional stores to those few bytes serialise in the L2 slices that own them (648 us at B*n*M = 12.6 M).
            if (present[pv[x]] == 0) present[pv[x]] = 1;
            pmin = min(pmin, pv[x]);
            pmax = max(pmax, pv[x]);
        }
        if (G == 4) {                                    // e % 4 == 0: aligned vector stores
            *reinterpret_cast<longlong4 *>(member_idx + e) = make_longlong4(mi[0], mi[1], mi[2], mi[3]);
            *reinterpret_cast<int4 *>(pe + e) = make_int4(pv[0], pv[1], pv[2], pv[3]);
            if (cmask) {
                if (mask64) *reinterpret_cast<longlong4 *>(mask64 + e) = make_longlong4(mk[0], mk[1], mk[2], mk[3]);
                if (mask8) *reinterpret_cast<uchar4 *>(mask8 + e) = make_uchar4(mk[0] != 0, mk[1] != 0, mk[2] != 0, mk[3] != 0);
            }
        } else {
            member_idx[e] = mi[0];
            pe[e] = pv[0];
            if (cmask) {
                if (mask64) mask64[e] = mk[0];
                if (mask8) mask8[e] = mk[0] != 0;
            }
        }
    }
    // span of the referenced table rows (neighbours are a few grid cells away: a small band of the 1023^2 rows)
    pmin = __reduce_min_sync(FULL, pmin);
    pmax = __reduce_max_sync(FULL, pmax);
    if ((threadIdx.x & 31) == 0 && pmax >= 0) { atomicMin(range, pmin); atomicMax(range + 1, pmax); }
}

// The same pass with the presence marks kept ON CHIP.  The B*n*M entries of a stage reference a few hundred to a few thousand
// distinct table rows, and marking them in the global byte map -- even read-first -- leaves thousands of same-address stores per
// row (the zeros every thread of the first wave reads come from its own L1): stage-prepare was 215 us at B*n*M = 12.6 M where
// the data movement is ~30 us, and 0.2 ms of the 4.9 ms AFF-Mini forward.  Here one CTA per SM keeps a BITMAP of all 1023^2
// rows in shared memory (128 KiB), marks with shared-memory atomics, and touches the global map once per CTA and row at the
// end (reading through L2 first).  The min / max of the referenced rows takes one pair of global atomics per CTA instead of
// one per warp.  Two thread-items are in flight per iteration: the pass is a chain of four dependent loads.
constexpr int PE_WORDS = (PE_ROWS + 31) / 32;            // 32 705 words
constexpr int PBM_THREADS = 1024;
constexpr size_t PBM_SMEM = (size_t)PE_WORDS * 4;

__global__ void __launch_bounds__(PBM_THREADS, 1)
prepare_bm_kernel(const int64_t *__restrict__ nearest, const int64_t *__restrict__ member, const int64_t *__restrict__ cmask,
                  const float *__restrict__ pos, int B, int n, int k, int m, int nnc,
                  int64_t *__restrict__ member_idx, int64_t *__restrict__ mask64, uint8_t *__restrict__ mask8,
                  int32_t *__restrict__ pe, uint8_t *__restrict__ present, int *__restrict__ range) {
    extern __shared__ uint32_t pbm[];
    __shared__ int s_min, s_max;
    for (int w = threadIdx.x; w < PE_WORDS; w += PBM_THREADS) pbm[w] = 0u;
    if (threadIdx.x == 0) { s_min = 0x7fffffff; s_max = -1; }
    __syncthreads();
    const int M = nnc * m;
    const int gpc = m >> 2;                              // groups of 4 members per cluster (m % 4 == 0 on this path)
    const int64_t total = (int64_t)B * n * nnc * gpc;
    const int64_t stride = (int64_t)gridDim.x * PBM_THREADS;
    int pmin = 0x7fffffff, pmax = -1;
    for (int64_t t0 = (int64_t)blockIdx.x * PBM_THREADS + threadIdx.x; t0 < total; t0 += 2 * stride) {
        bool ok[2];
        int64_t src[2], e[2], brow[2];
        float qx[2], qy[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int64_t t = t0 + u * stride;
            ok[u] = t < total;
            const int64_t tt = ok[u] ? t : t0;
            const int gq = (int)(tt % gpc);
            const int64_t tc = tt / gpc;                 // (b * n + i) * nnc + c
            const int c = (int)(tc % nnc);
            const int64_t bi = tc / nnc;                 // b * n + i
            const int b = (int)(bi / n);
            const int64_t cl = __ldg(nearest + tc);
            src[u] = ((int64_t)b * k + cl) * m + gq * 4;
            e[u] = bi * M + c * m + gq * 4;
            brow[u] = (int64_t)b * n;
            const float2 pi = __ldg(reinterpret_cast<const float2 *>(pos + bi * 2));
            qx[u] = __fsub_rn(pi.x, PE_HALF);
            qy[u] = __fsub_rn(pi.y, PE_HALF);
        }
        longlong4 mi[2], mk[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            mi[u] = *reinterpret_cast<const longlong4 *>(member + src[u]);
            mk[u] = cmask ? *reinterpret_cast<const longlong4 *>(cmask + src[u]) : make_longlong4(1, 1, 1, 1);
        }
        float2 pn[2][4];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            pn[u][0] = __ldg(reinterpret_cast<const float2 *>(pos + (brow[u] + mi[u].x) * 2));
            pn[u][1] = __ldg(reinterpret_cast<const float2 *>(pos + (brow[u] + mi[u].y) * 2));
            pn[u][2] = __ldg(reinterpret_cast<const float2 *>(pos + (brow[u] + mi[u].z) * 2));
            pn[u][3] = __ldg(reinterpret_cast<const float2 *>(pos + (brow[u] + mi[u].w) * 2));
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (!ok[u]) continue;
            int pv[4];
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                float rx = __fsub_rn(pn[u][x].x, qx[u]), ry = __fsub_rn(pn[u][x].y, qy[u]);
                rx = fminf(fmaxf(rx, 0.f), (float)(PE_W - 1));
                ry = fminf(fmaxf(ry, 0.f), (float)(PE_W - 1));
                pv[x] = (int)__fadd_rn(__fmul_rn(ry, (float)PE_W), rx);       // (rel.y * 1023 + rel.x).long(), fp32 arithmetic
                const uint32_t bit = 1u << (pv[x] & 31);
                uint32_t *wd = pbm + (pv[x] >> 5);
                if (!(*reinterpret_cast<volatile uint32_t *>(wd) & bit)) atomicOr(wd, bit);
                pmin = min(pmin, pv[x]);
                pmax = max(pmax, pv[x]);
            }
            *reinterpret_cast<longlong4 *>(member_idx + e[u]) = mi[u];
            *reinterpret_cast<int4 *>(pe + e[u]) = make_int4(pv[0], pv[1], pv[2], pv[3]);
            if (cmask) {
                if (mask64) *reinterpret_cast<longlong4 *>(mask64 + e[u]) = mk[u];
                if (mask8) *reinterpret_cast<uchar4 *>(mask8 + e[u]) = make_uchar4(mk[u].x != 0, mk[u].y != 0, mk[u].z != 0, mk[u].w != 0);
            }
        }
    }
    pmin = __reduce_min_sync(FULL, pmin);
    pmax = __reduce_max_sync(FULL, pmax);
    if ((threadIdx.x & 31) == 0 && pmax >= 0) { atomicMin(&s_min, pmin); atomicMax(&s_max, pmax); }
    __syncthreads();
    if (s_max < 0) return;
    // this CTA's marks -> the global byte map the ranking scan reads (benign race: every writer stores 1)
    for (int w = (s_min >> 5) + threadIdx.x; w <= (s_max >> 5); w += PBM_THREADS) {
        uint32_t v = pbm[w];
        while (v) {
            const int p = w * 32 + __ffs(v) - 1;
            v &= v - 1;
            if (__ldcg(present + p) == 0) present[p] = 1;
        }
    }
    if (threadIdx.x == 0) { atomicMin(range, s_min); atomicMax(range + 1, s_max); }
}

// rank of every present table row (exclusive prefix count), U = number of present rows, uniq[rank] = row id.
// One CTA of 1024 threads walks the 1023^2 flags in chunks: 1 M flags, ~1 MB -> a few microseconds; no second kernel, no sync.
__global__ void __launch_bounds__(1024)
rank_kernel(const uint8_t *__restrict__ present, int32_t *__restrict__ rank, int32_t *__restrict__ uniq, int32_t *__restrict__ count,
            int cap, const int *__restrict__ range) {
    __shared__ int wsum[32];
    __shared__ int base_s, chunk_tot;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) base_s = 0;
    __syncthreads();
    constexpr int PER = 16;                              // flags per thread per chunk (one 16-byte load)
    const int lo = range[0], hi = range[1];              // only the chunks that hold referenced rows are walked
    for (int c0 = lo / (1024 * PER) * (1024 * PER); c0 < PE_ROWS && c0 <= hi; c0 += 1024 * PER) {
        const int p0 = c0 + tid * PER;
        uint32_t w[4] = {0u, 0u, 0u, 0u};
        if (p0 + PER <= PE_ROWS) {
            const uint4 v = *reinterpret_cast<const uint4 *>(present + p0);
            w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
        } else {
            for (int x = 0; x < PER; ++x)
                if (p0 + x < PE_ROWS && present[p0 + x]) w[x >> 2] |= 1u << (8 * (x & 3));
        }
        int cnt = 0;
#pragma unroll
        for (int x = 0; x < 4; ++x) cnt += __popc(w[x] & 0x01010101u);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += y;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const int v = wsum[lane];
            int iv = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(FULL, iv, o);
                if (lane >= o) iv += y;
            }
            wsum[lane] = iv - v;                         // exclusive over the warps
            if (lane == 31) chunk_tot = iv;
        }
        __syncthreads();
        int r = base_s + wsum[warp] + incl - cnt;
#pragma unroll
        for (int x = 0; x < PER; ++x) {
            if ((w[x >> 2] >> (8 * (x & 3))) & 1u) {
                rank[p0 + x] = r;
                if (r < cap) uniq[r] = p0 + x;
                ++r;
            }
        }
        __syncthreads();
        if (tid == 0) base_s += chunk_tot;
        __syncthreads();
    }
    if (tid == 0) *count = base_s;
}

__global__ void __launch_bounds__(256)
rerank_kernel(const int32_t *pe, const int32_t *__restrict__ rank, int32_t *inv, int64_t total) {      // (pe may alias inv)
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x)
        inv[e] = __ldg(rank + pe[e]);
}

// standalone ranking of table rows given as int64 indices (PointConv's pe_idx, msdeformattn_pc.py:305-306): mark + narrow to int32
__global__ void __launch_bounds__(256)
mark_rows_kernel(const int64_t *__restrict__ pe64, int64_t total, int32_t *__restrict__ pe32, uint8_t *__restrict__ present, int *__restrict__ range) {
    int pmin = 0x7fffffff, pmax = -1;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int v = (int)min(max(pe64[e], (int64_t)0), (int64_t)PE_ROWS - 1);
        pe32[e] = v;
        if (__ldcg(present + v) == 0) present[v] = 1;         // read first (through L2): see prepare_kernel
        pmin = min(pmin, v);
        pmax = max(pmax, v);
    }
    pmin = __reduce_min_sync(FULL, pmin);
    pmax = __reduce_max_sync(FULL, pmax);
    if ((threadIdx.x & 31) == 0 && pmax >= 0) { atomicMin(range, pmin); atomicMax(range + 1, pmax); }
}

static bool prepare_bitmap_enabled() {
    const char *e = getenv("CLUSTEN_PREPARE_BITMAP");                  // A/B switch, read per call: 0 = the global byte-map marks
    return !(e && e[0] == '0');
}
static int sm_count_prepare() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

}  // namespace clusten

using namespace clusten;

extern "C" int clusten_table_rank(const int64_t *pe_idx, int64_t total, int32_t *inverse, int32_t *uniq, int uniq_cap, int32_t *count,
                                  void *workspace, size_t workspace_bytes, void *stream) {
    if (total < 0 || uniq_cap <= 0) return set_error(CLUSTEN_EINVAL, "bad sizes total=%lld cap=%d", (long long)total, uniq_cap);
    if (!pe_idx || !inverse || !uniq || !count || !workspace) return set_error(CLUSTEN_EINVAL, "null pointer");
    if (workspace_bytes < clusten_prepare_workspace_bytes()) return set_error(CLUSTEN_EWORKSPACE, "workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t *present = reinterpret_cast<uint8_t *>(workspace);
    const size_t pm = (size_t)((PE_ROWS + 255) & ~255);
    int32_t *rank = reinterpret_cast<int32_t *>(present + pm);
    int *range = reinterpret_cast<int *>(reinterpret_cast<char *>(workspace) + pm + (size_t)PE_ROWS * 4);
    cudaMemsetAsync(present, 0, pm, st);
    cudaMemsetAsync(range, 0x7f, 4, st);
    cudaMemsetAsync(range + 1, 0, 4, st);
    if (total == 0) { cudaMemsetAsync(count, 0, 4, st); return check_launch("table_rank memset"); }
    const int grid = (int)std::min<int64_t>(148 * 16, (total + 255) / 256);
    mark_rows_kernel<<<grid, 256, 0, st>>>(pe_idx, total, inverse, present, range);
    rank_kernel<<<1, 1024, 0, st>>>(present, rank, uniq, count, uniq_cap, range);
    rerank_kernel<<<grid, 256, 0, st>>>(inverse, rank, inverse, total);          // in place: every thread reads its element, then writes it
    note_launches(3);
    return check_launch("table_rank");
}

extern "C" size_t clusten_prepare_workspace_bytes(void) {
    // presence map (padded) + rank table
    return (size_t)((PE_ROWS + 255) & ~255) + (size_t)PE_ROWS * 4 + 256;
}

extern "C" int clusten_stage_prepare(const int64_t *nearest, const int64_t *member, const int64_t *cluster_mask, const float *pos,
                                     int B, int n, int k, int m, int nnc,
                                     int64_t *member_idx, int64_t *mask64, uint8_t *mask8, int32_t *pe_idx, int32_t *bias_idx,
                                     int32_t *uniq, int uniq_cap, int32_t *count, void *workspace, size_t workspace_bytes,
                                     void *stream) {
    if (B < 0 || n < 0 || k <= 0 || m <= 0 || nnc <= 0 || uniq_cap <= 0)
        return set_error(CLUSTEN_EINVAL, "bad sizes B=%d n=%d k=%d m=%d nnc=%d", B, n, k, m, nnc);
    if (!nearest || !member || !pos || !member_idx || !pe_idx || !bias_idx || !uniq || !count || !workspace)
        return set_error(CLUSTEN_EINVAL, "null pointer");
    if (workspace_bytes < clusten_prepare_workspace_bytes()) return set_error(CLUSTEN_EWORKSPACE, "workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t *present = reinterpret_cast<uint8_t *>(workspace);
    const size_t pm = (size_t)((PE_ROWS + 255) & ~255);
    int32_t *rank = reinterpret_cast<int32_t *>(present + pm);
    cudaMemsetAsync(present, 0, pm, st);
    int *range = reinterpret_cast<int *>(reinterpret_cast<char *>(workspace) + pm + (size_t)PE_ROWS * 4);      // {min, max} referenced row
    cudaMemsetAsync(range, 0x7f, 4, st);
    cudaMemsetAsync(range + 1, 0, 4, st);
    const int64_t total = (int64_t)B * n * nnc * m;
    if (total == 0) { cudaMemsetAsync(count, 0, 4, st); return check_launch("prepare memset"); }
    const int64_t pthreads = (m & 3) == 0 ? total / 4 : total;
    const int pgrid = (int)std::min<int64_t>(148 * 16, (pthreads + 255) / 256);
    const int grid = (int)std::min<int64_t>(148 * 16, (total + 255) / 256);
    // (longlong4 accesses: 32-byte aligned rows -- member / mask / member_idx come from the allocator, m % 4 == 0 keeps the groups aligned)
    const bool vec_ok = (m & 3) == 0 && ((reinterpret_cast<uintptr_t>(member) | reinterpret_cast<uintptr_t>(member_idx) |
                                           reinterpret_cast<uintptr_t>(cluster_mask) | reinterpret_cast<uintptr_t>(mask64)) & 31u) == 0 &&
                        (reinterpret_cast<uintptr_t>(pe_idx) & 15u) == 0 && (reinterpret_cast<uintptr_t>(mask8) & 3u) == 0;
    if (vec_ok && prepare_bitmap_enabled()) {
        static bool attr_set = false;
        if (!attr_set) {
            if (cudaFuncSetAttribute(prepare_bm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PBM_SMEM) != cudaSuccess)
                return set_error(CLUSTEN_EUNSUPPORTED, "stage_prepare: cannot reserve %zu bytes of shared memory", PBM_SMEM);
            attr_set = true;
        }
        const int bgrid = (int)std::min<int64_t>(sm_count_prepare(), (pthreads + 2 * PBM_THREADS - 1) / (2 * PBM_THREADS));
        prepare_bm_kernel<<<bgrid, PBM_THREADS, PBM_SMEM, st>>>(nearest, member, cluster_mask, pos, B, n, k, m, nnc, member_idx, mask64, mask8,
                                                               pe_idx, present, range);
    } else {
        prepare_kernel<<<pgrid, 256, 0, st>>>(nearest, member, cluster_mask, pos, B, n, k, m, nnc, member_idx, mask64, mask8, pe_idx, present, range);
    }
    rank_kernel<<<1, 1024, 0, st>>>(present, rank, uniq, count, uniq_cap, range);
    rerank_kernel<<<grid, 256, 0, st>>>(pe_idx, rank, bias_idx, total);
    note_launches(3);
    return check_launch("stage_prepare");
}
