// Segmented, stable LSD radix sort of (uint32 key, uint32 value) pairs -- the one sorting primitive behind the
// inverse neighbour list (clusten_csr_build), the space-filling-curve ordering (clusten_sfc_cluster) and the canonical
// top-k (clusten_topk_select).  B independent segments of n pairs; 8 bits per pass; three kernels per pass:
//   hist    : per-tile digit histogram               grid (tiles, B)
//   scan    : exclusive scan over (digit, tile)      grid (B)
//   scatter : stable in-tile ranking + scatter       grid (tiles, B)
// Stability (ties keep their input order) is what makes every consumer deterministic and gives the canonical
// "ties -> lower index" rule of the oracle; the in-tile ranking uses warp match_any so no atomics order the data.
#include "common.cuh"

namespace clusten {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_IPT = 8;                              // items per thread
constexpr int RS_TILE = RS_THREADS * RS_IPT;           // 2048 items per CTA
constexpr int RS_RADIX = 256;

static inline int rs_tiles(int n) { return (n + RS_TILE - 1) / RS_TILE; }

size_t radix_sort_workspace_bytes(int B, int n) {
    return (size_t)B * RS_RADIX * rs_tiles(n) * sizeof(int) + 256;
}

__global__ void __launch_bounds__(RS_THREADS)
rs_hist_kernel(const uint32_t *__restrict__ keys, int n, int shift, int *__restrict__ hist, int T, const int *skip) {
    if (skip && skip[0] == 0) return;
    __shared__ int h[RS_RADIX];
    const int t = blockIdx.x, b = blockIdx.y;
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t *kin = keys + (int64_t)b * n;
    const int base = t * RS_TILE;
#pragma unroll
    for (int s = 0; s < RS_IPT; ++s) {
        const int p = base + s * RS_THREADS + threadIdx.x;
        if (p < n) atomicAdd(&h[(kin[p] >> shift) & 255u], 1);
    }
    __syncthreads();
    hist[((int64_t)b * RS_RADIX + threadIdx.x) * T + t] = h[threadIdx.x];
}

// exclusive scan of L = 256*T ints per segment, in place; one CTA of 1024 threads per segment
__global__ void __launch_bounds__(1024) rs_scan_kernel(int *__restrict__ hist, int L, const int *skip) {
    if (skip && skip[0] == 0) return;
    __shared__ int warp_tot[32];
    __shared__ int carry_s;
    int *a = hist + (int64_t)blockIdx.x * L;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < L; base += 1024) {
        const int p = base + threadIdx.x;
        const int v = p < L ? a[p] : 0;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(FULL, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) warp_tot[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            int w = warp_tot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(FULL, w, o);
                if (lane >= o) w += t;
            }
            warp_tot[lane] = w;                      // inclusive scan of warp totals
        }
        __syncthreads();
        const int carry = carry_s;
        const int excl = carry + (warp ? warp_tot[warp - 1] : 0) + inc - v;
        if (p < L) a[p] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + warp_tot[31];
        __syncthreads();
    }
}

__global__ void __launch_bounds__(RS_THREADS)
rs_scatter_kernel(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                  uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out,
                  int n, int shift, const int *__restrict__ hist, int T, const int *skip) {
    if (skip && skip[0] == 0) return;
    __shared__ int cnt[RS_WARPS][RS_RADIX + 1];
    __shared__ int base_s[RS_WARPS][RS_RADIX];
    const int t = blockIdx.x, b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int x = threadIdx.x; x < RS_WARPS * (RS_RADIX + 1); x += RS_THREADS) (&cnt[0][0])[x] = 0;
    __syncthreads();
    const uint32_t *kin = keys_in + (int64_t)b * n;
    const uint32_t *vin = vals_in ? vals_in + (int64_t)b * n : nullptr;
    // warp w owns the contiguous items [w*256, (w+1)*256) of the tile, 32 at a time -> input order is preserved
    const int wbase = t * RS_TILE + warp * (32 * RS_IPT);
    uint32_t key[RS_IPT];
    int rnk[RS_IPT];
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll
    for (int s = 0; s < RS_IPT; ++s) {
        const int p = wbase + s * 32 + lane;
        const bool valid = p < n;
        key[s] = valid ? kin[p] : 0u;
        const int d = valid ? (int)((key[s] >> shift) & 255u) : RS_RADIX;
        const unsigned peers = __match_any_sync(FULL, d);
        const int c = cnt[warp][d];
        __syncwarp();
        if (lane == __ffs(peers) - 1) cnt[warp][d] = c + __popc(peers);
        __syncwarp();
        rnk[s] = c + __popc(peers & lt);
    }
    __syncthreads();
    {
        const int d = threadIdx.x;                    // RS_THREADS == RS_RADIX
        int run = hist[((int64_t)b * RS_RADIX + d) * T + t];
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) {
            base_s[w][d] = run;
            run += cnt[w][d];
        }
    }
    __syncthreads();
    uint32_t *kout = keys_out + (int64_t)b * n;
    uint32_t *vout = vals_out + (int64_t)b * n;
#pragma unroll
    for (int s = 0; s < RS_IPT; ++s) {
        const int p = wbase + s * 32 + lane;
        if (p < n) {
            const int d = (int)((key[s] >> shift) & 255u);
            const int dst = base_s[warp][d] + rnk[s];
            kout[dst] = key[s];
            vout[dst] = vin ? vin[p] : (uint32_t)p;
        }
    }
}

// ---- short segments: every pass inside ONE kernel -------------------------------------------------------------------------------
// n <= 4096 pairs per segment (every sort of stages 1-3 at 512 x 512: 4096 / 1024 / 256 tokens per image) fit in shared memory with
// both ping-pong buffers, so the 3 kernels x 4 passes of the tiled path (launch- and dependency-bound at this size: ~18 us per pass
// for a few KB of data) become one launch of one CTA per segment.  Same stable ranking as rs_scatter_kernel: a warp owns a contiguous
// run of items and ranks them 32 at a time with match_any, so ties keep their input order.
constexpr int SS_THREADS = 512;
constexpr int SS_WARPS = SS_THREADS / 32;
constexpr int SS_IPT = 8;
constexpr int SS_CAP = SS_THREADS * SS_IPT;            // 4096 items
constexpr size_t SS_SMEM = (size_t)4 * SS_CAP * 4 + (size_t)SS_WARPS * (RS_RADIX + 1) * 4 + 64;

__global__ void __launch_bounds__(SS_THREADS)
rs_small_kernel(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in, uint32_t *__restrict__ keys_out,
                uint32_t *__restrict__ vals_out, int n, int passes, const int *skip) {
    if (skip && skip[0] == 0) return;
    extern __shared__ uint32_t ss_smem[];
    uint32_t *ka = ss_smem, *va = ka + SS_CAP, *kb = va + SS_CAP, *vb = kb + SS_CAP;
    int (*cnt)[RS_RADIX + 1] = reinterpret_cast<int (*)[RS_RADIX + 1]>(vb + SS_CAP);
    int *warp_tot = reinterpret_cast<int *>(cnt + SS_WARPS);
    const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t *kin = keys_in + (int64_t)b * n;
    const uint32_t *vin = vals_in ? vals_in + (int64_t)b * n : nullptr;
    for (int i = threadIdx.x; i < n; i += SS_THREADS) {
        ka[i] = kin[i];
        va[i] = vin ? vin[i] : (uint32_t)i;
    }
    const unsigned lt = (1u << lane) - 1u;
    for (int p = 0; p < passes; ++p) {
        const int shift = 8 * p;
        for (int x = threadIdx.x; x < SS_WARPS * (RS_RADIX + 1); x += SS_THREADS) (&cnt[0][0])[x] = 0;
        __syncthreads();
        uint32_t key[SS_IPT], val[SS_IPT];
        int rnk[SS_IPT];
        const int wbase = warp * (32 * SS_IPT);
#pragma unroll
        for (int s = 0; s < SS_IPT; ++s) {
            const int q = wbase + s * 32 + lane;
            const bool valid = q < n;
            key[s] = valid ? ka[q] : 0u;
            val[s] = valid ? va[q] : 0u;
            const int d = valid ? (int)((key[s] >> shift) & 255u) : RS_RADIX;
            const unsigned peers = __match_any_sync(FULL, d);
            const int c = cnt[warp][d];
            __syncwarp();
            if (lane == __ffs(peers) - 1) cnt[warp][d] = c + __popc(peers);
            __syncwarp();
            rnk[s] = c + __popc(peers & lt);
        }
        __syncthreads();
        // digit d: counts of the 16 warps -> start of (warp, digit) in the output; exclusive scan of the digit totals over 256 threads
        int tot = 0;
        if (threadIdx.x < RS_RADIX) {
#pragma unroll
            for (int w = 0; w < SS_WARPS; ++w) tot += cnt[w][threadIdx.x];
        }
        int inc = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(FULL, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) warp_tot[warp] = inc;
        __syncthreads();
        if (threadIdx.x < RS_RADIX) {
            int run = inc - tot;
            for (int w = 0; w < warp; ++w) run += warp_tot[w];           // warps 0..7 hold the 256 digits
#pragma unroll
            for (int w = 0; w < SS_WARPS; ++w) {
                const int c = cnt[w][threadIdx.x];
                cnt[w][threadIdx.x] = run;
                run += c;
            }
        }
        __syncthreads();
#pragma unroll
        for (int s = 0; s < SS_IPT; ++s) {
            const int q = wbase + s * 32 + lane;
            if (q < n) {
                const int dst = cnt[warp][(key[s] >> shift) & 255u] + rnk[s];
                kb[dst] = key[s];
                vb[dst] = val[s];
            }
        }
        __syncthreads();
        uint32_t *t0 = ka; ka = kb; kb = t0;
        uint32_t *t1 = va; va = vb; vb = t1;
    }
    uint32_t *kout = keys_out + (int64_t)b * n;
    uint32_t *vout = vals_out + (int64_t)b * n;
    for (int i = threadIdx.x; i < n; i += SS_THREADS) {
        kout[i] = ka[i];
        vout[i] = va[i];
    }
}

int radix_sort_pairs(uint32_t *keys_in, const uint32_t *vals_in, uint32_t *keys_tmp, uint32_t *vals_tmp,
                     uint32_t *keys_out, uint32_t *vals_out, int B, int n, int key_bits,
                     void *hist_ws, cudaStream_t st, const int *skip) {
    if (B <= 0 || n <= 0) return 0;
    const int passes = key_bits <= 0 ? 1 : (key_bits + 7) / 8;
    if (n <= SS_CAP) {
        static bool attr_set = false;
        if (!attr_set) {
            cudaFuncSetAttribute(rs_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SS_SMEM);
            attr_set = true;
        }
        rs_small_kernel<<<B, SS_THREADS, SS_SMEM, st>>>(keys_in, vals_in, keys_out, vals_out, n, passes, skip);
        note_launches(1);
        return check_launch("radix_sort");
    }
    const int T = rs_tiles(n);
    int *hist = reinterpret_cast<int *>(hist_ws);
    const dim3 grid(T, B);
    const uint32_t *kin = keys_in;
    const uint32_t *vin = vals_in;
    for (int p = 0; p < passes; ++p) {
        // ping-pong so that the LAST pass writes (keys_out, vals_out)
        const bool to_out = ((passes - 1 - p) % 2) == 0;
        uint32_t *ko = to_out ? keys_out : keys_tmp;
        uint32_t *vo = to_out ? vals_out : vals_tmp;
        rs_hist_kernel<<<grid, RS_THREADS, 0, st>>>(kin, n, 8 * p, hist, T, skip);
        rs_scan_kernel<<<B, 1024, 0, st>>>(hist, RS_RADIX * T, skip);
        rs_scatter_kernel<<<grid, RS_THREADS, 0, st>>>(kin, vin, ko, vo, n, 8 * p, hist, T, skip);
        note_launches(3);
        kin = ko;
        vin = vo;
    }
    return check_launch("radix_sort");
}

}  // namespace clusten
