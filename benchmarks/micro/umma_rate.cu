// Issue-rate micro-benchmark for tcgen05.mma kind::tf32 on one SM per CTA: cycles per MMA (M = 128, K = 8) for N in {64, 128, 256},
// A from shared memory (SS) or from tensor memory (TS), accumulating into one accumulator or alternating between two.
// The operands are whatever is in shared / tensor memory (zeros): only the timing matters.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o umma_rate umma_rate.cu && ./umma_rate
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3fffu) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
    return pred != 0;
}

template <int N, bool TS, int ND>
__global__ void __launch_bounds__(128, 1) rate_kernel(int iters, long long *cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) uint64_t bar;
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    for (int i = threadIdx.x; i < (16384 + N * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem + (base - smem_u32(smem)))[i] = 0;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (warp == 1) {
        long long t0 = 0, t1 = 0;
        if (elect_one()) {
            const uint64_t adesc = umma_desc(base), bdesc = umma_desc(base + 16384);
            const uint32_t a_tmem = tmem + (ND * N < 448 ? 448 : 480);
            t0 = clock64();
            for (int i = 0; i < iters; ++i) {
#pragma unroll
                for (int k = 0; k < 12; ++k) {
                    const uint32_t d = tmem + ((k % ND) * N);
                    if (TS)
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
                                     ::"r"(d), "r"(a_tmem + 8 * (k & 3)), "l"(bdesc + 2 * (k & 3)), "r"(IDESC), "r"(1u) : "memory");
                    else
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                                     ::"r"(d), "l"(adesc + 2 * (k & 3)), "l"(bdesc + 2 * (k & 3)), "r"(IDESC), "r"(1u) : "memory");
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            uint32_t ok = 0;
            while (!ok)
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                             : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
            t1 = clock64();
            if (blockIdx.x == 0) *cycles = t1 - t0;
        }
        __syncwarp();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

template <int N, bool TS, int ND> void run(const char *name) {
    long long *d, h = 0;
    cudaMalloc(&d, 8);
    const int smem = 16384 + N * 128 + 1024, iters = 400;
    cudaFuncSetAttribute(rate_kernel<N, TS, ND>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int rep = 0; rep < 2; ++rep) {
        rate_kernel<N, TS, ND><<<148, 128, smem>>>(iters, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    }
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    const double per = (double)h / (iters * 12.0);
    printf("%-28s %7.1f cycles / MMA   -> %6.0f TFLOP/s (TF32, 148 SMs @ 1.9 GHz)\n", name, per, 2.0 * 128 * N * 8 / per * 148 * 1.9e9 / 1e12);
    cudaFree(d);
}

int main() {
    run<64, false, 1>("SS N=64  one accumulator");
    run<128, false, 1>("SS N=128 one accumulator");
    run<256, false, 1>("SS N=256 one accumulator");
    run<128, false, 2>("SS N=128 two accumulators");
    run<64, true, 1>("TS N=64  one accumulator");
    run<128, true, 1>("TS N=128 one accumulator");
    run<256, true, 1>("TS N=256 one accumulator");
    run<128, true, 2>("TS N=128 two accumulators");
    return 0;
}
