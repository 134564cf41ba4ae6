// Shared device helpers for libclusten_b200 (sm_100a).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/clusten_b200.h"

namespace clusten {

int set_error(int code, const char *fmt, ...);
int check_launch(const char *what);
void note_launches(int n);   // bumps the counter behind clusten_kernel_launches()

constexpr int WARPS_PER_CTA = 8;
constexpr int CTA_THREADS = WARPS_PER_CTA * 32;
constexpr unsigned FULL = 0xffffffffu;

template <typename T> struct Vec { static constexpr int VPT = 16 / sizeof(T); };   // elements per 128-bit access

__device__ __forceinline__ float to_f(float x) { return x; }
__device__ __forceinline__ float to_f(__half x) { return __half2float(x); }
__device__ __forceinline__ float to_f(__nv_bfloat16 x) { return __bfloat162float(x); }
template <typename T> __device__ __forceinline__ T from_f(float x);
template <> __device__ __forceinline__ float from_f<float>(float x) { return x; }
template <> __device__ __forceinline__ __half from_f<__half>(float x) { return __float2half_rn(x); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float x) { return __float2bfloat16_rn(x); }

// 128-bit read-only load of VPT elements, widened to fp32.  p must be 16-byte aligned.
__device__ __forceinline__ void load16(const float *p, float (&o)[4]) {
    const float4 v = __ldg(reinterpret_cast<const float4 *>(p));
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}
__device__ __forceinline__ void load16(const __half *p, float (&o)[8]) {
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p));
    const __half2 *h = reinterpret_cast<const __half2 *>(&v);
#pragma unroll
    for (int t = 0; t < 4; ++t) { const float2 f = __half22float2(h[t]); o[2 * t] = f.x; o[2 * t + 1] = f.y; }
}
__device__ __forceinline__ void load16(const __nv_bfloat16 *p, float (&o)[8]) {
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p));
    const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&v);
#pragma unroll
    for (int t = 0; t < 4; ++t) { const float2 f = __bfloat1622float2(h[t]); o[2 * t] = f.x; o[2 * t + 1] = f.y; }
}
__device__ __forceinline__ void store16(float *p, const float (&o)[4]) {
    *reinterpret_cast<float4 *>(p) = make_float4(o[0], o[1], o[2], o[3]);
}
__device__ __forceinline__ void store16(__half *p, const float (&o)[8]) {
    uint4 v;
    __half2 *h = reinterpret_cast<__half2 *>(&v);
#pragma unroll
    for (int t = 0; t < 4; ++t) h[t] = __floats2half2_rn(o[2 * t], o[2 * t + 1]);
    *reinterpret_cast<uint4 *>(p) = v;
}
__device__ __forceinline__ void store16(__nv_bfloat16 *p, const float (&o)[8]) {
    uint4 v;
    __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&v);
#pragma unroll
    for (int t = 0; t < 4; ++t) h[t] = __floats2bfloat162_rn(o[2 * t], o[2 * t + 1]);
    *reinterpret_cast<uint4 *>(p) = v;
}

// sum over the G lanes of an aligned lane group (G power of two); every lane of the group gets the sum
template <int G> __device__ __forceinline__ float group_sum(float s) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
    return s;
}
// sum across the 32/G lane groups (lanes with equal lane%G); every lane gets the sum
template <int G> __device__ __forceinline__ float cross_group_sum(float s) {
#pragma unroll
    for (int o = G; o < 32; o <<= 1) s += __shfl_xor_sync(FULL, s, o);
    return s;
}

inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// smallest supported lane-group size covering nchunk 16-byte chunks per row (<= 32 chunks)
inline int pick_group(int nchunk) { return nchunk <= 4 ? 4 : nchunk <= 8 ? 8 : nchunk <= 16 ? 16 : 32; }

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- segmented radix sort (sort.cu) ----
size_t radix_sort_workspace_bytes(int B, int n);          // histogram scratch only
// Stable LSD radix sort of B segments of n (key,val) pairs over key bits [0, key_bits).
// vals_in == nullptr -> implicit iota (val = position within segment).  Result lands in (keys_out, vals_out);
// (keys_tmp, vals_tmp) and (keys_in) are scratch / clobbered.  All buffers B*n uint32.
// skip_if_zero != nullptr: every kernel of the sort exits at once when skip_if_zero[0] == 0 (device-side dispatch).
int radix_sort_pairs(uint32_t *keys_in, const uint32_t *vals_in, uint32_t *keys_tmp, uint32_t *vals_tmp,
                     uint32_t *keys_out, uint32_t *vals_out, int B, int n, int key_bits,
                     void *hist_ws, cudaStream_t stream, const int *skip_if_zero = nullptr);

}  // namespace clusten

#define CLUSTEN_DISPATCH_DTYPE(dtype, ...)                                                  \
    switch (dtype) {                                                                        \
        case CLUSTEN_F32: { using T = float; __VA_ARGS__; break; }                          \
        case CLUSTEN_F16: { using T = __half; __VA_ARGS__; break; }                         \
        case CLUSTEN_BF16: { using T = __nv_bfloat16; __VA_ARGS__; break; }                 \
        default: return clusten::set_error(CLUSTEN_EDTYPE, "unknown dtype %d", dtype);      \
    }

#define CLUSTEN_DISPATCH_GROUP(G_, ...)                                                     \
    switch (G_) {                                                                           \
        case 4: { constexpr int G = 4; __VA_ARGS__; break; }                                \
        case 8: { constexpr int G = 8; __VA_ARGS__; break; }                                \
        case 16: { constexpr int G = 16; __VA_ARGS__; break; }                              \
        default: { constexpr int G = 32; __VA_ARGS__; break; }                              \
    }
