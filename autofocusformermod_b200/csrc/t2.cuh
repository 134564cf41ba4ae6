// Device helpers shared by the second-generation tile kernels (clusten_tile2.cu, clusten_fused_bwd.cu).
#pragma once
#include "tile.cuh"

namespace clusten {
namespace t2 {

constexpr int TW2 = 8;                   // warps per CTA

__device__ __forceinline__ void mma_bf16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_f16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
template <typename T> __device__ __forceinline__ void mma16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    if constexpr (std::is_same<T, __half>::value) mma_f16(d, a0, a1, a2, a3, b0, b1);
    else mma_bf16(d, a0, a1, a2, a3, b0, b1);
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void split_tf32(uint32_t x, uint32_t &hi, uint32_t &lo) {
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(__uint_as_float(x)));
    const float r = __uint_as_float(x) - __uint_as_float(hi);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}
template <typename T> __device__ __forceinline__ uint32_t pack_pair(float a, float b) {
    if constexpr (std::is_same<T, __half>::value) { const __half2 v = __floats2half2_rn(a, b); return *reinterpret_cast<const uint32_t *>(&v); }
    else { const __nv_bfloat162 v = __floats2bfloat162_rn(a, b); return *reinterpret_cast<const uint32_t *>(&v); }
}
template <typename T> __device__ __forceinline__ void st_pair(T *p, float a, float b) {
    if constexpr (sizeof(T) == 4) *reinterpret_cast<float2 *>(p) = make_float2(a, b);
    else *reinterpret_cast<uint32_t *>(p) = pack_pair<T>(a, b);
}
// NB bytes (8, 16 or 32) of one row chunk into 32-bit registers (ld.global.nc spelled out: the pointers below are made
// opaque to the optimiser, which would otherwise demote __ldg to generic loads)
__device__ __forceinline__ uint4 ldg16(const void *p) {
    uint4 v;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint2 ldg8(const void *p) {
    uint2 v;
    asm volatile("ld.global.nc.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t ldg4(const void *p) {
    uint32_t v;
    asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
template <int NB> __device__ __forceinline__ void ld_chunk(uint32_t (&r)[NB / 4], const void *p) {
    if constexpr (NB == 8) {
        const uint2 v = ldg8(p);
        r[0] = v.x; r[1] = v.y;
    } else {
#pragma unroll
        for (int x = 0; x < NB / 16; ++x) {
            const uint4 v = ldg16(reinterpret_cast<const uint4 *>(p) + x);
            r[4 * x] = v.x; r[4 * x + 1] = v.y; r[4 * x + 2] = v.z; r[4 * x + 3] = v.w;
        }
    }
}
__device__ __forceinline__ int sbyte(uint32_t w, int j) { return (int)(int8_t)(w >> (8 * j)); }
// base + off elements with ONE 32x32->64 multiply-add (IMAD.WIDE) instead of a 64-bit add chain
template <typename T> __device__ __forceinline__ T *at(T *base, int off) { return base + off; }
// hide a per-warp base pointer from the optimiser: otherwise it re-associates base + offset into 64-bit add chains
template <typename T> __device__ __forceinline__ T *opaque(T *p) {
    asm volatile("" : "+l"(p));
    return p;
}
// predicated 32-bit / 64-bit global stores (kept as predicates: the compiler otherwise emits a branch per store)
__device__ __forceinline__ void st32_if(void *p, uint32_t v, int s) {
    asm volatile("{ .reg .pred p; setp.ge.s32 p, %2, 0; @p st.global.b32 [%0], %1; }" ::"l"(p), "r"(v), "r"(s) : "memory");
}
__device__ __forceinline__ void st64_if(void *p, float a, float b, int s) {
    asm volatile("{ .reg .pred p; setp.ge.s32 p, %3, 0; @p st.global.v2.f32 [%0], {%1, %2}; }" ::"l"(p), "f"(a), "f"(b), "r"(s) : "memory");
}
template <typename T> __device__ __forceinline__ void st_pair_if(T *p, float a, float b, int s) {
    if constexpr (sizeof(T) == 4) st64_if(p, a, b, s);
    else st32_if(p, pack_pair<T>(a, b), s);
}

__device__ __forceinline__ void cp16(uint32_t smem, const void *gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem), "l"(gmem));
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }
__device__ __forceinline__ void ldsm4t(uint32_t (&r)[4], uint32_t s) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(s));
}
__device__ __forceinline__ uint32_t lds32(uint32_t s) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(s));
    return v;
}
__device__ __forceinline__ uint32_t lds16(uint32_t s) {
    uint32_t v;
    asm volatile("{ .reg .u16 t; ld.shared.u16 t, [%1]; cvt.u32.u16 %0, t; }" : "=r"(v) : "r"(s));
    return v;
}
// two consecutive 16-bit elements at byte address s (2-byte aligned) of shared memory
template <bool AL4> __device__ __forceinline__ uint32_t lds_pair(uint32_t s) {
    if constexpr (AL4) return lds32(s);
    else return lds16(s) | (lds16(s + 2) << 16);
}
__device__ __forceinline__ void cp16_zfill(uint32_t smem, const void *gmem, bool pred) {
    const int sz = pred ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem), "l"(gmem), "r"(sz));
}

// ---- impure tokens (tile.cuh): left out of the tile structure, listed in pk.imp_list and computed one (token, head) per warp
// by whichever warps of the grid finish first -- see slow_items().  16-bit types, C % 8 == 0, C <= 32.
template <typename T> __device__ __forceinline__ float dot8(const uint4 &a, const uint4 &b) {
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float2 fa, fb;
        if constexpr (std::is_same<T, __half>::value) {
            fa = __half22float2(*reinterpret_cast<const __half2 *>(&aw[q]));
            fb = __half22float2(*reinterpret_cast<const __half2 *>(&bw[q]));
        } else {
            fa = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&aw[q]));
            fb = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&bw[q]));
        }
        s = fmaf(fa.x, fb.x, s);
        s = fmaf(fa.y, fb.y, s);
    }
    return s;
}
template <typename T> __device__ __forceinline__ void unpack8(const uint4 &a, float (&f)[8]) {
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float2 fa;
        if constexpr (std::is_same<T, __half>::value) fa = __half22float2(*reinterpret_cast<const __half2 *>(&aw[q]));
        else fa = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&aw[q]));
        f[2 * q] = fa.x; f[2 * q + 1] = fa.y;
    }
}
// out[j] = x . y[idx[j]]: one neighbour per lane, rows read as 16-byte chunks
template <typename T>
__device__ __noinline__ void dot_row_fast(const T *xrow, const T *ybase, int y_sn, const int64_t *irow, T *orow, int C, int M, int lane) {
    const int nch = C >> 3;
    uint4 xq[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) xq[c] = c < nch ? ldg16(xrow + 8 * c) : make_uint4(0u, 0u, 0u, 0u);
    for (int j = lane; j < M; j += 32) {
        const T *y = ybase + (int)__ldg(irow + j) * y_sn;
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c)
            if (c < nch) s += dot8<T>(xq[c], ldg16(y + 8 * c));
        orow[j] = from_f<T>(s);
    }
}
// out[:] = sum_j w[j] y[idx[j]][:]: lane = (neighbour group lane >> 2, 8-channel block lane & 3), shuffle-reduced over the groups
template <typename T>
__device__ __noinline__ void axpy_row_fast(const T *wrow, const T *ybase, int y_sn, const int64_t *irow, T *orow, int C, int M, int lane) {
    const int jg = lane >> 2, cb = lane & 3;
    const bool act = 8 * cb < C;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
    for (int j = jg; j < M; j += 8) {
        const float w = to_f(wrow[j]);
        const uint4 v = ldg16(ybase + (int)__ldg(irow + j) * y_sn + (act ? 8 * cb : 0));
        float f[8];
        unpack8<T>(v, f);
#pragma unroll
        for (int x = 0; x < 8; ++x) acc[x] = fmaf(w, f[x], acc[x]);
    }
#pragma unroll
    for (int x = 0; x < 8; ++x) {
        acc[x] += __shfl_xor_sync(FULL, acc[x], 4);
        acc[x] += __shfl_xor_sync(FULL, acc[x], 8);
        acc[x] += __shfl_xor_sync(FULL, acc[x], 16);
    }
    if (jg == 0 && act) {
        uint4 o;
        o.x = pack_pair<T>(acc[0], acc[1]); o.y = pack_pair<T>(acc[2], acc[3]);
        o.z = pack_pair<T>(acc[4], acc[5]); o.w = pack_pair<T>(acc[6], acc[7]);
        *reinterpret_cast<uint4 *>(orow + 8 * cb) = o;
    }
}
// number of (impure token, head) items and this warp's first item / stride over the whole grid
struct SlowIter { int n, first, stride; };
__device__ __forceinline__ SlowIter slow_items(const PackView &pk, int H) {
    SlowIter it;
    it.n = min(pk.flags[2], pk.imp_cap) * H;
    const int wpc = blockDim.x >> 5;
    it.first = (blockIdx.y * gridDim.x + blockIdx.x) * wpc + (threadIdx.x >> 5);
    it.stride = gridDim.x * gridDim.y * wpc;
    return it;
}

}  // namespace t2
}  // namespace clusten
