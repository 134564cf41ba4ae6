// Tile-union tensor-core kernels for CLUSTEN QK / AV (forward and backward), sm_100a.
//
// Why: the row-gather kernels of clusten_attn.cu are instruction-issue bound (ncu: 1500 warp instructions per token,
// DRAM 8 % busy) because every key row is re-fetched and re-multiplied by each of the ~48 tokens that reference it.
// Here one warp owns a tile of 16 consecutive tokens and ONE head, walks the union of key octets the tile references
// (tile.cuh; ~13 octets for the reference's curve-ordered clusters) and computes the dense 16 x 8 block against every
// octet with warp-level tensor-core MMAs; each token then keeps only the blocks of the octets it references:
//   dot    out[i, 8s+r]  = X[i,:] . Y[8o+r,:]                 QK fwd (X=q, Y=k);  d_attn of AV bwd (X=d_feat, Y=v)
//   axpy   out[i,:]      = sum_{s,r} W[i, 8s+r] * Y[8o+r,:]   AV fwd (W=attn, Y=v); d_q of QK bwd (W=d_attn, Y=k)
//   scat   out[8o+r,:]   = sum_{i} W[i, 8s+r] * X[i,:]        d_k of QK bwd (W=d_attn, X=q); d_v of AV bwd (W=attn, X=d_feat)
// with o = octet referenced by token i at slot s.  16-bit types use mma.sync.m16n8k16 (fp32 accumulate); fp32 uses
// 3xTF32 (m16n8k8 on the hi/lo split of both operands: a_lo*b_hi + a_hi*b_lo + a_hi*b_hi), which keeps fp32-level
// accuracy (measured <= 2e-6 relative).  Operand fragments come straight from 128-bit row loads: lane (g, t) of the mma
// layout loads channels [CH*t, CH*t+CH) of row g, and the SAME channel->k permutation is used for both operands, so no
// shuffle or shared-memory transpose is needed for the dot shape.  Every instruction does 2048 (bf16) MACs instead
// of 32: ~20 instructions per token and head instead of ~500.
#include "tile.cuh"

namespace clusten {

constexpr int TW = 8;                    // warps per CTA

__device__ __forceinline__ void mma16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                         uint32_t b0, uint32_t b1, __nv_bfloat16) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                         uint32_t b0, uint32_t b1, __half) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma1688(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                        uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t tf32_hi(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
// x = hi + lo with hi = x truncated to tf32 (one LOP3) and lo = x - hi (exact; the tensor cores read its upper 19 bits): two
// instructions where `cvt.rna.tf32.f32` costs seven twice over on sm_100 (the rounding is emulated).  hi + lo keeps 21 mantissa bits.
__device__ __forceinline__ void tf32_split(float x, uint32_t &hi, uint32_t &lo) {
    hi = __float_as_uint(x) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}
// 3xTF32: d += a*b with fp32-level accuracy
__device__ __forceinline__ void mma_3xtf32(float (&d)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4],
                                           uint32_t b0h, uint32_t b1h, uint32_t b0l, uint32_t b1l) {
    mma1688(d, al[0], al[1], al[2], al[3], b0h, b1h);
    mma1688(d, ah[0], ah[1], ah[2], ah[3], b0l, b1l);
    mma1688(d, ah[0], ah[1], ah[2], ah[3], b0h, b1h);
}

template <typename T> __device__ __forceinline__ void store_pair(T *p, float a, float b);
template <> __device__ __forceinline__ void store_pair<float>(float *p, float a, float b) {
    *reinterpret_cast<float2 *>(p) = make_float2(a, b);
}
template <> __device__ __forceinline__ void store_pair<__half>(__half *p, float a, float b) {
    *reinterpret_cast<__half2 *>(p) = __floats2half2_rn(a, b);
}
template <> __device__ __forceinline__ void store_pair<__nv_bfloat16>(__nv_bfloat16 *p, float a, float b) {
    *reinterpret_cast<__nv_bfloat162 *>(p) = __floats2bfloat162_rn(a, b);
}

// CH channels of one row as 32-bit registers: 16-bit types CH/2 regs, fp32 CH regs
template <typename T, int CH> struct RowFrag { uint32_t r[CH * sizeof(T) / 4]; };

template <typename T, int CH> __device__ __forceinline__ void load_frag(RowFrag<T, CH> &f, const T *p, bool pred) {
    constexpr int NB = CH * sizeof(T);                 // 8, 16 or 32 bytes
#pragma unroll
    for (int x = 0; x < NB / 4; ++x) f.r[x] = 0u;
    if (!pred) return;
    if constexpr (NB == 8) {
        const uint2 v = __ldg(reinterpret_cast<const uint2 *>(p));
        f.r[0] = v.x; f.r[1] = v.y;
    } else {
#pragma unroll
        for (int x = 0; x < NB / 16; ++x) {
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p) + x);
            f.r[4 * x] = v.x; f.r[4 * x + 1] = v.y; f.r[4 * x + 2] = v.z; f.r[4 * x + 3] = v.w;
        }
    }
}

struct TileCtx {
    int b, h, bt, i0, U, g, t, lane;
    int o0, o1;
};

__device__ __forceinline__ bool tile_ctx(TileCtx &c, const PackView &pk, int B, int H) {
    const int lane = threadIdx.x & 31;
    const int64_t item = (int64_t)blockIdx.x * TW + (threadIdx.x >> 5);
    if (item >= (int64_t)B * pk.T * H) return false;
    c.lane = lane;
    c.h = (int)(item % H);
    c.bt = (int)(item / H);
    c.b = c.bt / pk.T;
    c.i0 = (c.bt - c.b * pk.T) * TILE_TOK;
    c.g = lane >> 2;
    c.t = lane & 3;
    c.U = pk.tile_u[c.bt];
    const int *octp = pk.tile_oct + (int64_t)c.bt * U_MAX;
    c.o0 = octp[lane];
    c.o1 = lane + 32 < U_MAX ? octp[lane + 32] : 0;
    return true;
}
__device__ __forceinline__ int tile_octet(const TileCtx &c, int u) {        // u is warp-uniform
    return __shfl_sync(FULL, u < 32 ? c.o0 : c.o1, u & 31);
}

// ---- slow in-kernel paths for impure tokens (rare: the padded last cluster of a stage, point_utils.py:282-283) -------------
// All warp-cooperative and warp-uniform; they read the int64 index tensor the interface delivered.
__device__ __forceinline__ uint32_t tile_imp_mask(const PackView &pk, int bt) {          // bit r = token row r is impure
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(pk.tok_imp + (int64_t)bt * TILE_TOK));
    uint32_t m = 0;
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int k = 0; k < 4; ++k) m |= ((w[q] >> (8 * k)) & 1u) << (4 * q + k);
    return m;
}
template <typename T>
__device__ __forceinline__ void slow_dot_row(const T *xrow, const T *ybase, int64_t y_sn, const int64_t *irow, T *orow,
                                             int C, int M, int lane) {
    for (int j = lane; j < M; j += 32) {
        const T *y = ybase + irow[j] * y_sn;
        float s = 0.f;
        for (int ch = 0; ch < C; ++ch) s = fmaf(to_f(xrow[ch]), to_f(y[ch]), s);
        orow[j] = from_f<T>(s);
    }
}
template <typename T>
__device__ __forceinline__ void slow_axpy_row(const T *wrow, const T *ybase, int64_t y_sn, const int64_t *irow, T *orow,
                                              int C, int M, int lane) {
    for (int ch = lane; ch < C; ch += 32) {
        float s = 0.f;
        for (int j = 0; j < M; ++j) s = fmaf(to_f(wrow[j]), to_f(ybase[irow[j] * y_sn + ch]), s);
        orow[ch] = from_f<T>(s);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// dot: 16-bit types
template <typename T, int CH>
__global__ void __launch_bounds__(TW * 32)
dot_tile16_kernel(const T *__restrict__ X, const T *__restrict__ Y, const int64_t *__restrict__ idx, const PackView pk,
                  T *__restrict__ out, int B, int H, int Nq, int C, int M,
                  int64_t x_sb, int64_t x_sh, int64_t x_sn, int64_t y_sb, int64_t y_sh, int64_t y_sn) {
    if (pk.flags[0]) return;
    TileCtx c;
    if (!tile_ctx(c, pk, B, H)) return;
    constexpr int UB = 4;                               // octets in flight
    const bool cact = CH * c.t < C;
    const int ra = c.i0 + c.g, rb = ra + 8;
    RowFrag<T, CH> xa, xb;
    const T *xbase = X + c.b * x_sb + c.h * x_sh + CH * c.t;
    load_frag<T, CH>(xa, xbase + (int64_t)ra * x_sn, cact && ra < Nq);
    load_frag<T, CH>(xb, xbase + (int64_t)rb * x_sn, cact && rb < Nq);
    const int8_t *sa = pk.slot_of + ((int64_t)c.bt * TILE_TOK + c.g) * U_MAX;
    const int8_t *sb = sa + 8 * U_MAX;
    const T *ybase = Y + c.b * y_sb + c.h * y_sh + CH * c.t + (int64_t)c.g * y_sn;
    T *oa = out + (((int64_t)c.b * H + c.h) * Nq + ra) * M + 2 * c.t;
    T *ob = oa + (int64_t)8 * M;
    for (int u0 = 0; u0 < c.U; u0 += UB) {
        RowFrag<T, CH> y[UB];
        int sga[UB], sgb[UB];
#pragma unroll
        for (int j = 0; j < UB; ++j) {
            const int u = u0 + j;
            const int o = tile_octet(c, u < c.U ? u : 0);
            load_frag<T, CH>(y[j], ybase + (int64_t)o * 8 * y_sn, cact && u < c.U);
            sga[j] = u < c.U ? (int)sa[u] : -1;
            sgb[j] = u < c.U ? (int)sb[u] : -1;
        }
#pragma unroll
        for (int j = 0; j < UB; ++j) {
            if (u0 + j >= c.U) break;
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int s = 0; s < CH / 4; ++s)            // k-step s: logical k (2t+e | 2t+8+e) <-> channel CH*t + 4s + (e | 2+e)
                mma16816(acc, xa.r[2 * s], xb.r[2 * s], xa.r[2 * s + 1], xb.r[2 * s + 1], y[j].r[2 * s], y[j].r[2 * s + 1], T());
            if (sga[j] >= 0) store_pair<T>(oa + 8 * sga[j], acc[0], acc[1]);
            if (sgb[j] >= 0) store_pair<T>(ob + 8 * sgb[j], acc[2], acc[3]);
        }
    }
    for (uint32_t imp = tile_imp_mask(pk, c.bt); imp; imp &= imp - 1) {          // impure tokens of this tile
        const int i = c.i0 + __ffs(imp) - 1;
        slow_dot_row<T>(X + c.b * x_sb + c.h * x_sh + (int64_t)i * x_sn, Y + c.b * y_sb + c.h * y_sh, y_sn,
                        idx + ((int64_t)c.b * Nq + i) * M, out + (((int64_t)c.b * H + c.h) * Nq + i) * M, C, M, c.lane);
    }
}

// dot: fp32 through 3xTF32
template <int CH>
__global__ void __launch_bounds__(TW * 32)
dot_tile32_kernel(const float *__restrict__ X, const float *__restrict__ Y, const int64_t *__restrict__ idx, const PackView pk,
                  float *__restrict__ out, int B, int H, int Nq, int C, int M,
                  int64_t x_sb, int64_t x_sh, int64_t x_sn, int64_t y_sb, int64_t y_sh, int64_t y_sn) {
    if (pk.flags[0]) return;
    TileCtx c;
    if (!tile_ctx(c, pk, B, H)) return;
    constexpr int UB = 2;
    constexpr int KS = CH / 2;                          // k-steps of 8
    const bool cact = CH * c.t < C;
    const int ra = c.i0 + c.g, rb = ra + 8;
    uint32_t ah[KS][4], al[KS][4];                      // A fragments (hi / lo), k-step s: channel CH*t + 2s + (0 | 1)
    {
        RowFrag<float, CH> xa, xb;
        const float *xbase = X + c.b * x_sb + c.h * x_sh + CH * c.t;
        load_frag<float, CH>(xa, xbase + (int64_t)ra * x_sn, cact && ra < Nq);
        load_frag<float, CH>(xb, xbase + (int64_t)rb * x_sn, cact && rb < Nq);
#pragma unroll
        for (int s = 0; s < KS; ++s) {
            tf32_split(__uint_as_float(xa.r[2 * s]), ah[s][0], al[s][0]);
            tf32_split(__uint_as_float(xb.r[2 * s]), ah[s][1], al[s][1]);
            tf32_split(__uint_as_float(xa.r[2 * s + 1]), ah[s][2], al[s][2]);
            tf32_split(__uint_as_float(xb.r[2 * s + 1]), ah[s][3], al[s][3]);
        }
    }
    const int8_t *sa = pk.slot_of + ((int64_t)c.bt * TILE_TOK + c.g) * U_MAX;
    const int8_t *sb = sa + 8 * U_MAX;
    const float *ybase = Y + c.b * y_sb + c.h * y_sh + CH * c.t + (int64_t)c.g * y_sn;
    float *oa = out + (((int64_t)c.b * H + c.h) * Nq + ra) * M + 2 * c.t;
    float *ob = oa + (int64_t)8 * M;
    for (int u0 = 0; u0 < c.U; u0 += UB) {
        RowFrag<float, CH> y[UB];
        int sga[UB], sgb[UB];
#pragma unroll
        for (int j = 0; j < UB; ++j) {
            const int u = u0 + j;
            const int o = tile_octet(c, u < c.U ? u : 0);
            load_frag<float, CH>(y[j], ybase + (int64_t)o * 8 * y_sn, cact && u < c.U);
            sga[j] = u < c.U ? (int)sa[u] : -1;
            sgb[j] = u < c.U ? (int)sb[u] : -1;
        }
#pragma unroll
        for (int j = 0; j < UB; ++j) {
            if (u0 + j >= c.U) break;
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int s = 0; s < KS; ++s) {
                uint32_t b0h, b0l, b1h, b1l;
                tf32_split(__uint_as_float(y[j].r[2 * s]), b0h, b0l);
                tf32_split(__uint_as_float(y[j].r[2 * s + 1]), b1h, b1l);
                mma_3xtf32(acc, ah[s], al[s], b0h, b1h, b0l, b1l);
            }
            if (sga[j] >= 0) store_pair<float>(oa + 8 * sga[j], acc[0], acc[1]);
            if (sgb[j] >= 0) store_pair<float>(ob + 8 * sgb[j], acc[2], acc[3]);
        }
    }
    for (uint32_t imp = tile_imp_mask(pk, c.bt); imp; imp &= imp - 1) {
        const int i = c.i0 + __ffs(imp) - 1;
        slow_dot_row<float>(X + c.b * x_sb + c.h * x_sh + (int64_t)i * x_sn, Y + c.b * y_sb + c.h * y_sh, y_sn,
                            idx + ((int64_t)c.b * Nq + i) * M, out + (((int64_t)c.b * H + c.h) * Nq + i) * M, C, M, c.lane);
    }
}

// ---- host side ---------------------------------------------------------------------------------------------------------
template <typename T> static bool rows_ok(const Rows4 &r) {
    constexpr int VPT = 16 / sizeof(T);
    return aligned16(r.p) && r.sb % VPT == 0 && r.sh % VPT == 0 && r.sn % VPT == 0;
}

// can the tile kernels take this call at all (the device-side flag decides the rest)
template <typename T> static bool tile_shape_ok(int C, int M) { return C % 8 == 0 && C >= 8 && C <= 32 && M % 8 == 0 && M <= 256; }

template <typename T>
int launch_dot_tile(const T *X, const T *Y, const int64_t *idx, const void *pack, T *out, int B, int H, int Nq, int Nk, int C,
                    int M, Rows4 x, Rows4 y, cudaStream_t st) {
    const PackView pk = pack_view(const_cast<void *>(pack), B, Nq, Nk);
    const int64_t items = (int64_t)B * pk.T * H;
    if (items == 0) return 0;
    const int grid = ceil_div(items, TW);
    if constexpr (sizeof(T) == 2) {
        if (C <= 16) dot_tile16_kernel<T, 4><<<grid, TW * 32, 0, st>>>(X, Y, idx, pk, out, B, H, Nq, C, M, x.sb, x.sh, x.sn, y.sb, y.sh, y.sn);
        else dot_tile16_kernel<T, 8><<<grid, TW * 32, 0, st>>>(X, Y, idx, pk, out, B, H, Nq, C, M, x.sb, x.sh, x.sn, y.sb, y.sh, y.sn);
    } else {
        if (C <= 16) dot_tile32_kernel<4><<<grid, TW * 32, 0, st>>>(X, Y, idx, pk, out, B, H, Nq, C, M, x.sb, x.sh, x.sn, y.sb, y.sh, y.sn);
        else dot_tile32_kernel<8><<<grid, TW * 32, 0, st>>>(X, Y, idx, pk, out, B, H, Nq, C, M, x.sb, x.sh, x.sn, y.sb, y.sh, y.sn);
    }
    note_launches(1);
    return check_launch("dot_tile");
}

template int launch_dot_tile<float>(const float *, const float *, const int64_t *, const void *, float *, int, int, int, int, int, int, Rows4, Rows4, cudaStream_t);
template int launch_dot_tile<__half>(const __half *, const __half *, const int64_t *, const void *, __half *, int, int, int, int, int, int, Rows4, Rows4, cudaStream_t);
template int launch_dot_tile<__nv_bfloat16>(const __nv_bfloat16 *, const __nv_bfloat16 *, const int64_t *, const void *, __nv_bfloat16 *, int, int, int, int, int, int, Rows4, Rows4, cudaStream_t);

template <typename T> bool tile_dot_eligible(int C, int M, Rows4 x, Rows4 y) { return tile_shape_ok<T>(C, M) && rows_ok<T>(x) && rows_ok<T>(y); }
template bool tile_dot_eligible<float>(int, int, Rows4, Rows4);
template bool tile_dot_eligible<__half>(int, int, Rows4, Rows4);
template bool tile_dot_eligible<__nv_bfloat16>(int, int, Rows4, Rows4);

// ---------------------------------------------------------------------------------------------------------------------
// axpy: out[16 x C] = sum over union octets of  W-block[16 x 8] * Y-octet[8 x C].  The contraction runs over KEYS, so the
// Y operand is needed key-major ("transposed"): octet rows are staged in shared memory with cp.async (double buffered,
// per warp) and read back with ldmatrix.trans (16-bit) or conflict-free scalar LDS (fp32).  W-blocks come straight from
// global memory: token row i contributes W[i, 8s .. 8s+7] where s is its slot for the octet, zeros when it has none.
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, bool pred) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    const int sz = pred ? 16 : 0;                                   // src-size 0 -> 16 bytes of zeros, no global read
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void *smem_row) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_row);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(s));
}

template <typename T> __device__ __forceinline__ uint32_t load_w_pair(const T *p, bool al4) {     // two consecutive 16-bit values
    if (al4) return __ldg(reinterpret_cast<const unsigned int *>(p));
    const unsigned short lo = __ldg(reinterpret_cast<const unsigned short *>(p));
    const unsigned short hi = __ldg(reinterpret_cast<const unsigned short *>(p) + 1);
    return (uint32_t)lo | ((uint32_t)hi << 16);
}

template <typename T, int NT>
__global__ void __launch_bounds__(TW * 32)
axpy_tile16_kernel(const T *__restrict__ W, const T *__restrict__ Y, const int64_t *__restrict__ idx, const PackView pk,
                   T *__restrict__ out, int B, int H, int Nq, int C, int M,
                   int64_t w_sb, int64_t w_sh, int64_t w_sn, int64_t y_sb, int64_t y_sh, int64_t y_sn,
                   int64_t o_sb, int64_t o_sh, int64_t o_sn, int w_al4) {
    constexpr int ROWB = NT * 16 + 16;                 // padded row: ldmatrix phases hit 8 distinct 16-byte bank groups
    constexpr int CPL = NT / 2;                        // 16-byte chunks per lane per stage (16 rows x NT chunks)
    __shared__ __align__(16) unsigned char smem_all[TW][2][16 * ROWB];
    if (pk.flags[0]) return;
    TileCtx c;
    if (!tile_ctx(c, pk, B, H)) return;
    unsigned char(*buf)[16 * ROWB] = smem_all[threadIdx.x >> 5];
    const int ra = c.i0 + c.g, rb = ra + 8;
    const int8_t *sa = pk.slot_of + ((int64_t)c.bt * TILE_TOK + c.g) * U_MAX;
    const int8_t *sb = sa + 8 * U_MAX;
    const T *wa = W + c.b * w_sb + c.h * w_sh + (int64_t)ra * w_sn + 2 * c.t;
    const T *wb = wa + 8 * w_sn;
    const T *ybase = Y + c.b * y_sb + c.h * y_sh;
    const int P = (c.U + 1) >> 1;                      // k-steps of two octets
    float acc[NT][4];
#pragma unroll
    for (int n = 0; n < NT; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;

    auto stage = [&](int p, int which) {
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
            const int ch = c.lane + 32 * j;            // chunk id: row = ch / NT (0..15), block = ch % NT
            const int row = ch / NT, blk = ch % NT;
            const int u = 2 * p + (row >> 3);
            const bool ok = u < c.U && 8 * blk < C;
            const int o = tile_octet(c, u < c.U ? u : 0);
            const T *src = ybase + ((int64_t)o * 8 + (row & 7)) * y_sn + 8 * blk;
            cp_async16(buf[which] + row * ROWB + blk * 16, ok ? (const void *)src : (const void *)Y, ok);
        }
        cp_async_commit();
    };

    if (P > 0) stage(0, 0);
    for (int p = 0; p < P; ++p) {
        if (p + 1 < P) stage(p + 1, (p + 1) & 1);
        const int u0 = 2 * p, u1 = 2 * p + 1;
        const int s00 = sa[u0], s10 = sb[u0];
        const int s01 = u1 < c.U ? (int)sa[u1] : -1, s11 = u1 < c.U ? (int)sb[u1] : -1;
        uint32_t a[4];
        a[0] = s00 >= 0 ? load_w_pair<T>(wa + 8 * s00, w_al4) : 0u;
        a[1] = s10 >= 0 ? load_w_pair<T>(wb + 8 * s10, w_al4) : 0u;
        a[2] = s01 >= 0 ? load_w_pair<T>(wa + 8 * s01, w_al4) : 0u;
        a[3] = s11 >= 0 ? load_w_pair<T>(wb + 8 * s11, w_al4) : 0u;
        if (p + 1 < P) cp_async_wait<1>(); else cp_async_wait<0>();
        __syncwarp();
        const unsigned char *bp = buf[p & 1];
        const int mi = c.lane >> 3;
        const unsigned char *lrow = bp + (((mi & 1) << 3) + (c.lane & 7)) * ROWB + (mi >> 1) * 16;
#pragma unroll
        for (int n = 0; n < NT; n += 2) {
            uint32_t bfr[4];
            ldmatrix_x4_trans(bfr, lrow + n * 16);
            mma16816(acc[n], a[0], a[1], a[2], a[3], bfr[0], bfr[1], T());
            mma16816(acc[n + 1], a[0], a[1], a[2], a[3], bfr[2], bfr[3], T());
        }
        __syncwarp();
    }
    const uint32_t impm = tile_imp_mask(pk, c.bt);
    T *oa = out + c.b * o_sb + c.h * o_sh + (int64_t)ra * o_sn + 2 * c.t;
    T *ob = oa + 8 * o_sn;
#pragma unroll
    for (int n = 0; n < NT; ++n) {
        if (8 * n + 2 * c.t < C) {
            if (ra < Nq && !((impm >> c.g) & 1u)) store_pair<T>(oa + 8 * n, acc[n][0], acc[n][1]);
            if (rb < Nq && !((impm >> (c.g + 8)) & 1u)) store_pair<T>(ob + 8 * n, acc[n][2], acc[n][3]);
        }
    }
    for (uint32_t imp = impm; imp; imp &= imp - 1) {
        const int i = c.i0 + __ffs(imp) - 1;
        slow_axpy_row<T>(W + c.b * w_sb + c.h * w_sh + (int64_t)i * w_sn, ybase, y_sn, idx + ((int64_t)c.b * Nq + i) * M,
                         out + c.b * o_sb + c.h * o_sh + (int64_t)i * o_sn, C, M, c.lane);
    }
}

template <int NT>
__global__ void __launch_bounds__(TW * 32)
axpy_tile32_kernel(const float *__restrict__ W, const float *__restrict__ Y, const int64_t *__restrict__ idx, const PackView pk,
                   float *__restrict__ out, int B, int H, int Nq, int C, int M,
                   int64_t w_sb, int64_t w_sh, int64_t w_sn, int64_t y_sb, int64_t y_sh, int64_t y_sn,
                   int64_t o_sb, int64_t o_sh, int64_t o_sn) {
    constexpr int RS = NT * 8 + 8;                     // padded row stride in floats: lanes (t, g) hit 32 distinct banks
    constexpr int CPL = NT / 2;                        // 16-byte chunks per lane per stage (8 rows x 2 NT chunks)
    __shared__ __align__(16) float smem_all[TW][2][8 * RS];
    if (pk.flags[0]) return;
    TileCtx c;
    if (!tile_ctx(c, pk, B, H)) return;
    float(*buf)[8 * RS] = smem_all[threadIdx.x >> 5];
    const int ra = c.i0 + c.g, rb = ra + 8;
    const int8_t *sa = pk.slot_of + ((int64_t)c.bt * TILE_TOK + c.g) * U_MAX;
    const int8_t *sb = sa + 8 * U_MAX;
    const float *wa = W + c.b * w_sb + c.h * w_sh + (int64_t)ra * w_sn + c.t;
    const float *wb = wa + 8 * w_sn;
    const float *ybase = Y + c.b * y_sb + c.h * y_sh;
    float acc[NT][4];
#pragma unroll
    for (int n = 0; n < NT; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;

    auto stage = [&](int u, int which) {
        const int o = tile_octet(c, u);
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
            const int ch = c.lane + 32 * j;            // row = ch / (2 NT) (0..7), 4-float block = ch % (2 NT)
            const int row = ch / (2 * NT), blk = ch % (2 * NT);
            const bool ok = 4 * blk < C;
            const float *src = ybase + ((int64_t)o * 8 + row) * y_sn + 4 * blk;
            cp_async16(buf[which] + row * RS + blk * 4, ok ? (const void *)src : (const void *)Y, ok);
        }
        cp_async_commit();
    };

    if (c.U > 0) stage(0, 0);
    for (int u = 0; u < c.U; ++u) {
        if (u + 1 < c.U) stage(u + 1, (u + 1) & 1);
        const int s0 = sa[u], s1 = sb[u];
        uint32_t ah[4], al[4];
        tf32_split(s0 >= 0 ? __ldg(wa + 8 * s0) : 0.f, ah[0], al[0]);
        tf32_split(s1 >= 0 ? __ldg(wb + 8 * s1) : 0.f, ah[1], al[1]);
        tf32_split(s0 >= 0 ? __ldg(wa + 8 * s0 + 4) : 0.f, ah[2], al[2]);
        tf32_split(s1 >= 0 ? __ldg(wb + 8 * s1 + 4) : 0.f, ah[3], al[3]);
        if (u + 1 < c.U) cp_async_wait<1>(); else cp_async_wait<0>();
        __syncwarp();
        const float *bp = buf[u & 1] + c.t * RS + c.g;
#pragma unroll
        for (int n = 0; n < NT; ++n) {
            uint32_t b0h, b0l, b1h, b1l;
            tf32_split(bp[8 * n], b0h, b0l);                 // (k = t,     n = g)  <-  Y[octet row t][8n + g]
            tf32_split(bp[4 * RS + 8 * n], b1h, b1l);        // (k = t + 4, n = g)
            mma_3xtf32(acc[n], ah, al, b0h, b1h, b0l, b1l);
        }
        __syncwarp();
    }
    const uint32_t impm = tile_imp_mask(pk, c.bt);
    float *oa = out + c.b * o_sb + c.h * o_sh + (int64_t)ra * o_sn + 2 * c.t;
    float *ob = oa + 8 * o_sn;
#pragma unroll
    for (int n = 0; n < NT; ++n) {
        if (8 * n + 2 * c.t < C) {
            if (ra < Nq && !((impm >> c.g) & 1u)) store_pair<float>(oa + 8 * n, acc[n][0], acc[n][1]);
            if (rb < Nq && !((impm >> (c.g + 8)) & 1u)) store_pair<float>(ob + 8 * n, acc[n][2], acc[n][3]);
        }
    }
    for (uint32_t imp = impm; imp; imp &= imp - 1) {
        const int i = c.i0 + __ffs(imp) - 1;
        slow_axpy_row<float>(W + c.b * w_sb + c.h * w_sh + (int64_t)i * w_sn, ybase, y_sn, idx + ((int64_t)c.b * Nq + i) * M,
                             out + c.b * o_sb + c.h * o_sh + (int64_t)i * o_sn, C, M, c.lane);
    }
}

template <typename T> bool tile_axpy_eligible(int C, int M, Rows4 w, Rows4 y, Rows4 o) {
    // W rows are read element-wise (any strides; 16-bit pairs need 4-byte alignment, checked at launch)
    (void)w;
    return tile_shape_ok<T>(C, M) && rows_ok<T>(y) && rows_ok<T>(o);
}

template <typename T>
int launch_axpy_tile(const T *W, const T *Y, const int64_t *idx, const void *pack, T *out, int B, int H, int Nq, int Nk, int C,
                     int M, Rows4 w, Rows4 y, Rows4 o, cudaStream_t st) {
    const PackView pk = pack_view(const_cast<void *>(pack), B, Nq, Nk);
    const int64_t items = (int64_t)B * pk.T * H;
    if (items == 0) return 0;
    const int grid = ceil_div(items, TW);
    if constexpr (sizeof(T) == 2) {
        const int al4 = ((reinterpret_cast<uintptr_t>(w.p) & 3u) == 0 && w.sb % 2 == 0 && w.sh % 2 == 0 && w.sn % 2 == 0) ? 1 : 0;
        if (C <= 16) axpy_tile16_kernel<T, 2><<<grid, TW * 32, 0, st>>>(W, Y, idx, pk, out, B, H, Nq, C, M, w.sb, w.sh, w.sn, y.sb, y.sh, y.sn, o.sb, o.sh, o.sn, al4);
        else axpy_tile16_kernel<T, 4><<<grid, TW * 32, 0, st>>>(W, Y, idx, pk, out, B, H, Nq, C, M, w.sb, w.sh, w.sn, y.sb, y.sh, y.sn, o.sb, o.sh, o.sn, al4);
    } else {
        if (C <= 16) axpy_tile32_kernel<2><<<grid, TW * 32, 0, st>>>(W, Y, idx, pk, out, B, H, Nq, C, M, w.sb, w.sh, w.sn, y.sb, y.sh, y.sn, o.sb, o.sh, o.sn);
        else axpy_tile32_kernel<4><<<grid, TW * 32, 0, st>>>(W, Y, idx, pk, out, B, H, Nq, C, M, w.sb, w.sh, w.sn, y.sb, y.sh, y.sn, o.sb, o.sh, o.sn);
    }
    note_launches(1);
    return check_launch("axpy_tile");
}

// ---------------------------------------------------------------------------------------------------------------------
// scatter: out[key rows of octets (o, o+1)][C] = sum over the tiles that reference them of  W-block^T[16 keys x 16 tok] *
// X-tile[16 tok x C].  One warp owns a PAIR of adjacent key octets (= one mma M-tile of 16 output rows; adjacent octets
// are referenced by nearly the same tiles) and one head, and walks the merged inverse lists of the pack (ascending tile
// order -> fixed summation order -> deterministic gradients, no atomics).  Both operands are needed token-major along
// k, so the X tile (cp.async, double buffered) and the W block are staged in shared memory and read with
// ldmatrix.trans (16-bit) / conflict-free scalar LDS (fp32, 3xTF32).
constexpr int TWS = 4;                   // warps per CTA for the scatter kernels (more shared memory per warp)

struct PairCtx {
    int b, h, o, lane, g, t;
    int pa, ea, pb, eb;                   // inverse-list cursors of octet o and o + 1
    const uint32_t *ent;
};

__device__ __forceinline__ bool pair_ctx(PairCtx &c, const PackView &pk, int B, int H) {
    const int NP = (pk.NO + 1) >> 1;
    const int64_t item = (int64_t)blockIdx.x * TWS + (threadIdx.x >> 5);
    if (item >= (int64_t)B * NP * H) return false;
    c.lane = threadIdx.x & 31;
    c.g = c.lane >> 2;
    c.t = c.lane & 3;
    c.h = (int)(item % H);
    const int bp = (int)(item / H);
    c.b = bp / NP;
    c.o = (bp - c.b * NP) * 2;
    const int *off = pk.oct_off + (int64_t)c.b * (pk.NO + 1);
    c.pa = off[c.o];
    c.ea = off[c.o + 1];
    c.pb = c.ea;
    c.eb = c.o + 1 < pk.NO ? off[c.o + 2] : c.ea;
    c.ent = pk.oct_ent + (int64_t)c.b * pk.T * U_MAX;
    return true;
}
// next tile of the merged lists; ua / ub = union position of octet o / o+1 in that tile or -1.  Warp-uniform.
__device__ __forceinline__ bool pair_next(PairCtx &c, int &tile, int &ua, int &ub) {
    if (c.pa >= c.ea && c.pb >= c.eb) return false;
    const unsigned ea = c.pa < c.ea ? __ldg(c.ent + c.pa) : 0xffffffffu;
    const unsigned eb = c.pb < c.eb ? __ldg(c.ent + c.pb) : 0xffffffffu;
    const unsigned ta = ea == 0xffffffffu ? 0xffffffffu : ea / U_MAX, tb = eb == 0xffffffffu ? 0xffffffffu : eb / U_MAX;
    const unsigned tm = min(ta, tb);
    tile = (int)tm;
    ua = ub = -1;
    if (ta == tm) { ua = (int)(ea - ta * U_MAX); ++c.pa; }
    if (tb == tm) { ub = (int)(eb - tb * U_MAX); ++c.pb; }
    return true;
}

// Contributions of impure tokens (left out of the tile structure) to the 16 key rows of this octet pair: walk the full
// inverse neighbour list of each flagged row (csr.cu; built because the pack has impure tokens), keep the entries of
// impure tokens, add them to the row just written.  Ascending (i, j) order -> deterministic.
template <typename T>
__device__ __forceinline__ void scat_fixup(const T *W, const T *X, const int32_t *csr_off, const uint32_t *csr_ent,
                                           const PackView &pk, T *out, const PairCtx &c, int Nq, int Nk, int C, int M,
                                           int64_t w_sb, int64_t w_sh, int64_t w_sn, int64_t x_sb, int64_t x_sh, int64_t x_sn,
                                           int64_t o_sb, int64_t o_sh, int64_t o_sn) {
    if (pk.flags[2] == 0) return;
    __syncwarp();
    for (int rr = 0; rr < 16; ++rr) {
        const int row = c.o * 8 + rr;
        if (row >= Nk || !pk.row_imp[(int64_t)c.b * Nk + row]) continue;
        const int lo = csr_off[(int64_t)c.b * (Nk + 1) + row], hi = csr_off[(int64_t)c.b * (Nk + 1) + row + 1];
        const uint32_t *ent = csr_ent + (int64_t)c.b * Nq * M;
        for (int ch = c.lane; ch < C; ch += 32) {
            float s = 0.f;
            for (int e = lo; e < hi; ++e) {
                const uint32_t pkd = ent[e];
                const int i = (int)(pkd >> 8), j = (int)(pkd & 255u);
                if (!pk.tok_imp[(int64_t)c.b * pk.T * TILE_TOK + i]) continue;
                s = fmaf(to_f(W[c.b * w_sb + c.h * w_sh + (int64_t)i * w_sn + j]), to_f(X[c.b * x_sb + c.h * x_sh + (int64_t)i * x_sn + ch]), s);
            }
            T *op = out + c.b * o_sb + c.h * o_sh + (int64_t)row * o_sn + ch;
            *op = from_f<T>(to_f(*op) + s);
        }
    }
}

template <typename T, int NT>
__global__ void __launch_bounds__(TWS * 32)
scat_tile16_kernel(const T *__restrict__ W, const T *__restrict__ X, const int32_t *__restrict__ csr_off,
                   const uint32_t *__restrict__ csr_ent, const PackView pk, T *__restrict__ out,
                   int B, int H, int Nq, int Nk, int C, int M,
                   int64_t w_sb, int64_t w_sh, int64_t w_sn, int64_t x_sb, int64_t x_sh, int64_t x_sn,
                   int64_t o_sb, int64_t o_sh, int64_t o_sn, int w_al16) {
    constexpr int ROWB = NT * 16 + 16;                 // X tile row (bytes), padded
    constexpr int WROWB = 48;                          // W block row: 16 keys x 2 bytes + 16 pad
    constexpr int CPL = NT / 2;
    __shared__ __align__(16) unsigned char xs_all[TWS][2][16 * ROWB];
    __shared__ __align__(16) unsigned char ws_all[TWS][16 * WROWB];
    if (pk.flags[0]) return;
    PairCtx c;
    if (!pair_ctx(c, pk, B, H)) return;
    unsigned char(*xs)[16 * ROWB] = xs_all[threadIdx.x >> 5];
    unsigned char *ws = ws_all[threadIdx.x >> 5];
    const T *xbase = X + c.b * x_sb + c.h * x_sh;
    const T *wbase = W + c.b * w_sb + c.h * w_sh;
    float acc[NT][4];
#pragma unroll
    for (int n = 0; n < NT; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;

    auto stage_x = [&](int tile, int which) {
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
            const int ch = c.lane + 32 * j;
            const int row = ch / NT, blk = ch % NT;
            const int i = tile * TILE_TOK + row;
            const bool ok = i < Nq && 8 * blk < C;
            cp_async16(xs[which] + row * ROWB + blk * 16, ok ? (const void *)(xbase + (int64_t)i * x_sn + 8 * blk) : (const void *)X, ok);
        }
        cp_async_commit();
    };

    int tile, ua, ub, ntile = -1, nua = -1, nub = -1;
    bool have = pair_next(c, tile, ua, ub);
    if (have) stage_x(tile, 0);
    int it = 0;
    while (have) {
        const bool more = pair_next(c, ntile, nua, nub);
        if (more) stage_x(ntile, (it + 1) & 1);
        // W block: lane -> (token = lane & 15, octet half = lane >> 4): 8 weights of that token for that octet or zeros
        {
            const int tok = c.lane & 15, half = c.lane >> 4;
            const int u = half ? ub : ua;
            const int i = tile * TILE_TOK + tok;
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (u >= 0) {
                const int s = pk.slot_of[((int64_t)(c.b * pk.T + tile) * TILE_TOK + tok) * U_MAX + u];
                if (s >= 0) {
                    const T *wp = wbase + (int64_t)i * w_sn + 8 * s;
                    if (w_al16) v = __ldg(reinterpret_cast<const uint4 *>(wp));
                    else {
                        const unsigned short *hp = reinterpret_cast<const unsigned short *>(wp);
                        v.x = __ldg(hp) | ((uint32_t)__ldg(hp + 1) << 16);
                        v.y = __ldg(hp + 2) | ((uint32_t)__ldg(hp + 3) << 16);
                        v.z = __ldg(hp + 4) | ((uint32_t)__ldg(hp + 5) << 16);
                        v.w = __ldg(hp + 6) | ((uint32_t)__ldg(hp + 7) << 16);
                    }
                }
            }
            *reinterpret_cast<uint4 *>(ws + tok * WROWB + half * 16) = v;
        }
        if (more) cp_async_wait<1>(); else cp_async_wait<0>();
        __syncwarp();
        const int mi = c.lane >> 3;
        uint32_t a[4];
        // A = W-block^T: matrices (tok 0-7 | keys o), (tok 0-7 | keys o+1), (tok 8-15 | keys o), (tok 8-15 | keys o+1)
        ldmatrix_x4_trans(a, ws + (((mi >> 1) << 3) + (c.lane & 7)) * WROWB + (mi & 1) * 16);
        const unsigned char *lrow = xs[it & 1] + (((mi & 1) << 3) + (c.lane & 7)) * ROWB + (mi >> 1) * 16;
#pragma unroll
        for (int n = 0; n < NT; n += 2) {
            uint32_t bfr[4];
            ldmatrix_x4_trans(bfr, lrow + n * 16);
            mma16816(acc[n], a[0], a[1], a[2], a[3], bfr[0], bfr[1], T());
            mma16816(acc[n + 1], a[0], a[1], a[2], a[3], bfr[2], bfr[3], T());
        }
        __syncwarp();
        have = more; tile = ntile; ua = nua; ub = nub; ++it;
    }
    const int ka = c.o * 8 + c.g, kb = ka + 8;
    T *oa = out + c.b * o_sb + c.h * o_sh + (int64_t)ka * o_sn + 2 * c.t;
    T *ob = oa + 8 * o_sn;
#pragma unroll
    for (int n = 0; n < NT; ++n) {
        if (8 * n + 2 * c.t < C) {
            if (ka < Nk) store_pair<T>(oa + 8 * n, acc[n][0], acc[n][1]);
            if (kb < Nk) store_pair<T>(ob + 8 * n, acc[n][2], acc[n][3]);
        }
    }
    scat_fixup<T>(W, X, csr_off, csr_ent, pk, out, c, Nq, Nk, C, M, w_sb, w_sh, w_sn, x_sb, x_sh, x_sn, o_sb, o_sh, o_sn);
}

template <int NT>
__global__ void __launch_bounds__(TWS * 32)
scat_tile32_kernel(const float *__restrict__ W, const float *__restrict__ X, const int32_t *__restrict__ csr_off,
                   const uint32_t *__restrict__ csr_ent, const PackView pk, float *__restrict__ out,
                   int B, int H, int Nq, int Nk, int C, int M,
                   int64_t w_sb, int64_t w_sh, int64_t w_sn, int64_t x_sb, int64_t x_sh, int64_t x_sn,
                   int64_t o_sb, int64_t o_sh, int64_t o_sn, int w_al16) {
    constexpr int RS = NT * 8 + 8;                     // X tile row stride (floats)
    constexpr int WRS = 24;                            // W block row stride: 16 keys + 8 pad
    constexpr int CPL = NT;                            // 16 rows x 2 NT chunks of 4 floats / 32 lanes
    __shared__ __align__(16) float xs_all[TWS][2][16 * RS];
    __shared__ __align__(16) float ws_all[TWS][16 * WRS];
    if (pk.flags[0]) return;
    PairCtx c;
    if (!pair_ctx(c, pk, B, H)) return;
    float(*xs)[16 * RS] = xs_all[threadIdx.x >> 5];
    float *ws = ws_all[threadIdx.x >> 5];
    const float *xbase = X + c.b * x_sb + c.h * x_sh;
    const float *wbase = W + c.b * w_sb + c.h * w_sh;
    float acc[NT][4];
#pragma unroll
    for (int n = 0; n < NT; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;

    auto stage_x = [&](int tile, int which) {
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
            const int ch = c.lane + 32 * j;
            const int row = ch / (2 * NT), blk = ch % (2 * NT);
            const int i = tile * TILE_TOK + row;
            const bool ok = i < Nq && 4 * blk < C;
            cp_async16(xs[which] + row * RS + blk * 4, ok ? (const void *)(xbase + (int64_t)i * x_sn + 4 * blk) : (const void *)X, ok);
        }
        cp_async_commit();
    };

    int tile, ua, ub, ntile = -1, nua = -1, nub = -1;
    bool have = pair_next(c, tile, ua, ub);
    if (have) stage_x(tile, 0);
    int it = 0;
    while (have) {
        const bool more = pair_next(c, ntile, nua, nub);
        if (more) stage_x(ntile, (it + 1) & 1);
        {
            const int tok = c.lane & 15, half = c.lane >> 4;
            const int u = half ? ub : ua;
            const int i = tile * TILE_TOK + tok;
            float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
            if (u >= 0) {
                const int s = pk.slot_of[((int64_t)(c.b * pk.T + tile) * TILE_TOK + tok) * U_MAX + u];
                if (s >= 0) {
                    const float *wp = wbase + (int64_t)i * w_sn + 8 * s;
                    if (w_al16) {
                        v0 = __ldg(reinterpret_cast<const float4 *>(wp));
                        v1 = __ldg(reinterpret_cast<const float4 *>(wp) + 1);
                    } else {
                        v0 = make_float4(__ldg(wp), __ldg(wp + 1), __ldg(wp + 2), __ldg(wp + 3));
                        v1 = make_float4(__ldg(wp + 4), __ldg(wp + 5), __ldg(wp + 6), __ldg(wp + 7));
                    }
                }
            }
            float4 *dst = reinterpret_cast<float4 *>(ws + tok * WRS + half * 8);
            dst[0] = v0;
            dst[1] = v1;
        }
        if (more) cp_async_wait<1>(); else cp_async_wait<0>();
        __syncwarp();
        const float *xb = xs[it & 1];
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {               // k-step = 8 tokens
            uint32_t ah[4], al[4];
            const float *wr = ws + (8 * ks + c.t) * WRS + c.g;
            tf32_split(wr[0], ah[0], al[0]);                     // (row = key g of o,     k = tok t)
            tf32_split(wr[8], ah[1], al[1]);                     // (row = key g of o + 1, k = tok t)
            tf32_split(wr[4 * WRS], ah[2], al[2]);               // k = tok t + 4
            tf32_split(wr[4 * WRS + 8], ah[3], al[3]);
            const float *bp = xb + (8 * ks + c.t) * RS + c.g;
#pragma unroll
            for (int n = 0; n < NT; ++n) {
                uint32_t b0h, b0l, b1h, b1l;
                tf32_split(bp[8 * n], b0h, b0l);
                tf32_split(bp[4 * RS + 8 * n], b1h, b1l);
                mma_3xtf32(acc[n], ah, al, b0h, b1h, b0l, b1l);
            }
        }
        __syncwarp();
        have = more; tile = ntile; ua = nua; ub = nub; ++it;
    }
    const int ka = c.o * 8 + c.g, kb = ka + 8;
    float *oa = out + c.b * o_sb + c.h * o_sh + (int64_t)ka * o_sn + 2 * c.t;
    float *ob = oa + 8 * o_sn;
#pragma unroll
    for (int n = 0; n < NT; ++n) {
        if (8 * n + 2 * c.t < C) {
            if (ka < Nk) store_pair<float>(oa + 8 * n, acc[n][0], acc[n][1]);
            if (kb < Nk) store_pair<float>(ob + 8 * n, acc[n][2], acc[n][3]);
        }
    }
    scat_fixup<float>(W, X, csr_off, csr_ent, pk, out, c, Nq, Nk, C, M, w_sb, w_sh, w_sn, x_sb, x_sh, x_sn, o_sb, o_sh, o_sn);
}

template <typename T> bool tile_scat_eligible(int C, int M, Rows4 w, Rows4 x, Rows4 o) {
    (void)w;
    return tile_shape_ok<T>(C, M) && rows_ok<T>(x) && rows_ok<T>(o);
}

template <typename T>
int launch_scat_tile(const T *W, const T *X, const int32_t *csr_off, const uint32_t *csr_ent, const void *pack, T *out,
                     int B, int H, int Nq, int Nk, int C, int M, Rows4 w, Rows4 x, Rows4 o, cudaStream_t st) {
    const PackView pk = pack_view(const_cast<void *>(pack), B, Nq, Nk);
    const int64_t items = (int64_t)B * ((pk.NO + 1) / 2) * H;
    if (items == 0) return 0;
    const int grid = ceil_div(items, TWS);
    constexpr int VPT = 16 / sizeof(T);
    const int al16 = (aligned16(w.p) && w.sb % VPT == 0 && w.sh % VPT == 0 && w.sn % VPT == 0) ? 1 : 0;
    if constexpr (sizeof(T) == 2) {
        if (C <= 16) scat_tile16_kernel<T, 2><<<grid, TWS * 32, 0, st>>>(W, X, csr_off, csr_ent, pk, out, B, H, Nq, Nk, C, M, w.sb, w.sh, w.sn, x.sb, x.sh, x.sn, o.sb, o.sh, o.sn, al16);
        else scat_tile16_kernel<T, 4><<<grid, TWS * 32, 0, st>>>(W, X, csr_off, csr_ent, pk, out, B, H, Nq, Nk, C, M, w.sb, w.sh, w.sn, x.sb, x.sh, x.sn, o.sb, o.sh, o.sn, al16);
    } else {
        if (C <= 16) scat_tile32_kernel<2><<<grid, TWS * 32, 0, st>>>(W, X, csr_off, csr_ent, pk, out, B, H, Nq, Nk, C, M, w.sb, w.sh, w.sn, x.sb, x.sh, x.sn, o.sb, o.sh, o.sn, al16);
        else scat_tile32_kernel<4><<<grid, TWS * 32, 0, st>>>(W, X, csr_off, csr_ent, pk, out, B, H, Nq, Nk, C, M, w.sb, w.sh, w.sn, x.sb, x.sh, x.sn, o.sb, o.sh, o.sn, al16);
    }
    note_launches(1);
    return check_launch("scat_tile");
}

#define INST(T) \
    template bool tile_axpy_eligible<T>(int, int, Rows4, Rows4, Rows4); \
    template bool tile_scat_eligible<T>(int, int, Rows4, Rows4, Rows4); \
    template int launch_axpy_tile<T>(const T *, const T *, const int64_t *, const void *, T *, int, int, int, int, int, int, Rows4, Rows4, Rows4, cudaStream_t); \
    template int launch_scat_tile<T>(const T *, const T *, const int32_t *, const uint32_t *, const void *, T *, int, int, int, int, int, int, Rows4, Rows4, Rows4, cudaStream_t);
INST(float) INST(__half) INST(__nv_bfloat16)
#undef INST

}  // namespace clusten
