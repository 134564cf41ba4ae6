// Second-generation tile-union kernels for the dot and axpy shapes of CLUSTEN QK / AV (see clusten_tile.cu for the
// scheme).  ncu on the first generation (profiles/r1_tile_small_s0_bf16_v1.csv) showed them instruction-issue bound:
// ~930 warp instructions per (16-token tile, head), most of it 64-bit address arithmetic, per-head re-reads of the tile
// metadata and `u < U` predication.  Here one warp owns a tile and a GROUP of HG heads:
//   * the per-octet work that does not depend on the head (octet id shuffle, slot bytes, row offsets) is done once per
//     group, the inner head loop is LDG.128 + MMAs + stores only;
//   * all offsets are 32-bit element offsets from per-warp 64-bit bases (the host checks the extents);
//   * the union is walked in whole groups of UB octets: positions >= U re-read the tile's last octet with slot -1, so
//     the loop body carries no bounds predicates.
#include "t2.cuh"

namespace clusten {

namespace t2 {

struct DotArgs {
    const void *X, *Y;
    const int64_t *idx;
    void *out;
    int B, H, Nq, C, M;
    int x_sh, x_sn, y_sh, y_sn;          // element strides within one batch element (< 2^31, host-checked)
    int64_t x_sb, y_sb;
};


// dot, 16-bit types.  CH = channels per lane (4: C <= 16, 8: C <= 32), HG heads per warp, UB octets per step.
template <typename T, int CH, int HG, int UB>
__global__ void __launch_bounds__(TW2 * 32)
dot2_kernel16(const DotArgs a, const PackView pk) {
    if (pk.flags[0]) return;
    constexpr int NR = CH / 2;                           // 32-bit registers per row chunk
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int NG = a.H / HG;
    const int tile = blockIdx.x * TW2 + (threadIdx.x >> 5);          // grid: (tiles / TW2, B * head groups)
    if (tile < pk.T) {
    const int b = blockIdx.y / NG, hg = blockIdx.y - b * NG;
    const int bt = b * pk.T + tile, i0 = tile * TILE_TOK;
    const int h0 = hg * HG;
    const int U = pk.tile_u[bt];
    const int *octp = pk.tile_oct + bt * U_MAX;
    const int oc0 = octp[lane], oc1 = octp[32 + (lane & 15)];
    const bool cact = CH * t < a.C;
    const T *Xb = opaque(reinterpret_cast<const T *>(a.X) + b * a.x_sb + h0 * a.x_sh);
    const T *Yb = opaque(reinterpret_cast<const T *>(a.Y) + b * a.y_sb + h0 * a.y_sh);
    T *Ob = opaque(reinterpret_cast<T *>(a.out) + ((int64_t)(b * a.H + h0) * a.Nq + i0) * a.M);
    const int NqM = a.Nq * a.M;
    // A fragments (rows g and g + 8 of the tile) of every head of the group
    uint32_t xa[HG][NR], xb[HG][NR];
    {
        // rows beyond Nq re-read the last row (their slots are all -1: nothing of them is stored)
        const int ra = min(i0 + g, a.Nq - 1), rb = min(i0 + g + 8, a.Nq - 1);
        const int offa = ra * a.x_sn + (cact ? CH * t : 0), offb = rb * a.x_sn + (cact ? CH * t : 0);
#pragma unroll
        for (int hh = 0; hh < HG; ++hh) {
            ld_chunk<CH * 2>(xa[hh], at(Xb, offa + hh * a.x_sh));
            ld_chunk<CH * 2>(xb[hh], at(Xb, offb + hh * a.x_sh));
            if (!cact) {
#pragma unroll
                for (int x = 0; x < NR; ++x) xa[hh][x] = xb[hh][x] = 0u;
            }
        }
    }
    const int8_t *sa = pk.slot_of + (bt * TILE_TOK + g) * U_MAX;
    const int8_t *sb = sa + 8 * U_MAX;
    const int ylane = g * a.y_sn + (cact ? CH * t : 0);  // inactive channel lanes read (and zero-multiply) chunk 0
    const int oa = g * a.M + 2 * t, ob = oa + 8 * a.M;
    const int y8 = 8 * a.y_sn;
    for (int u0 = 0; u0 < U; u0 += UB) {
        uint32_t sa4, sb4;
        if constexpr (UB == 4) {
            sa4 = __ldg(reinterpret_cast<const uint32_t *>(sa + u0));
            sb4 = __ldg(reinterpret_cast<const uint32_t *>(sb + u0));
        } else {
            sa4 = __ldg(reinterpret_cast<const unsigned short *>(sa + u0));
            sb4 = __ldg(reinterpret_cast<const unsigned short *>(sb + u0));
        }
        uint32_t y[UB][HG][NR];
#pragma unroll
        for (int j = 0; j < UB; ++j) {
            const int u = min(u0 + j, U - 1);
            const int o = __shfl_sync(FULL, u < 32 ? oc0 : oc1, u & 31);
            const int yo = o * y8 + ylane;
#pragma unroll
            for (int hh = 0; hh < HG; ++hh) { const int off = yo + hh * a.y_sh; ld_chunk<CH * 2>(y[j][hh], at(Yb, off)); }
        }
#pragma unroll
        for (int j = 0; j < UB; ++j) {
            const int s0 = sbyte(sa4, j), s1 = sbyte(sb4, j);
#pragma unroll
            for (int hh = 0; hh < HG; ++hh) {
                float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int s = 0; s < CH / 4; ++s)
                    mma16<T>(acc, xa[hh][2 * s], xb[hh][2 * s], xa[hh][2 * s + 1], xb[hh][2 * s + 1], y[j][hh][2 * s], y[j][hh][2 * s + 1]);
                const int fa = hh * NqM + oa + 8 * s0, fb = hh * NqM + ob + 8 * s1;
                st_pair_if<T>(at(Ob, fa), acc[0], acc[1], s0);
                st_pair_if<T>(at(Ob, fb), acc[2], acc[3], s1);
            }
        }
    }
    }
    // impure tokens of the whole call, spread over the grid: item = (list position, head)
    const SlowIter si = slow_items(pk, a.H);
    for (int it = si.first; it < si.n; it += si.stride) {
        const int gi = pk.imp_list[it / a.H], h = it % a.H;
        const int b = gi / a.Nq, i = gi - b * a.Nq;
        dot_row_fast<T>(reinterpret_cast<const T *>(a.X) + b * a.x_sb + h * a.x_sh + i * a.x_sn,
                        reinterpret_cast<const T *>(a.Y) + b * a.y_sb + h * a.y_sh, a.y_sn, a.idx + (int64_t)gi * a.M,
                        reinterpret_cast<T *>(a.out) + ((int64_t)(b * a.H + h) * a.Nq + i) * a.M, a.C, a.M, lane);
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// axpy, 16-bit types: out[16 x C] = sum over union octets of W-block[16 x 8] * Y-octet[8 x C] for HG heads.
// Per warp in shared memory: the W tiles of the HG heads (16 rows x w_sn elements each, one contiguous 16-byte-aligned
// block of global memory -> cp.async, coalesced) and a double-buffered stage of two Y octets x HG heads (key-major,
// read back with ldmatrix.trans).  A fragments are 32-bit LDS from the W tile at (row, 8 * slot + 2t).
struct AxpyArgs {
    const void *W, *Y;
    const int64_t *idx;
    void *out;
    int B, H, Nq, C, M;
    int w_sh, w_sn, y_sh, y_sn, o_sh, o_sn;
    int64_t w_sb, y_sb, o_sb;
    int smem_per_warp, wtile_bytes;      // wtile_bytes = 16 * w_sn * 2 rounded up to 16
};


template <typename T, int NT, int HG, bool AL4>
__global__ void __launch_bounds__(128)
axpy2_kernel16(const AxpyArgs a, const PackView pk) {
    extern __shared__ __align__(16) unsigned char dyn2[];
    if (pk.flags[0]) return;
    constexpr int ROWB = NT * 16 + 16;                   // staged Y row (one head): NT 16-byte chunks + pad
    constexpr int HSTG = 16 * ROWB;                      // one head of one stage
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int NG = a.H / HG;
    const int tile = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (tile < pk.T) {
    const int b = blockIdx.y / NG, hg = blockIdx.y - b * NG;
    const int bt = b * pk.T + tile, i0 = tile * TILE_TOK;
    const int h0 = hg * HG;
    const int U = pk.tile_u[bt];
    const int *octp = pk.tile_oct + bt * U_MAX;
    const int oc0 = octp[lane], oc1 = octp[32 + (lane & 15)];
    const uint32_t sm0 = (uint32_t)__cvta_generic_to_shared(dyn2) + (threadIdx.x >> 5) * a.smem_per_warp;
    const uint32_t smW = sm0 + 2 * HG * HSTG;            // W tiles behind the two Y stages
    const T *Wb = opaque(reinterpret_cast<const T *>(a.W) + b * a.w_sb + h0 * a.w_sh);
    const T *Yb = opaque(reinterpret_cast<const T *>(a.Y) + b * a.y_sb + h0 * a.y_sh);
    // ---- W tiles: rows i0 .. i0+15 of every head are one contiguous block; rows beyond Nq are not read (their slots are -1)
    {
        const int nbytes = min(TILE_TOK, a.Nq - i0) * a.w_sn * 2;
        const int nchunk = nbytes >> 4, tail = (nbytes & 15) >> 1;        // whole 16-byte chunks, then < 8 single elements
#pragma unroll
        for (int hh = 0; hh < HG; ++hh) {
            const unsigned char *src = reinterpret_cast<const unsigned char *>(at(Wb, hh * a.w_sh + i0 * a.w_sn));
            for (int c = lane; c < nchunk; c += 32) cp16(smW + hh * a.wtile_bytes + 16 * c, src + 16 * c);
            if (lane < tail) {
                const unsigned short v = __ldg(reinterpret_cast<const unsigned short *>(src + 16 * nchunk) + lane);
                asm volatile("st.shared.u16 [%0], %1;" ::"r"(smW + hh * a.wtile_bytes + 16 * nchunk + 2 * lane), "h"(v));
            }
        }
    }
    // ---- Y staging: lane -> (16-byte block blk = lane % NT, row lane / NT of a pass of 32 / NT rows); passes walk the
    // 16 rows (two octets) of every head, so only two per-lane offsets stay live across the main loop
    constexpr int RPP = 32 / NT;                         // rows per pass (8 or 16)
    const int blk = lane % NT, prow = lane / NT;
    const bool act = 8 * blk < a.C;
    const int src_lane = (prow & 7) * a.y_sn + 8 * blk;
    const uint32_t dst_lane = sm0 + prow * ROWB + blk * 16;
    const int y8 = 8 * a.y_sn;
    auto octet = [&](int u) { u = min(u, U - 1); return __shfl_sync(FULL, u < 32 ? oc0 : oc1, u & 31); };
    auto stage = [&](int p, int which) {
        const int o0 = octet(2 * p) * y8 + src_lane, o1 = octet(2 * p + 1) * y8 + src_lane;
        if (act) {
#pragma unroll
            for (int hh = 0; hh < HG; ++hh) {
                if constexpr (RPP == 8) {
                    cp16(dst_lane + which * (HG * HSTG) + hh * HSTG, at(Yb, o0 + hh * a.y_sh));
                    cp16(dst_lane + which * (HG * HSTG) + hh * HSTG + 8 * ROWB, at(Yb, o1 + hh * a.y_sh));
                } else {
                    cp16(dst_lane + which * (HG * HSTG) + hh * HSTG, at(Yb, (prow >= 8 ? o1 : o0) + hh * a.y_sh));
                }
            }
        }
        cp_commit();
    };
    float acc[HG][NT][4];
#pragma unroll
    for (int hh = 0; hh < HG; ++hh)
#pragma unroll
        for (int n = 0; n < NT; ++n) acc[hh][n][0] = acc[hh][n][1] = acc[hh][n][2] = acc[hh][n][3] = 0.f;
    if (8 * NT > a.C) {                                  // channel blocks beyond C are never staged: zero them once
        for (int x = lane; x < 2 * HG * HSTG / 16; x += 32)
            asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(sm0 + 16 * x), "r"(0u));
        __syncwarp();
    }
    const int8_t *sa = pk.slot_of + (bt * TILE_TOK + g) * U_MAX;
    const int8_t *sb = sa + 8 * U_MAX;
    const int P = (U + 1) >> 1;
    const uint32_t wa = smW + (g * a.w_sn + 2 * t) * 2, wb = wa + 16 * a.w_sn;      // (row g | g+8, col 2t) of head 0
    const int mi = lane >> 3;
    const uint32_t lrow = sm0 + (((mi & 1) << 3) + (lane & 7)) * ROWB + (mi >> 1) * 16;
    if (P > 0) stage(0, 0); else cp_commit();
    for (int p = 0; p < P; ++p) {
        if (p + 1 < P) stage(p + 1, (p + 1) & 1);
        const uint32_t s2a = __ldg(reinterpret_cast<const unsigned short *>(sa + 2 * p));
        const uint32_t s2b = __ldg(reinterpret_cast<const unsigned short *>(sb + 2 * p));
        const int s00 = sbyte(s2a, 0), s10 = sbyte(s2b, 0);
        int s01 = sbyte(s2a, 1), s11 = sbyte(s2b, 1);
        if (2 * p + 1 >= U) s01 = s11 = -1;              // odd U: the second octet of the last pair is a re-read of the first
        if (p + 1 < P) cp_wait<1>(); else cp_wait<0>();
        __syncwarp();
        const uint32_t yst = lrow + (p & 1) * (HG * HSTG);
#pragma unroll
        for (int hh = 0; hh < HG; ++hh) {
            uint32_t af[4];
            af[0] = lds_pair<AL4>(wa + hh * a.wtile_bytes + 16 * max(s00, 0));
            af[1] = lds_pair<AL4>(wb + hh * a.wtile_bytes + 16 * max(s10, 0));
            af[2] = lds_pair<AL4>(wa + hh * a.wtile_bytes + 16 * max(s01, 0));
            af[3] = lds_pair<AL4>(wb + hh * a.wtile_bytes + 16 * max(s11, 0));
            if (s00 < 0) af[0] = 0u;
            if (s10 < 0) af[1] = 0u;
            if (s01 < 0) af[2] = 0u;
            if (s11 < 0) af[3] = 0u;
#pragma unroll
            for (int n = 0; n < NT; n += 2) {
                uint32_t bfr[4];
                ldsm4t(bfr, yst + hh * HSTG + n * 16);
                mma16<T>(acc[hh][n], af[0], af[1], af[2], af[3], bfr[0], bfr[1]);
                mma16<T>(acc[hh][n + 1], af[0], af[1], af[2], af[3], bfr[2], bfr[3]);
            }
        }
        __syncwarp();
    }
    cp_wait<0>();
    uint32_t impm = 0;
    {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(pk.tok_imp + bt * TILE_TOK));
        if (v.x | v.y | v.z | v.w) {
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int k = 0; k < 4; ++k) impm |= ((w[q] >> (8 * k)) & 1u) << (4 * q + k);
        }
    }
    T *Ob = opaque(reinterpret_cast<T *>(a.out) + b * a.o_sb + h0 * a.o_sh);
    {
        const int ra = i0 + g, rb = ra + 8;
        const int sta = (ra < a.Nq && !((impm >> g) & 1u)) ? 0 : -1, stb = (rb < a.Nq && !((impm >> (g + 8)) & 1u)) ? 0 : -1;
        const int oa = ra * a.o_sn + 2 * t, ob = rb * a.o_sn + 2 * t;
#pragma unroll
        for (int hh = 0; hh < HG; ++hh)
#pragma unroll
            for (int n = 0; n < NT; ++n) {
                const int ok = 8 * n + 2 * t < a.C ? 0 : -1;
                st_pair_if<T>(at(Ob, hh * a.o_sh + oa + 8 * n), acc[hh][n][0], acc[hh][n][1], sta | ok);
                st_pair_if<T>(at(Ob, hh * a.o_sh + ob + 8 * n), acc[hh][n][2], acc[hh][n][3], stb | ok);
            }
    }
    }
    const SlowIter si = slow_items(pk, a.H);
    for (int it = si.first; it < si.n; it += si.stride) {
        const int gi = pk.imp_list[it / a.H], h = it % a.H;
        const int b = gi / a.Nq, i = gi - b * a.Nq;
        axpy_row_fast<T>(reinterpret_cast<const T *>(a.W) + b * a.w_sb + h * a.w_sh + i * a.w_sn,
                         reinterpret_cast<const T *>(a.Y) + b * a.y_sb + h * a.y_sh, a.y_sn, a.idx + (int64_t)gi * a.M,
                         reinterpret_cast<T *>(a.out) + b * a.o_sb + h * a.o_sh + i * a.o_sn, a.C, a.M, lane);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// scatter, 16-bit types: out[16 key rows of octets (o, o+1)][C] = sum over the tiles referencing them of
// W-block^T[16 keys x 16 tok] * X-tile[16 tok x C], for HG heads per warp.  The merged inverse lists of the pack are walked
// once per head group; X tiles and W blocks are staged with cp.async (double buffered; W blocks zero-filled where a token
// does not reference the octet) and read with ldmatrix.trans.  Ascending tile order -> deterministic sums, no atomics.
struct ScatArgs {
    const void *W, *X;
    const int32_t *csr_off;
    const uint32_t *csr_ent;
    void *out;
    int B, H, Nq, Nk, C, M;
    int w_sh, w_sn, x_sh, x_sn, o_sh, o_sn;
    int64_t w_sb, x_sb, o_sb;
    int smem_per_warp;
};


template <typename T, int NT, int HG, bool WAL>
__global__ void __launch_bounds__(128)
scat2_kernel16(const ScatArgs a, const PackView pk) {
    extern __shared__ __align__(16) unsigned char dyn2[];
    if (pk.flags[0]) return;
    constexpr int ROWB = NT * 16 + 16;                   // X tile row of one head
    constexpr int XSTG = 16 * ROWB;
    constexpr int WROWB = 48;                            // W block row: 16 keys x 2 bytes + 16 pad
    constexpr int WSTG = 16 * WROWB;
    constexpr int HSTG = XSTG + WSTG;                    // one head of one stage
    constexpr int RPP = 32 / NT;
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int NG = a.H / HG;
    const int NP = (pk.NO + 1) >> 1;
    const int pair = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (pair >= NP) return;
    const int b = blockIdx.y / NG, hg = blockIdx.y - b * NG;
    const int h0 = hg * HG, o = pair * 2;
    const uint32_t sm0 = (uint32_t)__cvta_generic_to_shared(dyn2) + (threadIdx.x >> 5) * a.smem_per_warp;
    const T *Wb = opaque(reinterpret_cast<const T *>(a.W) + b * a.w_sb + h0 * a.w_sh);
    const T *Xb = opaque(reinterpret_cast<const T *>(a.X) + b * a.x_sb + h0 * a.x_sh);
    const int8_t *slotb = opaque(pk.slot_t + (int64_t)b * pk.T * (TILE_TOK * U_MAX));
    // inverse-list cursors of octets o and o + 1; the first 32 entries of either list are fetched once (one per lane) and
    // handed out by shuffle, so the walk does not start every step with a dependent global load
    const int *off = pk.oct_off + b * (pk.NO + 1) + o;
    const int pa0 = off[0];
    int pa = pa0;
    const int ea = off[1];
    const int pb0 = ea;
    int pb = ea;
    const int eb = o + 1 < pk.NO ? off[2] : ea;
    const uint32_t *ent = opaque(pk.oct_ent + (int64_t)b * pk.T * U_MAX);
    const unsigned pre_a = pa0 + lane < ea ? ldg4(ent + pa0 + lane) : 0xffffffffu;
    const unsigned pre_b = pb0 + lane < eb ? ldg4(ent + pb0 + lane) : 0xffffffffu;
    auto next = [&](int &tile, int &ua, int &ub) -> bool {
        if (pa >= ea && pb >= eb) return false;
        unsigned va = __shfl_sync(FULL, pre_a, (pa - pa0) & 31), vb = __shfl_sync(FULL, pre_b, (pb - pb0) & 31);
        if (pa - pa0 >= 32) va = ldg4(ent + min(pa, ea - 1));          // (lists longer than a warp: rare)
        if (pb - pb0 >= 32) vb = ldg4(ent + min(pb, max(eb - 1, pb0)));
        if (pa >= ea) va = 0xffffffffu;
        if (pb >= eb) vb = 0xffffffffu;
        const unsigned ta = va == 0xffffffffu ? va : va / U_MAX, tb = vb == 0xffffffffu ? vb : vb / U_MAX;
        const unsigned tm = min(ta, tb);
        tile = (int)tm;
        ua = ub = -1;
        if (ta == tm) { ua = (int)(va - ta * U_MAX); ++pa; }
        if (tb == tm) { ub = (int)(vb - tb * U_MAX); ++pb; }
        return true;
    };
    // staging lanes: X tile -> (block lane % NT, row lane / NT per pass); W block -> (token lane & 15, octet half lane >> 4)
    const int blk = lane % NT, prow = lane / NT;
    const bool xact = 8 * blk < a.C;
    const uint32_t xdst = sm0 + prow * ROWB + blk * 16;
    const int tok = lane & 15, half = lane >> 4;
    const uint32_t wdst = sm0 + XSTG + tok * WROWB + half * 16;
    auto stage = [&](int tile, int ua, int ub, int which) {
        const uint32_t sbase = which * (HG * HSTG);
        if (xact) {
#pragma unroll
            for (int ps = 0; ps < 16 / RPP; ++ps) {
                const int row = min(tile * TILE_TOK + ps * RPP + prow, a.Nq - 1);   // rows beyond Nq: weights are zero
                const int xo = row * a.x_sn + 8 * blk;
#pragma unroll
                for (int hh = 0; hh < HG; ++hh) cp16(xdst + sbase + hh * HSTG + ps * RPP * ROWB, at(Xb, xo + hh * a.x_sh));
            }
        }
        const int u = half ? ub : ua;
        int s = -1;
        if (u >= 0) s = (int)slotb[(tile * U_MAX + u) * TILE_TOK + tok];      // 16 contiguous bytes per (tile, u)
        const int wo = (tile * TILE_TOK + tok) * a.w_sn + 8 * max(s, 0);
#pragma unroll
        for (int hh = 0; hh < HG; ++hh) {
            if constexpr (WAL) {
                cp16_zfill(wdst + sbase + hh * HSTG, at(Wb, wo + hh * a.w_sh), s >= 0);
            } else {
                uint4 v = make_uint4(0u, 0u, 0u, 0u);
                if (s >= 0) {
                    const unsigned short *hp = reinterpret_cast<const unsigned short *>(at(Wb, wo + hh * a.w_sh));
                    v.x = __ldg(hp) | ((uint32_t)__ldg(hp + 1) << 16);
                    v.y = __ldg(hp + 2) | ((uint32_t)__ldg(hp + 3) << 16);
                    v.z = __ldg(hp + 4) | ((uint32_t)__ldg(hp + 5) << 16);
                    v.w = __ldg(hp + 6) | ((uint32_t)__ldg(hp + 7) << 16);
                }
                asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(wdst + sbase + hh * HSTG), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w));
            }
        }
        cp_commit();
    };
    float acc[HG][NT][4];
#pragma unroll
    for (int hh = 0; hh < HG; ++hh)
#pragma unroll
        for (int n = 0; n < NT; ++n) acc[hh][n][0] = acc[hh][n][1] = acc[hh][n][2] = acc[hh][n][3] = 0.f;
    if (8 * NT > a.C) {
        for (int x = lane; x < 2 * HG * HSTG / 16; x += 32)
            asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(sm0 + 16 * x), "r"(0u));
        __syncwarp();
    }
    const int mi = lane >> 3;
    const uint32_t a_row = sm0 + XSTG + (((mi >> 1) << 3) + (lane & 7)) * WROWB + (mi & 1) * 16;
    const uint32_t b_row = sm0 + (((mi & 1) << 3) + (lane & 7)) * ROWB + (mi >> 1) * 16;
    int tile, ua, ub, ntile = 0, nua = -1, nub = -1;
    bool have = next(tile, ua, ub);
    if (have) stage(tile, ua, ub, 0);
    int it = 0;
    while (have) {
        const bool more = next(ntile, nua, nub);
        if (more) stage(ntile, nua, nub, (it + 1) & 1);
        if (more) cp_wait<1>(); else cp_wait<0>();
        __syncwarp();
        const uint32_t sbase = (it & 1) * (HG * HSTG);
#pragma unroll
        for (int hh = 0; hh < HG; ++hh) {
            uint32_t af[4];
            // A = W-block^T: matrices (tok 0-7 | keys o), (tok 0-7 | keys o+1), (tok 8-15 | keys o), (tok 8-15 | keys o+1)
            ldsm4t(af, a_row + sbase + hh * HSTG);
#pragma unroll
            for (int n = 0; n < NT; n += 2) {
                uint32_t bfr[4];
                ldsm4t(bfr, b_row + sbase + hh * HSTG + n * 16);
                mma16<T>(acc[hh][n], af[0], af[1], af[2], af[3], bfr[0], bfr[1]);
                mma16<T>(acc[hh][n + 1], af[0], af[1], af[2], af[3], bfr[2], bfr[3]);
            }
        }
        __syncwarp();
        have = more; tile = ntile; ua = nua; ub = nub; ++it;
    }
    T *Ob = opaque(reinterpret_cast<T *>(a.out) + b * a.o_sb + h0 * a.o_sh);
    const int ka = o * 8 + g, kb = ka + 8;
    {
        const int sta = ka < a.Nk ? 0 : -1, stb = kb < a.Nk ? 0 : -1;
        const int oa = ka * a.o_sn + 2 * t, ob = kb * a.o_sn + 2 * t;
#pragma unroll
        for (int hh = 0; hh < HG; ++hh)
#pragma unroll
            for (int n = 0; n < NT; ++n) {
                const int ok = 8 * n + 2 * t < a.C ? 0 : -1;
                st_pair_if<T>(at(Ob, hh * a.o_sh + oa + 8 * n), acc[hh][n][0], acc[hh][n][1], sta | ok);
                st_pair_if<T>(at(Ob, hh * a.o_sh + ob + 8 * n), acc[hh][n][2], acc[hh][n][3], stb | ok);
            }
    }
}

// Contributions of impure tokens (left out of the tile structure) to the key rows they reference: one (flagged row, head) per
// warp, grid-strided over pk.rimp_list.  The warp scans the row's full inverse neighbour list (csr.cu) 32 entries at a time,
// ballots the entries that come from impure tokens and adds them in ascending (i, j) order -> deterministic.  Runs AFTER
// the scatter kernel (same stream), which wrote the tile part of the row.  Exits at once when nothing is flagged.
template <typename T>
__global__ void __launch_bounds__(256)
scat_fixup_kernel(const ScatArgs a, const PackView pk) {
    if (pk.flags[0]) return;
    const int nr = min(pk.flags[5], pk.rimp_cap);
    if (nr == 0) return;
    const int lane = threadIdx.x & 31;
    const int n = nr * a.H;
    for (int it = blockIdx.x * 8 + (threadIdx.x >> 5); it < n; it += gridDim.x * 8) {
        const int rid = pk.rimp_list[it / a.H], h = it % a.H;
        const int b = rid / a.Nk, row = rid - b * a.Nk;
        const int lo = a.csr_off[(int64_t)b * (a.Nk + 1) + row], hi = a.csr_off[(int64_t)b * (a.Nk + 1) + row + 1];
        const uint32_t *cent = a.csr_ent + (int64_t)b * a.Nq * a.M;
        const uint8_t *timp = pk.tok_imp + (int64_t)b * pk.T * TILE_TOK;
        const T *Wh = reinterpret_cast<const T *>(a.W) + b * a.w_sb + h * a.w_sh;
        const T *Xh = reinterpret_cast<const T *>(a.X) + b * a.x_sb + h * a.x_sh;
        float sum = 0.f;                                            // channel `lane`
        for (int e0 = lo; e0 < hi; e0 += 32) {
            const int e = e0 + lane;
            uint32_t pkd = 0;
            bool imp = false;
            if (e < hi) { pkd = __ldg(cent + e); imp = timp[pkd >> 8] != 0; }
            for (unsigned m = __ballot_sync(FULL, imp); m; m &= m - 1) {
                const uint32_t q = __shfl_sync(FULL, pkd, __ffs(m) - 1);
                const int i = (int)(q >> 8), j = (int)(q & 255u);
                if (lane < a.C) sum = fmaf(to_f(Wh[i * a.w_sn + j]), to_f(Xh[i * a.x_sn + lane]), sum);
            }
        }
        if (lane < a.C) {
            T *op = reinterpret_cast<T *>(a.out) + b * a.o_sb + h * a.o_sh + row * a.o_sn + lane;
            *op = from_f<T>(to_f(*op) + sum);
        }
    }
}

}  // namespace t2

static inline int head_group(int H) { return H % 4 == 0 ? 4 : H % 3 == 0 ? 3 : H % 2 == 0 ? 2 : 1; }
static inline bool fits31(int64_t v) { return v >= 0 && v < (1LL << 31); }

// 0 = launched, -1 = not applicable (caller uses the first-generation kernel), > 0 = CUDA error
template <typename T>
int launch_dot_tile2(const T *X, const T *Y, const int64_t *idx, const void *pack, T *out, int B, int H, int Nq, int Nk, int C,
                     int M, Rows4 x, Rows4 y, cudaStream_t st) {
    if constexpr (sizeof(T) != 2) return -1;
    else {
        const PackView pk = pack_view(const_cast<void *>(pack), B, Nq, Nk);
        if (!fits31((int64_t)H * x.sh) || !fits31((int64_t)Nq * x.sn + (int64_t)H * x.sh) || !fits31((int64_t)Nk * y.sn + (int64_t)H * y.sh) ||
            !fits31((int64_t)H * Nq * M) || !fits31((int64_t)B * pk.T * TILE_TOK * U_MAX) || x.sh < 0 || x.sn < 0 || y.sh < 0 || y.sn < 0)
            return -1;
        const int HG = head_group(H);
        if ((int64_t)B * pk.T == 0) return 0;
        if ((int64_t)B * (H / HG) > 65535) return -1;
        const dim3 grid(ceil_div(pk.T, t2::TW2), B * (H / HG));
        t2::DotArgs a{X, Y, idx, out, B, H, Nq, C, M, (int)x.sh, (int)x.sn, (int)y.sh, (int)y.sn, x.sb, y.sb};
#define DOT2(CH_, HG_, UB_) t2::dot2_kernel16<T, CH_, HG_, UB_><<<grid, t2::TW2 * 32, 0, st>>>(a, pk)
        if (C <= 16) {
            switch (HG) { case 4: DOT2(4, 4, 4); break; case 3: DOT2(4, 3, 4); break; case 2: DOT2(4, 2, 4); break; default: DOT2(4, 1, 4); }
        } else {
            switch (HG) { case 4: DOT2(8, 4, 2); break; case 3: DOT2(8, 3, 2); break; case 2: DOT2(8, 2, 4); break; default: DOT2(8, 1, 4); }
        }
#undef DOT2
        note_launches(1);
        return check_launch("dot_tile2");
    }
}

template <typename T>
int launch_axpy_tile2(const T *W, const T *Y, const int64_t *idx, const void *pack, T *out, int B, int H, int Nq, int Nk, int C,
                      int M, Rows4 w, Rows4 y, Rows4 o, cudaStream_t st) {
    if constexpr (sizeof(T) != 2) return -1;
    else {
        const PackView pk = pack_view(const_cast<void *>(pack), B, Nq, Nk);
        if (!fits31((int64_t)Nq * w.sn + (int64_t)H * w.sh) || !fits31((int64_t)Nk * y.sn + (int64_t)H * y.sh) ||
            !fits31((int64_t)Nq * o.sn + (int64_t)H * o.sh) || !fits31((int64_t)B * pk.T * TILE_TOK * U_MAX) ||
            w.sh < 0 || w.sn < M || y.sh < 0 || y.sn < 0 || o.sh < 0 || o.sn < 0)
            return -1;
        // the W tile is copied as one 16-byte-aligned contiguous block per (head, tile)
        if (!aligned16(w.p) || w.sb % 8 || w.sh % 8 || w.sn > 1024) return -1;
        const int NT = C <= 16 ? 2 : 4;
        const int wtile = 32 * (int)w.sn;
        int HG = 0;
        for (int cand = 4; cand >= 1; --cand)
            if (H % cand == 0 && 2 * cand * 16 * (NT * 16 + 16) + cand * wtile <= 16 * 1024) { HG = cand; break; }
        if (!HG || (int64_t)B * (H / HG) > 65535) return -1;
        if ((int64_t)B * pk.T == 0) return 0;
        const int spw = 2 * HG * 16 * (NT * 16 + 16) + HG * wtile;
        const int WPC = 4;
        const dim3 grid(ceil_div(pk.T, WPC), B * (H / HG));
        const size_t smem = (size_t)spw * WPC;
        const bool al4 = w.sn % 2 == 0;
        t2::AxpyArgs a{W, Y, idx, out, B, H, Nq, C, M, (int)w.sh, (int)w.sn, (int)y.sh, (int)y.sn, (int)o.sh, (int)o.sn, w.sb, y.sb, o.sb, spw, wtile};
#define AXPY2_(NT_, HG_, AL_) do { auto kfn = t2::axpy2_kernel16<T, NT_, HG_, AL_>; \
            if (smem > 48 * 1024) cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            kfn<<<grid, WPC * 32, smem, st>>>(a, pk); } while (0)
#define AXPY2(NT_, HG_) do { if (al4) AXPY2_(NT_, HG_, true); else AXPY2_(NT_, HG_, false); } while (0)
        if (NT == 2) {
            switch (HG) { case 4: AXPY2(2, 4); break; case 3: AXPY2(2, 3); break; case 2: AXPY2(2, 2); break; default: AXPY2(2, 1); }
        } else {
            switch (HG) { case 4: AXPY2(4, 4); break; case 3: AXPY2(4, 3); break; case 2: AXPY2(4, 2); break; default: AXPY2(4, 1); }
        }
#undef AXPY2
#undef AXPY2_
        note_launches(1);
        return check_launch("axpy_tile2");
    }
}

template int launch_axpy_tile2<float>(const float *, const float *, const int64_t *, const void *, float *, int, int, int, int, int, int, Rows4, Rows4, Rows4, cudaStream_t);
template int launch_axpy_tile2<__half>(const __half *, const __half *, const int64_t *, const void *, __half *, int, int, int, int, int, int, Rows4, Rows4, Rows4, cudaStream_t);
template int launch_axpy_tile2<__nv_bfloat16>(const __nv_bfloat16 *, const __nv_bfloat16 *, const int64_t *, const void *, __nv_bfloat16 *, int, int, int, int, int, int, Rows4, Rows4, Rows4, cudaStream_t);

template <typename T>
int launch_scat_tile2(const T *W, const T *X, const int32_t *csr_off, const uint32_t *csr_ent, const void *pack, T *out,
                      int B, int H, int Nq, int Nk, int C, int M, Rows4 w, Rows4 x, Rows4 o, cudaStream_t st) {
    if constexpr (sizeof(T) != 2) return -1;
    else {
        const PackView pk = pack_view(const_cast<void *>(pack), B, Nq, Nk);
        if (!fits31((int64_t)Nq * w.sn + (int64_t)H * w.sh) || !fits31((int64_t)Nq * x.sn + (int64_t)H * x.sh) ||
            !fits31((int64_t)Nk * o.sn + (int64_t)H * o.sh) || !fits31((int64_t)pk.T * TILE_TOK * U_MAX) ||
            w.sh < 0 || w.sn < 0 || x.sh < 0 || x.sn < 0 || o.sh < 0 || o.sn < 0)
            return -1;
        const int NT = C <= 16 ? 2 : 4;
        const int hstg = 16 * (NT * 16 + 16) + 16 * 48;
        int HG = 0;
        for (int cand = 4; cand >= 1; --cand)
            if (H % cand == 0) { HG = cand; break; }
        if ((int64_t)B * (H / HG) > 65535) return -1;
        const int NP = (pk.NO + 1) / 2;
        if ((int64_t)B * NP == 0) return 0;
        const int spw = 2 * HG * hstg;
        const int WPC = 4;
        const dim3 grid(ceil_div(NP, WPC), B * (H / HG));
        const size_t smem = (size_t)spw * WPC;
        const bool wal = aligned16(w.p) && w.sb % 8 == 0 && w.sh % 8 == 0 && w.sn % 8 == 0;
        t2::ScatArgs a{W, X, csr_off, csr_ent, out, B, H, Nq, Nk, C, M, (int)w.sh, (int)w.sn, (int)x.sh, (int)x.sn, (int)o.sh, (int)o.sn,
                       w.sb, x.sb, o.sb, spw};
#define SCAT2_(NT_, HG_, AL_) do { auto kfn = t2::scat2_kernel16<T, NT_, HG_, AL_>; \
            if (smem > 48 * 1024) cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            kfn<<<grid, WPC * 32, smem, st>>>(a, pk); } while (0)
#define SCAT2(NT_, HG_) do { if (wal) SCAT2_(NT_, HG_, true); else SCAT2_(NT_, HG_, false); } while (0)
        if (NT == 2) {
            switch (HG) { case 4: SCAT2(2, 4); break; case 3: SCAT2(2, 3); break; case 2: SCAT2(2, 2); break; default: SCAT2(2, 1); }
        } else {
            switch (HG) { case 4: SCAT2(4, 4); break; case 3: SCAT2(4, 3); break; case 2: SCAT2(4, 2); break; default: SCAT2(4, 1); }
        }
#undef SCAT2
#undef SCAT2_
        t2::scat_fixup_kernel<T><<<148 * 2, 256, 0, st>>>(a, pk);     // impure tokens' contributions (exits at once when none)
        note_launches(2);
        return check_launch("scat_tile2");
    }
}

#define INST_SCAT2(T) template int launch_scat_tile2<T>(const T *, const T *, const int32_t *, const uint32_t *, const void *, T *, int, int, int, int, int, int, Rows4, Rows4, Rows4, cudaStream_t);
INST_SCAT2(float) INST_SCAT2(__half) INST_SCAT2(__nv_bfloat16)
#undef INST_SCAT2

template int launch_dot_tile2<float>(const float *, const float *, const int64_t *, const void *, float *, int, int, int, int, int, int, Rows4, Rows4, cudaStream_t);
template int launch_dot_tile2<__half>(const __half *, const __half *, const int64_t *, const void *, __half *, int, int, int, int, int, int, Rows4, Rows4, cudaStream_t);
template int launch_dot_tile2<__nv_bfloat16>(const __nv_bfloat16 *, const __nv_bfloat16 *, const int64_t *, const void *, __nv_bfloat16 *, int, int, int, int, int, int, Rows4, Rows4, cudaStream_t);

}  // namespace clusten
