// clusten_pack_build: analyse an index tensor once (per AFF stage) and emit the tile pack of tile.cuh.
#include "tile.cuh"

namespace clusten {

constexpr int PACK_WARPS = 4;

__global__ void __launch_bounds__(PACK_WARPS * 32)
pack_tile_kernel(const int64_t *__restrict__ idx, const uint8_t *__restrict__ mask, int B, int Nq, int M, int Nk, PackView pk, GroupView gv) {
    __shared__ int oct_rs[PACK_WARPS][TILE_TOK][S_MAX];
    __shared__ __align__(16) int8_t slot_s[PACK_WARPS][TILE_TOK][U_MAX];
    __shared__ int row_bad[PACK_WARPS][TILE_TOK];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bt = blockIdx.x * PACK_WARPS + warp;
    if (bt >= B * pk.T) return;
    const int b = bt / pk.T, tile = bt - b * pk.T;
    const int i0 = tile * TILE_TOK;
    const int S = M >> 3;
    // step 1: every (row, slot) -> octet id, -1 (impure) or -2 (row beyond Nq)
    for (int item = lane; item < TILE_TOK * S; item += 32) {
        const int r = item / S, s = item - r * S;
        const int i = i0 + r;
        int o = -2;
        if (i < Nq) {
            const int64_t *p = idx + ((int64_t)b * Nq + i) * M + 8 * s;
            if (!mask) {
                const int64_t base = p[0];
                bool pure = base >= 0 && (base & 7) == 0 && base + 7 < (int64_t)Nk;
#pragma unroll
                for (int k = 1; k < 8; ++k) pure = pure && (p[k] == base + k);
                o = pure ? (int)(base >> 3) : -1;
            } else {
                // mask-aware (fused attention only): masked entries are wildcards -- their probability is exp(-100 - ...) = 0
                // whatever key row they read -- so a padded cluster (point_utils.py:282-283) is a pure, possibly partial, octet
                const uint8_t *mk = mask + ((int64_t)b * Nq + i) * M + 8 * s;
                int64_t base = -1;
                bool pure = true;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (!mk[k]) continue;
                    if (base < 0) base = p[k] - k;
                    pure = pure && p[k] == base + k && p[k] < (int64_t)Nk;
                }
                if (base < 0) o = -3;                        // fully masked slot: resolved below (continues the previous slot's run)
                else {
                    pure = pure && (base & 7) == 0 && base < (int64_t)Nk;
                    o = pure ? (int)(base >> 3) : -1;
                }
            }
        }
        oct_rs[warp][r][s] = o;
    }
    for (int x = lane; x < TILE_TOK * U_MAX / 4; x += 32) reinterpret_cast<int *>(&slot_s[warp][0][0])[x] = -1;
    if (lane < TILE_TOK) row_bad[warp][lane] = 0;
    __syncwarp();
    if (mask && lane < TILE_TOK) {
        // a fully masked slot (the all-padding octets of a padded cluster of m = 24) takes the octet after its predecessor's:
        // a distinct id, possibly beyond the last real octet -- its rows are clamped by the kernels, its weights are ~ 0
        for (int s2 = 0; s2 < S; ++s2)
            if (oct_rs[warp][lane][s2] == -3) oct_rs[warp][lane][s2] = (s2 > 0 && oct_rs[warp][lane][s2 - 1] >= 0) ? oct_rs[warp][lane][s2 - 1] + 1 : -1;
    }
    __syncwarp();
    for (int item = lane; item < TILE_TOK * S; item += 32)
        if (oct_rs[warp][item / S][item % S] == -1) row_bad[warp][item / S] = 1;      // benign race: all writers store 1
    __syncwarp();
    // step 2: union in first-seen order (lane u holds union position u / u + 32)
    int my0 = -1, my1 = -1, U = 0;
    for (int r = 0; r < TILE_TOK; ++r) {
        if (row_bad[warp][r]) continue;                      // impure token: contributes nothing to the union
        for (int s = 0; s < S; ++s) {
            const int o = oct_rs[warp][r][s];
            if (o < 0) continue;
            const unsigned m0 = __ballot_sync(FULL, my0 == o), m1 = __ballot_sync(FULL, my1 == o);
            int pos;
            if (m0) pos = __ffs(m0) - 1;
            else if (m1) pos = 32 + __ffs(m1) - 1;
            else {
                pos = U;
                if (U < 32) { if (lane == U) my0 = o; }
                else if (U < 64) { if (lane == U - 32) my1 = o; }
                ++U;
            }
            if (pos < U_MAX) {
                const int prev = slot_s[warp][r][pos];
                __syncwarp();
                if (prev != -1) { if (lane == 0) row_bad[warp][r] = 1; }   // one octet twice: impure token
                else if (lane == 0) slot_s[warp][r][pos] = (int8_t)s;
                __syncwarp();
            }
        }
    }
    __syncwarp();
    // impure tokens: no slots in the tile structure, flagged for the slow paths, their key rows flagged for the scatter
    int bad = 0;
    unsigned bad_rows = 0;
    for (int r = 0; r < TILE_TOK; ++r) {
        const int rb_ = row_bad[warp][r];
        if (lane == 0) pk.tok_imp[(int64_t)bt * TILE_TOK + r] = (uint8_t)rb_;
        if (!rb_) continue;
        if (lane == 0) bad_rows |= 1u << r;
        ++bad;
        for (int u = lane; u < U_MAX; u += 32) slot_s[warp][r][u] = -1;
        const int64_t *p = idx + ((int64_t)b * Nq + i0 + r) * M;
        for (int j = lane; j < M; j += 32) {
            const int64_t v = p[j];
            if (v >= 0 && v < (int64_t)Nk) {
                // first marker of a key row also appends it to the flagged-row list (byte-wide exchange through the aligned word)
                unsigned *wp = reinterpret_cast<unsigned *>(pk.row_imp + (((int64_t)b * Nk + v) & ~(int64_t)3));
                const unsigned bit = 1u << (8 * (((int64_t)b * Nk + v) & 3));
                const unsigned old = atomicOr(wp, bit);
                if (!(old & bit)) {
                    const int pos = atomicAdd(pk.flags + 5, 1);
                    if (pos < pk.rimp_cap) pk.rimp_list[pos] = (int)((int64_t)b * Nk + v);
                }
            }
        }
    }
    __syncwarp();
    const int Uc = min(U, U_MAX);
    pk.tile_oct[(int64_t)bt * U_MAX + lane] = lane < Uc ? my0 : 0;
    if (lane + 32 < U_MAX) pk.tile_oct[(int64_t)bt * U_MAX + 32 + lane] = (lane + 32 < Uc) ? my1 : 0;
    const int4 *src = reinterpret_cast<const int4 *>(&slot_s[warp][0][0]);
    int4 *dst = reinterpret_cast<int4 *>(pk.slot_of + (int64_t)bt * TILE_TOK * U_MAX);
    for (int x = lane; x < TILE_TOK * U_MAX / 16; x += 32) dst[x] = src[x];
    // transposed copy [u][token] for the scatter kernels: one 16-byte vector per (tile, union position)
    for (int u = lane; u < U_MAX; u += 32) {
        uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int r = 0; r < TILE_TOK; ++r) w[r >> 2] |= (uint32_t)(uint8_t)slot_s[warp][r][u] << (8 * (r & 3));
        *reinterpret_cast<uint4 *>(pk.slot_t + ((int64_t)bt * U_MAX + u) * 16) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    // the same table in mma-fragment order (rows g and g + 8 of one union position side by side) for the TMA-staged kernels
    for (int x = lane; x < 8 * U_MAX; x += 32) {
        const int gr = x / U_MAX, u = x - gr * U_MAX;
        const uint16_t v = (uint16_t)(uint8_t)slot_s[warp][gr][u] | (uint16_t)((uint16_t)(uint8_t)slot_s[warp][gr + 8][u] << 8);
        *reinterpret_cast<uint16_t *>(gv.slot_g + (int64_t)bt * SLOT_G_TILE + gr * SLOT_G_ROW + 2 * u) = v;
    }
    if (lane == 0) {
        pk.tile_u[bt] = Uc;
        atomicMax(pk.flags + 1, U);
        if (bad) {
            int pos = atomicAdd(pk.flags + 2, bad);
            for (; bad_rows; bad_rows &= bad_rows - 1, ++pos)
                if (pos < pk.imp_cap) pk.imp_list[pos] = b * Nq + i0 + __ffs(bad_rows) - 1;
        }
        if (U > U_MAX) atomicAdd(pk.flags + 3, 1);
    }
}

// union of GROUP_TILES consecutive tiles' unions, first-seen order (one warp per group; lane l holds list entries l, l+32, ...)
__global__ void __launch_bounds__(PACK_WARPS * 32)
pack_group_kernel(PackView pk, GroupView gv, int B) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bg = blockIdx.x * PACK_WARPS + warp;
    if (bg >= B * gv.TG) return;
    const int b = bg / gv.TG, grp = bg - b * gv.TG;
    constexpr int PER_LANE = GU_MAX / 32;
    int mine[PER_LANE];
#pragma unroll
    for (int x = 0; x < PER_LANE; ++x) mine[x] = -1;
    int GU = 0;
    for (int tl = 0; tl < GROUP_TILES; ++tl) {
        const int tile = grp * GROUP_TILES + tl;
        if (tile >= pk.T) break;
        const int64_t bt = (int64_t)b * pk.T + tile;
        const int U = pk.tile_u[bt];
        for (int u = 0; u < U; ++u) {
            const int o = pk.tile_oct[bt * U_MAX + u];
            int pos = -1;
#pragma unroll
            for (int x = 0; x < PER_LANE; ++x) {
                const unsigned hit = __ballot_sync(FULL, mine[x] == o);
                if (hit && pos < 0) pos = 32 * x + __ffs(hit) - 1;
            }
            if (pos < 0) {
                pos = GU++;
#pragma unroll
                for (int x = 0; x < PER_LANE; ++x)
                    if (pos >> 5 == x && lane == (pos & 31)) mine[x] = o;
            }
            if (lane == 0) gv.sub_pos[bt * U_MAX + u] = (uint8_t)pos;
        }
        for (int u = U + lane; u < U_MAX; u += 32) gv.sub_pos[bt * U_MAX + u] = 0;      // (positions beyond the union: a staged octet)
    }
#pragma unroll
    for (int x = 0; x < PER_LANE; ++x) gv.grp_oct[(int64_t)bg * GU_MAX + 32 * x + lane] = (32 * x + lane < GU) ? mine[x] : 0;
    if (lane == 0) {
        gv.grp_u[bg] = GU;
        atomicMax(pk.flags + 6, GU);
    }
}

// generic path when a union overflowed or more than 1/16 of the tokens (and more than 64) are impure
__global__ void pack_decide_kernel(int *flags, int tokens) {
    const bool generic = flags[3] > 0 || (flags[2] > 64 && flags[2] > tokens / 16);
    flags[0] = generic ? 1 : 0;
    flags[4] = (generic || flags[2] > 0) ? 1 : 0;
}

// inverse lists: one (key = octet, implicit value = tile*U_MAX + u) pair per union entry, NO = sentinel for padding
__global__ void pack_inv_keys_kernel(PackView pk, int B, uint32_t *__restrict__ keys) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t per_b = (int64_t)pk.T * U_MAX;
    if (p >= (int64_t)B * per_b) return;
    const int64_t bt = p / U_MAX;
    const int u = (int)(p - bt * U_MAX);
    keys[p] = u < pk.tile_u[bt] ? (uint32_t)min(pk.tile_oct[p], pk.NO) : (uint32_t)pk.NO;     // (octets past the last real one: no key rows)
}

__global__ void pack_inv_finalize_kernel(const uint32_t *__restrict__ skeys, const uint32_t *__restrict__ svals,
                                         PackView pk, int nseg) {
    const int b = blockIdx.y;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nseg) return;
    const uint32_t *k = skeys + (int64_t)b * nseg;
    int *off = pk.oct_off + (int64_t)b * (pk.NO + 1);
    const int key = (int)min(k[p], (uint32_t)pk.NO);
    const int prev = p ? (int)min(k[p - 1], (uint32_t)pk.NO) : -1;
    for (int r = prev + 1; r <= key; ++r) off[r] = p;
    if (p == nseg - 1)
        for (int r = key + 1; r <= pk.NO; ++r) off[r] = nseg;
    pk.oct_ent[(int64_t)b * nseg + p] = svals[(int64_t)b * nseg + p];
}

__global__ void pack_disable_kernel(int *flags) { flags[0] = 1; flags[4] = 1; }

}  // namespace clusten

using namespace clusten;

extern "C" size_t clusten_pack_bytes(int B, int Nq, int M, int Nk) {
    (void)M;
    if (B <= 0 || Nq <= 0 || Nk <= 0) return 256;
    return pack_layout(B, Nq, Nk).total;
}

extern "C" int clusten_pack_build(const int64_t *nbhd_idx, const uint8_t *mask, int B, int Nq, int M, int Nk, void *pack,
                                  size_t pack_bytes, void *stream) {
    if (B < 0 || Nq < 0 || M <= 0 || Nk <= 0) return set_error(CLUSTEN_EINVAL, "bad sizes B=%d Nq=%d M=%d Nk=%d", B, Nq, M, Nk);
    if (!nbhd_idx || !pack) return set_error(CLUSTEN_EINVAL, "null pointer");
    if (pack_bytes < clusten_pack_bytes(B, Nq, M, Nk))
        return set_error(CLUSTEN_EWORKSPACE, "pack buffer too small: %zu < %zu", pack_bytes, clusten_pack_bytes(B, Nq, M, Nk));
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(pack, 0, 256, st);
    if (B == 0 || Nq == 0) return check_launch("pack memset");
    PackView pk = pack_view(pack, B, Nq, Nk);
    if ((M & 7) || (M >> 3) > S_MAX || (int64_t)pk.T * U_MAX >= (1LL << 31) / 2) {
        pack_disable_kernel<<<1, 1, 0, st>>>(pk.flags);       // no octet structure: generic kernels only
        note_launches(1);
        return check_launch("pack_disable");
    }
    const int bt = B * pk.T;
    cudaMemsetAsync(pk.row_imp, 0, (size_t)B * Nk, st);
    const GroupView gv = group_view(pack, B, Nq, Nk);
    pack_tile_kernel<<<ceil_div(bt, PACK_WARPS), PACK_WARPS * 32, 0, st>>>(nbhd_idx, mask, B, Nq, M, Nk, pk, gv);
    pack_decide_kernel<<<1, 1, 0, st>>>(pk.flags, B * Nq);
    pack_group_kernel<<<ceil_div(B * gv.TG, PACK_WARPS), PACK_WARPS * 32, 0, st>>>(pk, gv, B);
    note_launches(3);
    return check_launch("pack_build");
}

extern "C" int clusten_pack_inverse(void *pack, size_t pack_bytes, int B, int Nq, int M, int Nk, void *stream) {
    if (B < 0 || Nq < 0 || M <= 0 || Nk <= 0) return set_error(CLUSTEN_EINVAL, "bad sizes B=%d Nq=%d M=%d Nk=%d", B, Nq, M, Nk);
    if (!pack) return set_error(CLUSTEN_EINVAL, "null pointer");
    if (pack_bytes < clusten_pack_bytes(B, Nq, M, Nk)) return set_error(CLUSTEN_EWORKSPACE, "pack buffer too small");
    cudaStream_t st = (cudaStream_t)stream;
    if (B == 0 || Nq == 0) return 0;
    PackView pk = pack_view(pack, B, Nq, Nk);
    if ((M & 7) || (M >> 3) > S_MAX || (int64_t)pk.T * U_MAX >= (1LL << 31) / 2) return 0;   // generic-only pack
    // inverse lists (key octet -> referencing (tile, u)), stable sort keeps ascending tile order -> deterministic sums
    const PackLayout L = pack_layout(B, Nq, Nk);
    const int nseg = pk.T * U_MAX;
    const size_t stride = pack_align((size_t)nseg * B * 4);
    char *ws = reinterpret_cast<char *>(pack) + L.sort_ws;
    uint32_t *kA = reinterpret_cast<uint32_t *>(ws);
    uint32_t *kB = reinterpret_cast<uint32_t *>(ws + stride);
    uint32_t *vA = reinterpret_cast<uint32_t *>(ws + 2 * stride);
    uint32_t *vB = reinterpret_cast<uint32_t *>(ws + 3 * stride);
    uint32_t *kC = reinterpret_cast<uint32_t *>(ws + 4 * stride);
    void *hist = ws + 5 * stride;
    pack_inv_keys_kernel<<<ceil_div((int64_t)B * nseg, 256), 256, 0, st>>>(pk, B, kA);
    int bits = 1;
    while ((1LL << bits) <= pk.NO) ++bits;
    if (int e = radix_sort_pairs(kA, nullptr, kB, vB, kC, vA, B, nseg, bits, hist, st)) return e;
    pack_inv_finalize_kernel<<<dim3(ceil_div(nseg, 256), B), 256, 0, st>>>(kC, vA, pk, nseg);
    note_launches(2);
    return check_launch("pack_inverse");
}
