// Tile pack: the per-index-tensor structure behind the tensor-core ("tile-union") kernels of clusten_tile.cu.
//
// Tokens are taken 16 at a time (one mma M-tile).  Key rows are taken 8 at a time ("octets": rows 8o..8o+7 -- a
// balanced cluster of the reference is m = 8 or 24 consecutive rows, point_utils.py:282-285, so a neighbourhood is
// M/8 octets).  For every tile the pack holds the UNION of octets its 16 tokens reference and, per (token, union
// position), the neighbour slot (j / 8) at which the token references that octet, or -1.  A slot is "pure" when
// idx[i, 8s + r] == 8o + r for r = 0..7.  A token with an impure slot (padded tail of the last cluster,
// point_utils.py:282-283, or an arbitrary index tensor) is an "impure token": the tile kernels leave it out of the
// tensor-core work and handle it whole in a slow in-kernel path (tok_imp / row_imp below).  Too many impure tokens,
// or a tile whose union exceeds U_MAX, switch the WHOLE tensor to the generic kernels through the device-side flag (no
// host synchronisation: both kernels are enqueued, one of them exits at once).
#pragma once
#include <algorithm>

#include "common.cuh"

namespace clusten {

constexpr int TILE_TOK = 16;     // tokens per tile (mma M)
constexpr int U_MAX = 48;        // max union octets per tile (measured: <= 18 for m = 8, <= 39 for m = 24 / M = 144)
constexpr int S_MAX = 32;        // max slots per token (M <= 256)

struct PackView {
    int *flags;          // [0] != 0 -> generic path; [1] max U seen; [2] impure tokens; [3] tiles over U_MAX;
                         // [4] != 0 -> the full inverse neighbour list (csr.cu) is needed (generic path or impure tokens)
    int *tile_u;         // [B*T]
    int *tile_oct;       // [B*T*U_MAX]
    int8_t *slot_of;     // [B*T*16*U_MAX]  slot of (token row, union position) or -1
    int8_t *slot_t;      // [B*T*U_MAX*16]  the same table transposed: the 16 token slots of one (tile, union position) are
                         //                 16 contiguous bytes at (tile*U_MAX + u)*16 = inverse-list entry * 16 (scatter kernels)
    int *oct_off;        // [B*(NO+1)]      inverse lists: for key octet o the (tile, u) pairs referencing it ...
    uint32_t *oct_ent;   // [B*T*U_MAX]     ... entry = tile*U_MAX + u, ascending tile order
    uint8_t *tok_imp;    // [B*T*16]        1 = impure token (handled by the slow in-kernel path)
    uint8_t *row_imp;    // [B*Nk]          1 = key row referenced by an impure token (scatter kernels fix it up)
    int *imp_list;       // [imp_cap]       global ids (b*Nq + i) of the impure tokens, any order; count = min(flags[2], imp_cap)
    int *rimp_list;      // [rimp_cap]      global ids (b*Nk + r) of the key rows with row_imp set (each once); count = min(flags[5], rimp_cap)
    int T, NO, imp_cap, rimp_cap;
};

struct Rows4 { const void *p; int64_t sb, sh, sn; };   // strided [B,H,N,C] operand: element strides, unit inner stride

struct PackLayout {
    size_t flags, tile_u, tile_oct, slot_of, slot_t, oct_off, oct_ent, tok_imp, row_imp, imp_list, rimp_list, sort_ws,
           grp_u, grp_oct, sub_pos, slot_g, total;
    int T, NO, imp_cap, rimp_cap, TG;
};

// Tile GROUPS (clusten_fused_tma.cu): GROUP_TILES consecutive tiles = the 64 tokens one CTA of the TMA-staged kernels owns.
// Per group the union of its tiles' unions (first-seen order: a CTA stages these octets in shared memory once, by TMA box
// loads) and, per (tile, union position), the position of that octet in the group's list.
constexpr int GROUP_TILES = 4;
constexpr int GU_MAX = GROUP_TILES * U_MAX;     // a group's union can never exceed the sum of its tiles' unions
constexpr int SLOT_G_ROW = 2 * U_MAX + 8;       // bytes of one row of slot_g (8 bytes of padding: 64-bit reads of 8 rows hit 16 different banks)
constexpr int SLOT_G_TILE = 8 * SLOT_G_ROW;     // bytes of one tile's slot_g block (a multiple of 16)

struct GroupView {
    int *grp_u;          // [B*TG]            octets in the group's union
    int *grp_oct;        // [B*TG*GU_MAX]     their ids, first-seen order over the group's tiles
    uint8_t *sub_pos;    // [B*T*U_MAX]       position in the group's list of (tile, union position u); 0 beyond the tile's union
    uint8_t *slot_g;     // [B*T][8][SLOT_G_ROW]  the slot table in mma-fragment order: row g, entry u (16 bits) = slot of token row g
                         //                   in the low byte, of token row g + 8 in the high byte (0xff = none)
    int TG;              // groups per batch sample = ceil(T / GROUP_TILES);  flags[6] = largest group union
};

inline size_t pack_align(size_t x) { return (x + 255) & ~(size_t)255; }

inline PackLayout pack_layout(int B, int Nq, int Nk) {
    PackLayout L;
    L.T = (Nq + TILE_TOK - 1) / TILE_TOK;
    L.NO = (Nk + 7) / 8;
    const size_t bt = (size_t)B * L.T;
    size_t o = 0;
    L.flags = o;    o += 256;
    L.tile_u = o;   o += pack_align(bt * 4);
    L.tile_oct = o; o += pack_align(bt * U_MAX * 4);
    L.slot_of = o;  o += pack_align(bt * TILE_TOK * U_MAX);
    L.slot_t = o;   o += pack_align(bt * TILE_TOK * U_MAX);
    L.oct_off = o;  o += pack_align((size_t)B * (L.NO + 1) * 4);
    L.oct_ent = o;  o += pack_align(bt * U_MAX * 4);
    L.tok_imp = o;  o += pack_align(bt * TILE_TOK);
    L.row_imp = o;  o += pack_align((size_t)B * Nk);
    // the tile path is only taken while impure tokens <= max(64, tokens / 16) (pack_decide_kernel): the lists never need more
    L.imp_cap = (int)std::min<int64_t>((int64_t)B * Nq, (int64_t)B * Nq / 16 + 64);
    L.rimp_cap = (int)std::min<int64_t>((int64_t)B * Nk, (int64_t)1 << 30);
    L.imp_list = o; o += pack_align((size_t)L.imp_cap * 4);
    L.rimp_list = o; o += pack_align((size_t)L.rimp_cap * 4);
    L.sort_ws = o;  o += 5 * pack_align((size_t)L.T * U_MAX * B * 4) + radix_sort_workspace_bytes(B, L.T * U_MAX) + 256;
    L.TG = (L.T + GROUP_TILES - 1) / GROUP_TILES;
    L.grp_u = o;    o += pack_align((size_t)B * L.TG * 4);
    L.grp_oct = o;  o += pack_align((size_t)B * L.TG * GU_MAX * 4);
    L.sub_pos = o;  o += pack_align(bt * U_MAX);
    L.slot_g = o;   o += pack_align(bt * SLOT_G_TILE);
    L.total = o;
    return L;
}

inline PackView pack_view(void *buf, int B, int Nq, int Nk) {
    const PackLayout L = pack_layout(B, Nq, Nk);
    char *p = reinterpret_cast<char *>(buf);
    PackView v;
    v.flags = reinterpret_cast<int *>(p + L.flags);
    v.tile_u = reinterpret_cast<int *>(p + L.tile_u);
    v.tile_oct = reinterpret_cast<int *>(p + L.tile_oct);
    v.slot_of = reinterpret_cast<int8_t *>(p + L.slot_of);
    v.slot_t = reinterpret_cast<int8_t *>(p + L.slot_t);
    v.oct_off = reinterpret_cast<int *>(p + L.oct_off);
    v.oct_ent = reinterpret_cast<uint32_t *>(p + L.oct_ent);
    v.tok_imp = reinterpret_cast<uint8_t *>(p + L.tok_imp);
    v.row_imp = reinterpret_cast<uint8_t *>(p + L.row_imp);
    v.imp_list = reinterpret_cast<int *>(p + L.imp_list);
    v.rimp_list = reinterpret_cast<int *>(p + L.rimp_list);
    v.imp_cap = L.imp_cap;
    v.rimp_cap = L.rimp_cap;
    v.T = L.T;
    v.NO = L.NO;
    return v;
}

inline GroupView group_view(void *buf, int B, int Nq, int Nk) {
    const PackLayout L = pack_layout(B, Nq, Nk);
    char *p = reinterpret_cast<char *>(buf);
    GroupView g;
    g.grp_u = reinterpret_cast<int *>(p + L.grp_u);
    g.grp_oct = reinterpret_cast<int *>(p + L.grp_oct);
    g.sub_pos = reinterpret_cast<uint8_t *>(p + L.sub_pos);
    g.slot_g = reinterpret_cast<uint8_t *>(p + L.slot_g);
    g.TG = L.TG;
    return g;
}

}  // namespace clusten
