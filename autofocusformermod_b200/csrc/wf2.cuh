// WF plan: the per-index-tensor structure behind the tensor-core WF kernels of clusten_wf2.cu (16-bit types).
//
// The PointConv merge (aff.py:332-361) gathers, for each of the N' kept tokens, M = 48 (or 144) rows of the [B,N,C]
// feature map: M/8 "octets" (rows 8o..8o+7 -- the members of one balanced cluster, point_utils.py:282-285).  The kept
// tokens arrive in top-k order (aff.py:320-324), i.e. spatially shuffled, so consecutive tokens share no rows.
// The plan holds, per sample,
//   perm      the tokens ordered by their first neighbour octet (a counting sort; curve order ~ spatial order), so that
//             the tokens one SM works on back to back re-use each other's rows out of L1 instead of L2;
//   oct_off / oct_ent   for every feature octet the (token, slot) pairs referencing it, ascending -- the backward
//             scatter d_f becomes one dense [channels x 4*entries] x [4*entries x 8 rows] product per octet;
//   imp_*     the slots that are not a pure octet (padded tail of the last cluster, point_utils.py:282-283): their rows
//             ("flagged rows") are recomputed whole by a small fix-up kernel;
//   flags[0]  != 0 when the index tensor has no octet structure worth using (too many impure slots, an octet with
//             too many references): d_f then takes the generic inverse-list kernels.  Decided on the device; flags[4]
//             mirrors it at the position clusten_csr_build reads its device-side skip flag from (tile.cuh).
#pragma once
#include "common.cuh"

namespace clusten {

constexpr int WFP_IMP_CAP = 256;       // impure slots per sample handled by the fix-up kernel
constexpr int WFP_LIST_MAX = 128;      // longest per-octet reference list the plan sorts in registers
constexpr int WF_UMAX = 64;            // max union octets of a 16-token tile (tiles beyond it go token by token)

struct WfPlanView {
    int *flags;          // [0] generic d_f path, [1] impure slots (all samples), [2] longest octet list, [3] tokens outside
                         // the tile structure (timp_list), [4] = [0], [5] largest tile union
    int *perm;           // [B*Nq]         token order (per sample) by first neighbour octet
    int *oct_off;        // [B*(NO+1)]
    uint32_t *oct_ent;   // [B*Nq*S]       (i << 5) | s, ascending within an octet
    int *imp_cnt;        // [B]            min(impure slots of the sample, WFP_IMP_CAP)
    uint32_t *imp_list;  // [B*WFP_IMP_CAP]   (i << 5) | s, ascending
    int *frow_cnt;       // [B]            flagged rows of the sample
    int *frow_list;      // [B*WFP_IMP_CAP*8]
    uint8_t *row_flag;   // [B*Nk]
    int *hist_oct;       // [B*NO]         build scratch: histogram, then fill cursor
    int *hist_tok;       // [B*NO]
    int *slot_oct;       // [B*Nq*S]       octet of every slot or -1 (impure)
    int *tok_key;        // [B*Nq]         first neighbour octet of every token
    // tiles: 16 consecutive tokens of perm; the union of their octets and, per (union position, token), the slot or -1
    int *tile_u;         // [B*T]
    int *tile_oct;       // [B*T*WF_UMAX]
    int8_t *slot_t;      // [B*T*WF_UMAX*16]   [tile][u][token row]
    int *timp_list;      // [B*Nq]         global ids (b*Nq + i) of the tokens left to the token-by-token kernels; count = flags[3]
    int NO, S, T;
};

struct WfPlanLayout {
    size_t flags, perm, oct_off, oct_ent, imp_list, frow_list, slot_oct, tok_key, zero_begin, imp_cnt, frow_cnt, row_flag, hist_oct, hist_tok,
        zero_end, tile_u, tile_oct, slot_t, timp_list, total;
    int NO, S, T;
};

inline size_t wfp_align(size_t x) { return (x + 255) & ~(size_t)255; }

inline WfPlanLayout wf_plan_layout(int B, int Nq, int M, int Nk) {
    WfPlanLayout L;
    L.NO = (Nk + 7) / 8;
    L.S = (M % 8 == 0) ? M / 8 : 0;
    size_t o = 0;
    L.flags = o;     o += 256;
    L.perm = o;      o += wfp_align((size_t)B * Nq * 4);
    L.oct_off = o;   o += wfp_align((size_t)B * (L.NO + 1) * 4);
    L.oct_ent = o;   o += wfp_align((size_t)B * Nq * (L.S ? L.S : 1) * 4);
    L.imp_list = o;  o += wfp_align((size_t)B * WFP_IMP_CAP * 4);
    L.frow_list = o; o += wfp_align((size_t)B * WFP_IMP_CAP * 8 * 4);
    L.slot_oct = o;  o += wfp_align((size_t)B * Nq * (L.S ? L.S : 1) * 4);
    L.tok_key = o;   o += wfp_align((size_t)B * Nq * 4);
    L.zero_begin = o;                                   // one memset clears everything from here to zero_end
    L.imp_cnt = o;   o += wfp_align((size_t)B * 4);
    L.frow_cnt = o;  o += wfp_align((size_t)B * 4);
    L.row_flag = o;  o += wfp_align((size_t)B * Nk + 4);
    L.hist_oct = o;  o += wfp_align((size_t)B * L.NO * 4);
    L.hist_tok = o;  o += wfp_align((size_t)B * L.NO * 4);
    L.zero_end = o;
    L.T = (Nq + 15) / 16;
    L.tile_u = o;    o += wfp_align((size_t)B * L.T * 4);
    L.tile_oct = o;  o += wfp_align((size_t)B * L.T * WF_UMAX * 4);
    L.slot_t = o;    o += wfp_align((size_t)B * L.T * WF_UMAX * 16);
    L.timp_list = o; o += wfp_align((size_t)B * Nq * 4);
    L.total = o;
    return L;
}

inline WfPlanView wf_plan_view(void *buf, int B, int Nq, int M, int Nk) {
    const WfPlanLayout L = wf_plan_layout(B, Nq, M, Nk);
    char *p = reinterpret_cast<char *>(buf);
    WfPlanView v;
    v.flags = reinterpret_cast<int *>(p + L.flags);
    v.perm = reinterpret_cast<int *>(p + L.perm);
    v.oct_off = reinterpret_cast<int *>(p + L.oct_off);
    v.oct_ent = reinterpret_cast<uint32_t *>(p + L.oct_ent);
    v.imp_cnt = reinterpret_cast<int *>(p + L.imp_cnt);
    v.imp_list = reinterpret_cast<uint32_t *>(p + L.imp_list);
    v.frow_cnt = reinterpret_cast<int *>(p + L.frow_cnt);
    v.frow_list = reinterpret_cast<int *>(p + L.frow_list);
    v.row_flag = reinterpret_cast<uint8_t *>(p + L.row_flag);
    v.hist_oct = reinterpret_cast<int *>(p + L.hist_oct);
    v.hist_tok = reinterpret_cast<int *>(p + L.hist_tok);
    v.slot_oct = reinterpret_cast<int *>(p + L.slot_oct);
    v.tok_key = reinterpret_cast<int *>(p + L.tok_key);
    v.tile_u = reinterpret_cast<int *>(p + L.tile_u);
    v.tile_oct = reinterpret_cast<int *>(p + L.tile_oct);
    v.slot_t = reinterpret_cast<int8_t *>(p + L.slot_t);
    v.timp_list = reinterpret_cast<int *>(p + L.timp_list);
    v.NO = L.NO;
    v.S = L.S;
    v.T = L.T;
    return v;
}

// entry points of clusten_wf2.cu used by the dispatchers of clusten_wf.cu; each returns 1 when it took the call
int wf3_fwd(const void *w, const void *f, const int64_t *idx, void *out, const void *plan, int B, int Nq, int Nk, int C, int M,
            int IC, int64_t f_sb, int64_t f_sn, int dtype, cudaStream_t st);
// token-by-token forward over plan.timp_list only (the tokens the tile kernel leaves out)
int wf2_fwd_listed(const void *w, const void *f, const int64_t *idx, void *out, const void *plan, int B, int Nq, int Nk, int C, int M,
                   int IC, int64_t f_sb, int64_t f_sn, int dtype, cudaStream_t st);
int wf2_fwd(const void *w, const void *f, const int64_t *idx, void *out, const void *plan, int B, int Nq, int Nk, int C, int M,
            int IC, int64_t f_sb, int64_t f_sn, int dtype, cudaStream_t st);
int wf2_dw(const void *d_out, const void *f, const int64_t *idx, void *d_w, const void *plan, int B, int Nq, int Nk, int C, int M,
           int IC, int64_t f_sb, int64_t f_sn, int dtype, cudaStream_t st);
int wf2_df(const void *d_out, const void *w, const int64_t *idx, void *d_f, const void *plan, int B, int Nq, int Nk, int C, int M,
           int IC, int64_t df_sb, int64_t df_sn, int dtype, cudaStream_t st);

}  // namespace clusten
