"""Point utilities of the CLUSTEN path on libclusten_b200.so: host-side mirror of the functions of the reference's
``mask2former/modeling/backbone/point_utils.py`` that the AFF backbone and the point-cloud pixel decoder call
(same names, argument meaning and return conventions):

    knn_keops(query, database, k, return_dist=False)                     point_utils.py:28-60   -> clusten_knn
    space_filling_cluster(pos, m, h, w, no_reorder=False, sf_type='', use_anchor=True)
                                                                         point_utils.py:135-287 -> clusten_sfc_cluster
    shepard_decay_weights(dist, power=3)                                 point_utils.py:63-75   (elementwise torch)
    upsample_feature_shepard(query, database, feature, ...)              point_utils.py:78-121  -> clusten_knn + clusten_wg_*
    topk_select / mask_select / merge_select                             aff.py:292-329         -> clusten_topk_select, clusten_mask_select

Tie rules are canonical (stable; ties -> lowest index), see DESIGN.md.  CUDA tensors only: no CPU path.
"""
import math

import torch

from . import _lib
from .ops import WEIGHTEDGATHERFunction, _call


def _f32c(t):
    t = t.detach()
    if t.dtype != torch.float32:
        t = t.to(torch.float32)                       # point_utils.py:45-49 (positions are always searched in fp32)
    return t.contiguous()


def knn_keops(query, database, k, return_dist=False):
    """k nearest database points of each query: idx int64 [B, n_q, k] ascending by distance (ties -> lowest index),
    plus the fp32 distances when ``return_dist`` (returned as ``(nn_idx, nn_dist)``, point_utils.py:55-57)."""
    dev = _lib.require_cuda(query, database)
    q, d = _f32c(query), _f32c(database)
    if q.dim() != 3 or d.dim() != 3 or q.shape[2] != 2 or d.shape[2] != 2 or q.shape[0] != d.shape[0]:
        raise RuntimeError(f"knn: expected [B,n,2] positions, got {tuple(q.shape)} / {tuple(d.shape)}")
    B, Nq, _ = q.shape
    Ndb = d.shape[1]
    idx = torch.empty((B, Nq, k), dtype=torch.int64, device=dev)
    dist = torch.empty((B, Nq, k), dtype=torch.float32, device=dev) if return_dist else None
    with torch.cuda.device(dev):
        _call("clusten_knn", dev, q.data_ptr(), d.data_ptr(), B, Nq, Ndb, k, idx.data_ptr(), _lib.ptr(dist))
    if return_dist:
        return idx, dist
    return idx


knn = knn_keops


def space_filling_cluster(pos, m, h, w, no_reorder=False, sf_type='', use_anchor=True):
    """Balanced clustering along the boustrophedon anchor curve (default branch of point_utils.py:135-287).
    Returns (pos [B,n,2] reordered, cluster_mean_pos [B,k,2], member_idx [B,k,m] int64,
    cluster_mask [B,k,m] int64 or None when k*m == n, pos_ranking [B,n,1] int64)."""
    if no_reorder or sf_type != '' or not use_anchor:
        raise NotImplementedError("only the default branch (no_reorder=False, sf_type='', use_anchor=True) is on the "
                                  "accelerated path; Peano/Hilbert orders are out of scope (DESIGN.md)")
    dev = _lib.require_cuda(pos)
    dtype_in = pos.dtype
    p = _f32c(pos)
    B, n, d = p.shape
    if d != 2:
        raise RuntimeError("space_filling_cluster: positions must be 2-D")
    k = int(math.ceil(n / float(m)))
    L = _lib.lib()
    pos_sorted = torch.empty_like(p)
    mean_pos = torch.empty((B, k, 2), dtype=torch.float32, device=dev)
    member_idx = torch.empty((B, k, m), dtype=torch.int64, device=dev)
    cluster_mask = torch.empty((B, k, m), dtype=torch.int64, device=dev) if k * m != n else None
    ranking = torch.empty((B, n, 1), dtype=torch.int64, device=dev)
    ws_bytes = L.clusten_sfc_workspace_bytes(B, n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _call("clusten_sfc_cluster", dev, p.data_ptr(), B, n, m, h, w, pos_sorted.data_ptr(), mean_pos.data_ptr(),
              member_idx.data_ptr(), _lib.ptr(cluster_mask), ranking.data_ptr(), ws.data_ptr(), ws_bytes)
    if dtype_in != torch.float32:
        pos_sorted = pos_sorted.to(dtype_in)
    return pos_sorted, mean_pos, member_idx, cluster_mask, ranking


PE_TABLE_WIDTH = 1023                                  # aff.py:17-19 (2 * (2048 // 4 - 1) + 1)


def stage_prepare(pos, nearest, member, cluster_mask, want_mask64=True, extent=None):
    """Neighbourhood assembly of one AFF stage (aff.py:475-485) in one pass: returns
    (member_idx int64 [B,n,M], mask int64 [B,n,M] or None, mask_u8 uint8 [B,n,M] or None, uniq int64 [U], bias_idx int32
    [B,n,M]) with uniq = the ascending distinct relative-position table rows ``pe_idx`` takes and bias_idx its inverse map
    (``torch.unique(pe_idx, return_inverse=True)`` without the sort).  One host read (U) -- unless ``extent = (h, w)`` of
    the stem grid the positions live on is given: then uniq is returned at its upper bound (2h-1)(2w-1) rows (the tail
    repeats table row 0) together with the device-side count as a sixth result, and nothing is read back."""
    dev = _lib.require_cuda(pos, nearest, member, cluster_mask)
    p = _f32c(pos)
    nearest, member = nearest.contiguous(), member.contiguous()
    B, n, nnc = nearest.shape
    k, m = member.shape[1], member.shape[2]
    M = nnc * m
    cm = None if cluster_mask is None else cluster_mask.contiguous()
    member_idx = torch.empty((B, n, M), dtype=torch.int64, device=dev)
    mask64 = torch.empty((B, n, M), dtype=torch.int64, device=dev) if (cm is not None and want_mask64) else None
    mask8 = torch.empty((B, n, M), dtype=torch.uint8, device=dev) if cm is not None else None
    pe = torch.empty((B, n, M), dtype=torch.int32, device=dev)
    bias_idx = torch.empty((B, n, M), dtype=torch.int32, device=dev)
    cap = max(1, min(PE_TABLE_WIDTH * PE_TABLE_WIDTH, B * n * M))
    if extent is not None:                         # |dx| <= w-1, |dy| <= h-1 (clamped to the table): at most that many distinct rows
        cap = min(cap, min(2 * int(extent[1]) - 1, PE_TABLE_WIDTH) * min(2 * int(extent[0]) - 1, PE_TABLE_WIDTH))
    uniq = (torch.zeros if extent is not None else torch.empty)(cap, dtype=torch.int32, device=dev)
    count = torch.empty(1, dtype=torch.int32, device=dev)
    L = _lib.lib()
    ws_bytes = L.clusten_prepare_workspace_bytes()
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _call("clusten_stage_prepare", dev, nearest.data_ptr(), member.data_ptr(), _lib.ptr(cm), p.data_ptr(), B, n, k, m, nnc,
              member_idx.data_ptr(), _lib.ptr(mask64), _lib.ptr(mask8), pe.data_ptr(), bias_idx.data_ptr(), uniq.data_ptr(), cap,
              count.data_ptr(), ws.data_ptr(), ws_bytes)
    if extent is not None:
        return member_idx, mask64, mask8, uniq.long(), bias_idx, count
    U = int(count.item())
    return member_idx, mask64, mask8, uniq[:U].long(), bias_idx


def table_rank(pe_idx):
    """``torch.unique(pe_idx, return_inverse=True)`` over relative-position table rows without the sort and without a host read
    (clusten_table_rank): returns (uniq int64 [cap] -- the ascending distinct rows in its first ``count`` entries, the tail repeats
    row 0 --, inverse int32 shaped like pe_idx, count int32 device scalar).  cap = min(1023^2, pe_idx.numel())."""
    dev = _lib.require_cuda(pe_idx)
    pe = pe_idx.contiguous()
    if pe.dtype != torch.int64:
        pe = pe.long()
    total = pe.numel()
    cap = max(1, min(PE_TABLE_WIDTH * PE_TABLE_WIDTH, total))
    inverse = torch.empty(pe.shape, dtype=torch.int32, device=dev)
    uniq = torch.zeros(cap, dtype=torch.int32, device=dev)
    count = torch.empty(1, dtype=torch.int32, device=dev)
    L = _lib.lib()
    ws_bytes = L.clusten_prepare_workspace_bytes()
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _call("clusten_table_rank", dev, pe.data_ptr(), total, inverse.data_ptr(), uniq.data_ptr(), cap, count.data_ptr(), ws.data_ptr(), ws_bytes)
    return uniq.long(), inverse, count


def topk_select(score, k, out=None):
    """Canonical ``score.topk(k, sorted=False)[1]`` (aff.py:320): first k of a stable descending sort, int64 [B,k]."""
    dev = _lib.require_cuda(score)
    s = _f32c(score)
    B, n = s.shape
    if out is None:
        out = torch.empty((B, k), dtype=torch.int64, device=dev)
    L = _lib.lib()
    ws_bytes = L.clusten_topk_workspace_bytes(B, n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _call("clusten_topk_select", dev, s.data_ptr(), B, n, k, out.data_ptr(), out.stride(0), ws.data_ptr(), ws_bytes)
    return out


def mask_select(mask, count, out=None):
    """``mask.nonzero(as_tuple=True)[1].reshape(B, count)`` (aff.py:323) without the dynamic shape / host sync."""
    dev = _lib.require_cuda(mask)
    mk = _f32c(mask)
    B, n = mk.shape
    if out is None:
        out = torch.empty((B, count), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        _call("clusten_mask_select", dev, mk.data_ptr(), B, n, count, out.data_ptr(), out.stride(0))
    return out


def merge_scores(pos, min_dist, learned_prob, stride, alpha, reserve_on=True):
    """(final_prob [B,n], reserve_mask [B,n] or None) of ClusterMerging.forward (aff.py:292-315) in one kernel (clusten_merge_scores)
    instead of ~22 element-wise launches: pos fp32 [B,n,2], min_dist fp32 [B,n,>=2] (the kNN-2 distances of aff.py:299) or None for
    stride 2, learned_prob [B,n,1] / [B,n] or None.  Bit-identical to the op-by-op formulation (each fp32 operation rounded alone)."""
    dev = _lib.require_cuda(pos, min_dist, learned_prob)
    B, n = pos.shape[0], pos.shape[1]
    p = _f32c(pos)
    md = None if min_dist is None else _f32c(min_dist)
    lp = None if learned_prob is None else _f32c(learned_prob.detach().reshape(B, n))
    final_prob = torch.empty((B, n), dtype=torch.float32, device=dev)
    reserve = torch.empty((B, n), dtype=torch.float32, device=dev) if reserve_on else None
    if B * n:
        with torch.cuda.device(dev):
            _call("clusten_merge_scores", dev, p.data_ptr(), _lib.ptr(md), 0 if md is None else md.shape[-1], _lib.ptr(lp), float(alpha),
                  int(stride), int(bool(reserve_on)), final_prob.data_ptr(), _lib.ptr(reserve), B, n)
    return final_prob, reserve


def merge_select(final_prob, reserve_mask, keep_num, reserve_num):
    """idx [B, keep_num, 1] = cat(canonical top-(keep-reserve) of final_prob, reserve tokens ascending) (aff.py:320-324)."""
    B = final_prob.shape[0]
    if keep_num < reserve_num:                                           # the reference fails in topk(negative k) (aff.py:316-320)
        raise RuntimeError(f"merge_select: keep_num {keep_num} < reserve_num {reserve_num} (ds_rate too small for this grid)")
    idx = torch.empty((B, keep_num), dtype=torch.int64, device=final_prob.device)
    k = keep_num - reserve_num
    if k > 0:
        topk_select(final_prob, k, out=idx)
    if reserve_num > 0:
        mask_select(reserve_mask, reserve_num, out=idx[:, k:])
    return idx.unsqueeze(2)


def shepard_decay_weights(dist, power=3):
    """Inverse-distance weights (point_utils.py:63-75)."""
    dist = dist.clamp(min=1e-2)
    ipd = 1.0 / (dist.pow(power) + 1e-6)
    return ipd / (ipd.sum(dim=2, keepdim=True) + 1e-6)


def upsample_feature_shepard(query, database, feature, database_idx=None, k=4, power=3, custom_kernel=True,
                             nn_idx=None, return_weight_only=False):
    """kNN inverse-distance interpolation of ``feature`` (at ``database``) to ``query`` (point_utils.py:78-121).
    ``custom_kernel`` is accepted for signature parity; the CUDA weighted gather is always used."""
    b, n_, d = database.shape
    n = query.shape[1]
    if (n == n_) and bool((query == database).all()):                    # point_utils.py:97
        return feature
    if nn_idx is not None:
        k = nn_idx.shape[-1]
    else:
        k = min(k, n_)
        nn_idx = knn_keops(query, database, k=k, return_dist=False)
    nn_pos = database.gather(index=nn_idx.view(b, -1, 1).expand(-1, -1, 2), dim=1).reshape(b, n, k, d)
    nn_dist = (query.unsqueeze(2) - nn_pos).pow(2).sum(-1)               # squared distance, point_utils.py:105
    nn_weights = shepard_decay_weights(nn_dist, power=power)
    if return_weight_only:
        return nn_weights
    c = feature.shape[-1]
    assert feature.shape[1] == n_
    up_features = WEIGHTEDGATHERFunction.apply(nn_idx, nn_weights, feature)
    if database_idx is not None:
        up_features.scatter_(dim=1, index=database_idx.long().expand(-1, -1, c), src=feature)
    return up_features
