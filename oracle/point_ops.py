"""CPU restatement of the integer / geometry side of the CLUSTEN path (TEST INFRASTRUCTURE).

Follows mask2former/modeling/backbone/point_utils.py and aff.py of the reference; every function
cites the lines it restates.  Where the reference leaves tie order to the library (torch.sort,
torch.topk(sorted=False), KeOps argKmin) the oracle fixes the canonical rule documented in
oracle/__init__.py: stable sorts, ties -> lowest index.
"""
import math

import numpy as np
import torch


def sqrt_rn(x):
    """Correctly rounded fp32 square root.  torch.sqrt on CPU is NOT correctly rounded on every build (measured in the
    authoring container: 6 of 640 multiples of 1/64 near 144 are off by one ulp), numpy's is (it matches the
    round-to-nearest result obtained through float64).  The CUDA path uses __fsqrt_rn."""
    return torch.from_numpy(np.sqrt(x.contiguous().numpy()))



# --------------------------------------------------------------------------------------- kNN
def knn(query, database, k, return_dist=False, chunk=2048):
    """Brute-force kNN, restating knn_keops (point_utils.py:28-60).

    dist = sqrt(sum_d (q - db)^2) in IEEE fp32 with every op rounded separately (no FMA); the k
    smallest per query in ascending distance, ties -> lowest database index.  Returns int64
    [B, Nq, k] contiguous (+ fp32 distances [B, Nq, k] when ``return_dist``) -- the reference's
    callers ``.view`` the result (aff.py:478).  pykeops is not vendored: PARITY UNPINNED, see
    oracle/__init__.py.
    """
    q = query.detach().to(torch.float32)
    d = database.detach().to(torch.float32)
    B, Nq, _ = q.shape
    idx_out = torch.empty(B, Nq, k, dtype=torch.int64)
    dist_out = torch.empty(B, Nq, k, dtype=torch.float32)
    ar = torch.arange(d.shape[1], dtype=torch.int64)
    for b in range(B):
        for s in range(0, Nq, chunk):
            diff = q[b, s:s + chunk, None, :] - d[b, None, :, :]          # nq x ndb x D
            sq = diff * diff
            acc = sq[..., 0]
            for t in range(1, sq.shape[-1]):
                acc = acc + sq[..., t]                                    # left-to-right, rounded adds
            dist = sqrt_rn(acc)
            # canonical order = ascending (distance, index): dist >= 0 so its bit pattern is order-preserving and
            # (bits << 32 | index) is a total order; the k smallest keys ARE the first k of a stable ascending sort.
            key = (dist.view(torch.int32).to(torch.int64) << 32) | ar
            top = torch.topk(key, k, dim=1, largest=False, sorted=True)[0]
            idx_out[b, s:s + chunk] = top & 0xFFFFFFFF
            dist_out[b, s:s + chunk] = (top >> 32).to(torch.int32).view(torch.float32)
    if return_dist:
        return idx_out, dist_out
    return idx_out


def knn_tie_report(query, database, idx):
    """For a kNN result, count queries whose k-th / (k+1)-th distances tie (set ambiguous in the
    reference) -- used by tests to report how much of the answer is canonical-rule dependent."""
    i2, d2 = knn(query, database, idx.shape[-1] + 1, return_dist=True)
    return int((d2[..., -1] == d2[..., -2]).sum())


# ------------------------------------------------------------------- space-filling clustering
def anchor_grid(n, m, h, w):
    """Host-side scalars of point_utils.py:167-176 (Python double arithmetic, banker's round)."""
    k = int(math.ceil(n / m))
    patch_len = (h * w / k) ** 0.5
    num_patch_h = int(round(h / patch_len))
    num_patch_w = int(round(w / patch_len))
    return k, num_patch_h, num_patch_w, h / num_patch_h, w / num_patch_w


def space_filling_cluster(pos, m, h, w):
    """Balanced clustering along a boustrophedon curve of anchors: the default branch
    (sf_type='', use_anchor=True, no_reorder=False) of point_utils.py:135-287.

    Returns (pos_sorted [B,n,2] f32, cluster_mean_pos [B,k,2] f32, member_idx [B,k,m] i64,
    cluster_mask [B,k,m] i64 or None, pos_ranking [B,n,1] i64).
    The reference's ``sort`` (point_utils.py:238) is unstable; the oracle uses stable=True.
    """
    pos = pos.detach().to(torch.float32)
    B, n, d = pos.shape
    k, nph, npw, plh, plw = anchor_grid(n, m, h, w)

    ys, xs = torch.meshgrid(torch.arange(nph), torch.arange(npw), indexing="ij")
    grid_pos = torch.stack([xs, ys], dim=2).reshape(-1, 2)                      # :187-191
    score = torch.where(ys % 2 == 0, xs, -xs) + ys * w                          # :203-206
    score = score + torch.where(ys % 2 == 1, w - 1, 0)                          # :207
    order_idx = score.reshape(-1).sort()[1]                                     # rank -> anchor  :209
    order_grid_idx = torch.empty_like(order_idx)
    order_grid_idx[order_idx] = torch.arange(order_idx.numel())                 # anchor -> rank  :210-212

    ordered_grid = grid_pos[order_idx]
    plen = torch.tensor([plw, plh], dtype=torch.float32)                        # :215
    init = ordered_grid * plen + plen / 2 - 0.5                                 # :217
    nump = init.shape[0]
    prev = torch.zeros_like(init)
    prev[1:] = init[:nump - 1]
    prev[0] = prev[1] - (prev[2] - prev[1])                                     # :222
    nxt = torch.zeros_like(init)
    nxt[:nump - 1] = init[1:]
    nxt[-1] = nxt[-2] + (nxt[-2] - nxt[-3])                                     # :225

    cell = (pos / plen).floor()                                                 # :227
    anchor = (cell[..., 0] + cell[..., 1] * npw).long()                         # :228
    rank = order_grid_idx[anchor]                                               # B x n      :229
    dp = pos - prev[rank]
    dn = pos - nxt[rank]
    dist_prev = dp[..., 0] * dp[..., 0] + dp[..., 1] * dp[..., 1]               # :233
    dist_next = dn[..., 0] * dn[..., 0] + dn[..., 1] * dn[..., 1]               # :234
    ratio = dist_prev / (dist_next + 1e-5)                                      # :235
    key = rank * (ratio.max() + 1) + ratio                                      # :237 (fp32, batch-global max)
    pos_ranking = key.sort(dim=1, stable=True)[1].unsqueeze(2)                  # :238

    pos = pos.gather(1, pos_ranking.expand(-1, -1, d))                          # :260
    if k * m == n:
        cluster_mask = None
        cluster_mean_pos = pos.reshape(B, k, m, d).mean(2)                      # :262-264
    else:
        pad = torch.zeros(B, k * m, d)
        pad[:, :n] = pos
        cluster_mask = torch.zeros(B, k * m, dtype=torch.int64)
        cluster_mask[:, :n] = 1
        cluster_mask = cluster_mask.reshape(B, k, m)
        cluster_mean_pos = pad.reshape(B, k, m, d).sum(2) / cluster_mask.sum(2, keepdim=True)  # :266-271
    member_idx = torch.arange(k * m)
    member_idx[n:] = 0                                                          # :282-283
    member_idx = member_idx.unsqueeze(0).expand(B, -1).reshape(B, k, m)
    return pos, cluster_mean_pos, member_idx, cluster_mask, pos_ranking


REL_POS_WIDTH = 2048 // 4 - 1          # aff.py:18
TABLE_WIDTH = 2 * REL_POS_WIDTH + 1    # aff.py:19


def build_pre_table():
    """The 1023^2 x 5 relative-position feature table (dx, dy, dist, dy/dist, dx/dist), NaN/inf
    centre zeroed (aff.py:21-31)."""
    r = torch.arange(TABLE_WIDTH).float() - REL_POS_WIDTH
    ys, xs = torch.meshgrid(r, r, indexing="ij")
    dis = (ys ** 2 + xs ** 2) ** 0.5
    t = torch.stack([xs, ys, dis, ys / dis, xs / dis], dim=2)
    t[torch.bitwise_or(t.isnan(), t.isinf())] = 0
    return t.reshape(-1, 5)


def assemble_neighbourhood(pos, cluster_mean_pos, member_idx, cluster_mask, nnc):
    """Neighbourhood assembly of BasicLayer.forward (aff.py:475-485): nnc nearest clusters ->
    member_idx [B,n,M], mask [B,n,M] or None, pe_idx [B,n,M] (index into the 1023^2 rel-pos table)."""
    B, n, d = pos.shape
    m = member_idx.shape[2]
    M = m * nnc
    nearest = knn(pos, cluster_mean_pos, nnc)                                                  # :475
    gi = nearest.reshape(B, -1, 1).expand(-1, -1, m)
    nb = member_idx.gather(1, gi).reshape(B, n, M)                                             # :478
    mask = None
    if cluster_mask is not None:
        mask = cluster_mask.gather(1, gi).reshape(B, n, M)                                     # :480
    pos_nb = pos.gather(1, nb.reshape(B, -1, 1).expand(-1, -1, d)).reshape(B, n, M, d)
    rel = pos_nb - (pos.unsqueeze(2) - REL_POS_WIDTH)                                          # :481-482
    rel = rel.clamp(0, TABLE_WIDTH - 1)                                                        # :484
    pe_idx = (rel[..., 1] * TABLE_WIDTH + rel[..., 0]).long()                                  # :485
    return nb, mask, pe_idx


# ------------------------------------------------------------------ adaptive-downsampling top-k
def topk_canonical(score, k):
    """First k of a stable descending sort: the canonical order for topk(sorted=False) (aff.py:320)."""
    return torch.sort(score, dim=1, descending=True, stable=True)[1][:, :k]


def merge_scores(pos, learned_prob, stride, alpha, min_dist=None):
    """final_prob and reserve_mask of ClusterMerging.forward (aff.py:292-315).
    ``min_dist``: distance to the nearest OTHER token (index 1 of self-kNN-2, aff.py:299-300);
    computed here when None and stride != 2."""
    B, n, _ = pos.shape
    pos_long = pos.long()
    if stride == 2:
        grid_prob = ((pos_long % stride) == 0).all(-1).float()                                 # :297
    else:
        if min_dist is None:
            min_dist = knn(pos, pos, 2, return_dist=True)[1][:, :, 1]                          # :299-300
        ada_stride = 2 ** (min_dist.log2().ceil() + 1)                                         # :301
        grid_prob = ((pos_long % ada_stride.unsqueeze(2).long()) == 0).all(-1).float()         # :302
    final_prob = grid_prob
    if learned_prob is not None:
        final_prob = final_prob + learned_prob.detach().view(B, n) * alpha                     # :307-310
    reserve_mask = ((pos_long % (stride * 2)) == 0).all(dim=-1).float()                        # :313
    final_prob = final_prob + reserve_mask * (-100)                                            # :314
    return final_prob, reserve_mask


def merge_select(pos, learned_prob, stride, alpha, ds_rate, reserve_num, min_dist=None):
    """Token selection of ClusterMerging.forward (aff.py:292-329), reserve_on=True.
    Returns idx [B, keep, 1] int64: canonical top-k picks followed by the reserve tokens ascending."""
    B, n, _ = pos.shape
    keep_num = int(n * ds_rate)                                                                # :292
    final_prob, reserve_mask = merge_scores(pos, learned_prob, stride, alpha, min_dist)
    sample_idx = topk_canonical(final_prob, keep_num - reserve_num)                            # :320
    reserve_idx = reserve_mask.nonzero(as_tuple=True)[1].reshape(B, reserve_num)               # :323
    return torch.cat([sample_idx, reserve_idx], dim=-1).unsqueeze(2)                           # :324


# ------------------------------------------------------------------------ Shepard upsampling
def shepard_decay_weights(dist, power=3):
    """Inverse-distance weights (point_utils.py:63-75); ``dist`` is the SQUARED distance at the
    upsample call site (point_utils.py:105)."""
    dist = dist.clamp(min=1e-2)
    ipd = 1.0 / (dist.pow(power) + 1e-6)
    return ipd / (ipd.sum(dim=2, keepdim=True) + 1e-6)


def upsample_feature_shepard(query, database, feature, k=4, power=3):
    """kNN-k inverse-distance interpolation (point_utils.py:78-121, default arguments)."""
    from .clusten_ops import wg_forward
    B, n_, d = database.shape
    n = query.shape[1]
    if n == n_ and bool((query == database).all()):                                            # :97
        return feature
    k = min(k, n_)
    nn_idx = knn(query, database, k)                                                           # :103
    nn_pos = database.gather(1, nn_idx.view(B, -1, 1).expand(-1, -1, d)).reshape(B, n, k, d)
    nn_dist = (query.unsqueeze(2) - nn_pos).pow(2).sum(-1)                                     # :105
    w = shepard_decay_weights(nn_dist, power=power)
    return wg_forward(nn_idx, w, feature)                                                      # :114
