"""CPU: size-independent properties of the oracle's integer path (what the GPU parity tests lean on at sizes the oracle cannot
reach): the clustering is a permutation into balanced clusters, kNN is sorted with the canonical tie rule, the merge
selection is a duplicate-free set holding every reserve token, neighbourhood indices stay in range -- on random and on
degenerate inputs (ties everywhere, n not a multiple of m, a handful of clusters).  (A single cluster is not a valid input
of the reference's clustering -- point_utils.py:220-225 needs three anchors -- and never reaches it: nbhd_size >= n takes the
global-attention branch, aff.py:442.)"""
import math

import pytest
import torch

from oracle import inputs
from oracle import point_ops as pt


@pytest.mark.parametrize("n,h,w,m", [(1024, 32, 32, 8), (1000, 32, 32, 8), (131, 16, 16, 8), (777, 64, 16, 24), (24, 8, 8, 8), (49, 8, 8, 8)])
def test_space_filling_cluster_is_a_balanced_permutation(n, h, w, m):
    B = 3
    pos = inputs.grid_positions(B, h, w) if n == h * w else inputs.random_positions(B, n, h, w, seed=n)
    spos, mean_pos, member, mask, ranking = pt.space_filling_cluster(pos, m, h, w)
    k = math.ceil(n / m)
    assert spos.shape == (B, n, 2) and mean_pos.shape == (B, k, 2) and member.shape == (B, k, m) and ranking.shape == (B, n, 1)
    for b in range(B):
        assert torch.equal(torch.sort(ranking[b, :, 0])[0], torch.arange(n))                        # a permutation of the tokens
    assert torch.equal(spos, pos.gather(1, ranking.expand(-1, -1, 2)))                                # pos reordered by it
    flat = member.reshape(B, -1)
    assert torch.equal(flat[:, :n], torch.arange(n).expand(B, -1))                                    # clusters = consecutive runs
    if k * m == n:
        assert mask is None
    else:
        assert mask.shape == (B, k, m) and int(mask.sum()) == B * n
        assert not flat[:, n:].any() and not mask.reshape(B, -1)[:, n:].any()                         # padded tail -> row 0, masked
        cnt = mask.sum(2, keepdim=True).clamp_min(1)
        want = (torch.cat([spos, spos.new_zeros(B, k * m - n, 2)], 1).reshape(B, k, m, 2) * mask.unsqueeze(3)).sum(2) / cnt
        assert torch.allclose(mean_pos, want, atol=1e-5)


@pytest.mark.parametrize("nq,ndb,k", [(300, 200, 6), (64, 64, 2), (50, 9, 9), (40, 500, 4)])
def test_knn_is_sorted_with_lowest_index_ties(nq, ndb, k):
    g = torch.Generator().manual_seed(nq + ndb)
    q = torch.randint(0, 12, (2, nq, 2), generator=g).float()                # small integer grid: many exact distance ties
    db = torch.randint(0, 12, (2, ndb, 2), generator=g).float()
    idx, dist = pt.knn(q, db, k, return_dist=True)
    assert idx.shape == (2, nq, k) and idx.dtype == torch.int64 and idx.is_contiguous()
    assert (dist[..., 1:] >= dist[..., :-1]).all()                                                    # ascending
    d_all = pt.sqrt_rn(((q[:, :, None] - db[:, None]) ** 2).sum(-1))
    assert torch.equal(dist, d_all.gather(2, idx))                                                    # the reported distances are the true ones
    same = dist[..., 1:] == dist[..., :-1]
    assert (idx[..., 1:][same] > idx[..., :-1][same]).all()                                           # ties -> lower database index first
    kth = dist[..., -1:]
    chosen = torch.zeros_like(d_all, dtype=torch.bool).scatter_(2, idx, True)
    assert (d_all[~chosen].reshape(2, nq, -1) >= kth).all()                                           # nothing closer was left out


@pytest.mark.parametrize("hw,stride,ds_rate", [(32, 2, 0.25), (32, 2, 0.2), (24, 2, 0.2)])
def test_merge_selection_is_a_set_with_all_reserve_tokens(hw, stride, ds_rate):
    B, n = 2, hw * hw
    pos = inputs.grid_positions(B, hw, hw)
    g = torch.Generator().manual_seed(hw)
    prob = torch.rand(B, n, 1, generator=g).round(decimals=1)               # coarse scores: ties between candidates
    reserve_num = math.ceil(hw / (stride * 2)) ** 2
    idx = pt.merge_select(pos, prob, stride, 4.0, ds_rate, reserve_num)
    keep = int(n * ds_rate)
    assert idx.shape == (B, keep, 1)
    for b in range(B):
        chosen = idx[b, :, 0]
        assert chosen.unique().numel() == keep and int(chosen.min()) >= 0 and int(chosen.max()) < n   # no duplicates, in range
        reserve = ((pos[b].long() % (2 * stride)) == 0).all(-1).nonzero()[:, 0]
        assert torch.equal(chosen[keep - reserve_num:], reserve)                                      # reserve tokens last, ascending
        score, _ = pt.merge_scores(pos, prob, stride, 4.0)
        picked, rest = score[b, chosen[:keep - reserve_num]], score[b].clone()
        rest[chosen] = -1e9
        assert float(picked.min()) >= float(rest.max())                                               # a true top-k of the scores
        assert (picked[1:] <= picked[:-1]).all()                                                      # canonical order: descending score


@pytest.mark.parametrize("n,m,nbhd", [(1000, 8, 48), (131, 8, 48), (600, 24, 144)])
def test_neighbourhood_indices_and_table_rows_stay_in_range(n, m, nbhd):
    B, hw = 2, 64
    pos, nb, mask, pe_idx = inputs.structured_neighbourhood(B, n, hw, hw, m, nbhd, seed=n)
    k = math.ceil(n / m)
    M = m * min(int(round(nbhd / float(m))), k)
    assert nb.shape == (B, n, M) and int(nb.min()) >= 0 and int(nb.max()) < n
    assert int(pe_idx.min()) >= 0 and int(pe_idx.max()) < 1023 * 1023
    if mask is not None:
        assert mask.shape == nb.shape and set(mask.unique().tolist()) <= {0, 1}
        assert not nb[mask == 0].any()                                       # padded slots point at row 0 (point_utils.py:283)
    # every token's own cluster is among its nearest clusters: it attends to itself
    own = torch.arange(n).view(1, n, 1).expand(B, -1, -1)
    assert ((nb == own) & ((mask if mask is not None else torch.ones_like(nb)) == 1)).any(-1).all()
