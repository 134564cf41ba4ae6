"""GPU parity of the integer path -- kNN, balanced space-filling-curve clustering, adaptive-downsampling selection,
Shepard upsampling -- against the golden vectors (outputs of the reference's own Python) and the CPU oracle.
Everything here is BIT-EXACT: indices equal, fp32 distances / cluster means compared as int32 bit patterns."""
import glob
import math
import os

import numpy as np
import pytest
import torch

from oracle import inputs
from oracle import point_ops as pt

from conftest import GOLDEN, rel_err

pytestmark = pytest.mark.gpu


def _bits(t):
    return t.detach().cpu().contiguous().view(torch.int32)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "sfc_*.npz"))), ids=os.path.basename)
def test_space_filling_cluster_golden(path):
    import autofocusformermod_b200 as P
    g = np.load(path)
    pos_in = torch.from_numpy(g["pos_in"].astype(np.float32)).cuda()
    p, mean, member, mask, rank = P.space_filling_cluster(pos_in, int(g["m"]), int(g["h"]), int(g["w"]))
    assert torch.equal(rank.cpu(), torch.from_numpy(g["rank"].astype(np.int64)))
    assert torch.equal(_bits(p), _bits(torch.from_numpy(g["pos"])))
    assert torch.equal(_bits(mean), _bits(torch.from_numpy(g["mean"])))
    assert torch.equal(member.cpu(), torch.from_numpy(g["member"].astype(np.int64)))
    if g["mask"].size:
        assert torch.equal(mask.cpu(), torch.from_numpy(g["mask"].astype(np.int64)))
    else:
        assert mask is None


SFC_CASES = [  # B, n, h, w, m      the exact token counts of the BASELINE configs
    (2, 16384, 128, 128, 8), (2, 4096, 128, 128, 8), (2, 3276, 128, 128, 8), (3, 1024, 128, 128, 8),
    (2, 655, 128, 128, 8), (2, 256, 128, 128, 8), (2, 131, 128, 128, 8),
    (1, 32768, 128, 256, 24), (2, 8192, 128, 256, 24), (2, 2048, 128, 256, 24), (2, 512, 128, 256, 24),
    (1, 131072, 256, 512, 8), (1, 32768, 256, 512, 8), (1, 49, 28, 28, 8),
]


@pytest.mark.parametrize("B,n,h,w,m", SFC_CASES)
def test_space_filling_cluster_vs_oracle(B, n, h, w, m):
    import autofocusformermod_b200 as P
    pos = inputs.grid_positions(B, h, w) if n == h * w else inputs.random_positions(B, n, h, w, seed=n)
    ref = pt.space_filling_cluster(pos, m, h, w)
    got = P.space_filling_cluster(pos.cuda(), m, h, w)
    names = ["pos", "cluster_mean_pos", "member_idx", "cluster_mask", "pos_ranking"]
    for name, a, b in zip(names, got, ref):
        if b is None:
            assert a is None, name
        elif b.is_floating_point():
            assert torch.equal(_bits(a), _bits(b)), name
        else:
            assert torch.equal(a.cpu(), b), name
    # size-independent property: pos_ranking is a permutation and pos == pos_in[pos_ranking]
    r = got[4].squeeze(2)
    assert torch.equal(r.sort(1)[0], torch.arange(n, device="cuda").expand(B, -1))
    assert torch.equal(got[0], pos.cuda().gather(1, got[4].expand(-1, -1, 2)))


KNN_CASES = [  # B, nq, ndb, k, kind
    (2, 16384, 2048, 6, "cluster"),      # stage-0 tokens -> cluster means (aff.py:475)
    (2, 4096, 512, 6, "cluster"),
    (2, 655, 82, 6, "cluster"),
    (2, 4096, 4096, 2, "self"),          # ClusterMerging nearest other token, with distances (aff.py:299)
    (1, 16384, 16384, 9, "self-grid"),   # PointConv self-kNN-9 on the res2 grid: 100 % distance ties (msdeformattn_pc.py:295)
    (2, 16384, 4096, 4, "up"),           # Shepard upsample grid <- tokens: 42 % set ties (point_utils.py:103)
    (1, 300, 5, 4, "up"), (1, 7, 16, 16, "up"), (1, 1, 1, 1, "up"),
]


@pytest.mark.parametrize("B,nq,ndb,k,kind", KNN_CASES)
def test_knn_bit_exact(B, nq, ndb, k, kind):
    import autofocusformermod_b200 as P
    if kind == "cluster":
        pos = inputs.grid_positions(B, 128, 128) if nq == 16384 else inputs.random_positions(B, nq, 128, 128, seed=nq)
        q, db = pt.space_filling_cluster(pos, 8, 128, 128)[:2]
    elif kind == "self":
        q = db = inputs.random_positions(B, nq, 128, 128, seed=1)
    elif kind == "self-grid":
        q = db = inputs.grid_positions(B, 128, 128)
    else:
        side = max(int(math.isqrt(nq)), 1)
        q = inputs.grid_positions(B, side, side)[:, :nq] if side * side >= nq else inputs.random_positions(B, nq, 64, 64, seed=2)
        db = inputs.random_positions(B, ndb, max(side, 4), max(side, 4), seed=3)
    ri, rd = pt.knn(q, db, k, return_dist=True)
    gi, gd = P.knn_keops(q.cuda(), db.cuda(), k, return_dist=True)
    assert gi.is_contiguous() and gi.dtype == torch.int64
    assert torch.equal(_bits(gd), _bits(rd)), "distances must be bit-identical"
    assert torch.equal(gi.cpu(), ri), "indices must follow the canonical tie rule"
    assert torch.equal(P.knn_keops(q.cuda(), db.cuda(), k).cpu(), ri)


def _stage_tokens(B, n, h, w, stride, seed):
    """Token positions of a stage >= 1: every reserve position (multiples of 2*stride) + random others."""
    g = torch.Generator().manual_seed(seed)
    cells = torch.arange(h * w)
    is_res = ((cells % w) % (2 * stride) == 0) & ((cells // w) % (2 * stride) == 0)
    rows = []
    for _ in range(B):
        others = cells[~is_res][torch.randperm(int((~is_res).sum()), generator=g)[:n - int(is_res.sum())]]
        sel = torch.cat([cells[is_res], others])[torch.randperm(n, generator=g)]
        rows.append(torch.stack([sel % w, sel // w], dim=1))
    return torch.stack(rows).float()


@pytest.mark.parametrize("n,stride,ds", [(16384, 2, 0.25), (4096, 4, 0.25), (1024, 8, 0.25), (16384, 2, 0.2), (3276, 4, 0.2), (655, 8, 0.2)])
def test_merge_selection_bit_exact(n, stride, ds):
    """ClusterMerging token selection (aff.py:292-329) incl. heavy score ties (the grid prior is 0/1)."""
    import autofocusformermod_b200 as P
    B, h, w = 2, 128, 128
    pos = inputs.grid_positions(B, h, w) if n == h * w else _stage_tokens(B, n, h, w, stride, seed=n)
    g = torch.Generator().manual_seed(n)
    lp = torch.rand(B, n, 1, generator=g)
    lp[:, ::7] = 0.5                                                       # exact score ties
    reserve_num = math.ceil(h / (stride * 2)) * math.ceil(w / (stride * 2))
    ref = pt.merge_select(pos, lp, stride, 4.0, ds, reserve_num)
    final_prob, reserve_mask = pt.merge_scores(pos, lp, stride, 4.0)
    got = P.merge_select(final_prob.cuda(), reserve_mask.cuda(), int(n * ds), reserve_num)
    assert torch.equal(got.cpu(), ref)


def test_topk_edge_cases():
    import autofocusformermod_b200 as P
    s = torch.tensor([[0.0, -0.0, 1.0, 1.0, -1.0, float("inf"), -float("inf"), 1.0]]).cuda()
    assert P.topk_select(s, 8).cpu().tolist() == [[5, 2, 3, 7, 0, 1, 4, 6]]
    assert P.topk_select(s, 1).cpu().tolist() == [[5]]
    m = torch.tensor([[0, 1, 1, 0, 1.0]]).cuda()
    assert P.mask_select(m, 3).cpu().tolist() == [[1, 2, 4]]
    assert P.mask_select(m, 5).cpu().tolist() == [[1, 2, 4, 0, 0]]           # missing slots are zero-filled
    big = torch.rand(3, 200000, device="cuda")
    ref = torch.sort(big, dim=1, descending=True, stable=True)[1][:, :5000]
    assert torch.equal(P.topk_select(big, 5000), ref)


def test_shepard_upsample_golden(golden_dir):
    import autofocusformermod_b200 as P
    g = np.load(os.path.join(golden_dir, "shepard_16x16_from64.npz"))
    q, d, f = (torch.from_numpy(g[k]).cuda() for k in ("query", "database", "feature"))
    w = P.upsample_feature_shepard(q, d, f, return_weight_only=True)
    assert rel_err(w, torch.from_numpy(g["weights"])) <= 1e-6
    up = P.upsample_feature_shepard(q, d, f)
    assert rel_err(up, torch.from_numpy(g["up"])) <= 1e-5
    assert P.upsample_feature_shepard(d, d, f) is f                           # point_utils.py:97 early-out


def test_inverse_neighbour_list_is_exact():
    """clusten_csr_build: every (i, j) appears exactly once under its row, ascending."""
    import autofocusformermod_b200 as P
    g = torch.Generator().manual_seed(0)
    B, Nq, M, Nk = 2, 3000, 48, 1000
    idx = torch.randint(0, Nk, (B, Nq, M), generator=g)
    off, ent = P.inverse_neighbour_list(idx.cuda(), Nk)
    off, ent = off.cpu().long(), ent.cpu().long() & 0xFFFFFFFF
    for b in range(B):
        flat = idx[b].reshape(-1)
        order = torch.sort(flat, stable=True)[1]
        assert torch.equal(off[b], torch.searchsorted(flat[order], torch.arange(Nk + 1)))
        assert torch.equal(ent[b], ((order // M) << 8) | (order % M))


@pytest.mark.parametrize("stride", [2, 4, 8])
@pytest.mark.parametrize("with_lp,reserve_on", [(True, True), (False, True), (True, False)])
@pytest.mark.parametrize("B,n,hw", [(2, 4096, 128), (3, 1000, 64), (1, 37, 16)])
def test_merge_scores_bit_exact(stride, with_lp, reserve_on, B, n, hw):
    """clusten_merge_scores against the op-by-op formulation of ClusterMerging.forward (aff.py:292-315) on the same device: final_prob
    and reserve_mask bit for bit (they feed a top-k), for the fixed stride of the first merge and the per-token adaptive stride
    2 ** (ceil(log2(d_nearest)) + 1) of the later ones, with / without learned probabilities and reserve tokens."""
    from autofocusformermod_b200 import point_utils as pu
    g = torch.Generator().manual_seed(100 * stride + n)
    cells = torch.stack([torch.randperm(hw * hw, generator=g)[:n] for _ in range(B)])
    pos = torch.stack([cells % hw, cells // hw], dim=-1).float().cuda()           # distinct integer positions
    lp = torch.rand(B, n, 1, generator=g).cuda() if with_lp else None
    alpha = 4.0
    pos_long = pos.long()
    min_dist = None
    if stride == 2:
        grid = ((pos_long % stride) == 0).all(-1).float()
    else:
        _, min_dist = pu.knn_keops(pos, pos, 2, return_dist=True)
        ada = 2 ** (min_dist[:, :, 1].log2().ceil() + 1)
        grid = ((pos_long % ada.unsqueeze(2).long()) == 0).all(-1).float()
    want = grid
    if lp is not None:
        want = want + lp.view(B, n).float() * alpha
    want_r = None
    if reserve_on:
        want_r = ((pos_long % (stride * 2)) == 0).all(dim=-1).float()
        want = want + want_r * (-100)
    got, got_r = pu.merge_scores(pos, min_dist, lp, stride, alpha, reserve_on)
    assert torch.equal(got, want)
    assert (got_r is None and want_r is None) or torch.equal(got_r, want_r)


def test_rel_pos_feature_rows_bit_exact(monkeypatch):
    """clusten_rel_pos_features against the torch formulation of the reference's pre_table rows (aff.py:21-31: dx, dy, dist, dy / dist,
    dx / dist with the centre zeroed), bit for bit -- incl. the centre row (0 / 0), the corner rows and a whole band of the table."""
    from autofocusformermod_b200 import aff, ops
    g = torch.Generator().manual_seed(7)
    centre = 511 * 1023 + 511
    rows = torch.cat([torch.tensor([0, centre, 1023 * 1023 - 1, 1022, centre + 1, centre - 1023]),
                      torch.randint(0, 1023 * 1023, (5000,), generator=g), torch.arange(centre - 3000, centre + 3000)]).cuda()
    before = ops.launch_count()
    got = aff.rel_pos_features(rows)
    assert ops.launch_count() == before + 1
    monkeypatch.setattr(aff, "NATIVE_REL_POS_FEATURES", False)
    want = aff.rel_pos_features(rows)
    assert got.shape == want.shape == (rows.numel(), 5) and got.dtype == torch.float32
    assert torch.equal(got.view(torch.int32), want.view(torch.int32))
    assert torch.equal(aff.rel_pos_features(rows.view(2, -1)[:, :100].contiguous()), want.view(2, -1, 5)[:, :100])      # leading dims
