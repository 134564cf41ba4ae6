"""One-shot check of clusten_knn against pykeops (for an integrator who HAS pykeops 2.1.1 and a GPU; neither the authoring
container nor the B200 box has it, so kNN parity is 'unpinned' in DESIGN.md until this has been run somewhere).

For each kNN call site of the reference (tokens -> cluster centres k=6, self kNN k=2 with distances, PointConv
self kNN k=9, grid -> level k=4, upsample k=4) it builds the inputs the backbone would see on a 128x128 stem grid, runs

    reference:  LazyTensor(query[:, :, None, :]) / LazyTensor(database[:, None, :, :]),  ((q - d) ** 2).sum(-1).sqrt().argKmin(k, dim=2)
                (backbone/point_utils.py:41-60)
    ours:       autofocusformermod_b200.knn_keops(query, database, k)        (canonical rule: ascending distance, ties -> lowest index)

and prints: identical / identical as sets per query / different, next to the number of queries whose answer contains a distance tie
(those are the only places where a different tie rule can show).

    python tools/knn_vs_pykeops.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def keops_knn(query, database, k, return_dist=False):
    from pykeops.torch import LazyTensor
    q, d = LazyTensor(query[:, :, None, :].contiguous()), LazyTensor(database[:, None, :, :].contiguous())
    dist = ((q - d) ** 2).sum(-1).sqrt()
    if return_dist:
        dd, idx = dist.Kmin_argKmin(k, dim=2)
        return idx, dd
    return dist.argKmin(k, dim=2)


def main():
    import autofocusformermod_b200 as P
    try:
        import pykeops  # noqa: F401
    except ImportError:
        raise SystemExit("pykeops is not installed: nothing to compare against (pip install pykeops==2.1.1)")
    g = torch.Generator().manual_seed(0)
    h = w = 128
    ys, xs = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    grid = torch.stack([xs, ys], dim=2).reshape(1, -1, 2).float().cuda()
    sub = lambda n: grid[:, torch.randperm(h * w, generator=g)[:n].cuda()]
    spos, mean, _, _, _ = P.space_filling_cluster(grid, 8, h, w)
    cases = {"tokens->clusters k=6 (aff.py:475)": (spos, mean, 6), "self k=2 + dist (aff.py:299)": (sub(4096), None, 2),
             "PointConv self k=9 (msdeformattn_pc.py:295)": (grid, None, 9), "grid->level k=4 (msdeformattn_pc.py:502)": (grid, sub(4096), 4),
             "upsample k=4 (point_utils.py:103)": (grid, sub(1024), 4)}
    for name, (q, d, k) in cases.items():
        d = q if d is None else d
        ours, od = P.knn_keops(q, d, k, return_dist=True)
        ref, rd = keops_knn(q, d, k, return_dist=True)
        same = (ours == ref).all(-1)
        same_set = (ours.sort(-1)[0] == ref.sort(-1)[0]).all(-1)
        _, d2 = P.knn_keops(q, d, min(k + 1, d.shape[1]), return_dist=True)
        tied = ((d2[..., 1:] == d2[..., :-1]).any(-1)).sum()
        print(f"{name:48s} identical {int(same.sum())}/{same.numel()}  same set {int(same_set.sum())}  queries with a tie {int(tied)}  "
              f"distances bit-equal {bool(torch.equal(od, rd))}")


if __name__ == "__main__":
    main()
