"""Index tensors for the benchmarks, produced by the PRODUCT pipeline on the GPU (space_filling_cluster -> knn_keops ->
stage_prepare, i.e. what BasicLayer.forward runs, backbone/aff.py:469-485): nothing here touches oracle/
(the one checker-side option of the benchmarks is ``op_bench.py --ref``, which times the reference's own kernels beside ours)."""
import torch


def grid_positions(B, h, w):
    ys, xs = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    return torch.stack([xs, ys], dim=2).reshape(1, -1, 2).float().expand(B, -1, -1).contiguous()


def random_positions(B, n, h, w, seed=0):
    """n distinct integer positions per sample on an h x w grid (what AFF stages >= 1 see)."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(B):
        perm = torch.randperm(h * w, generator=g)[:n]
        out.append(torch.stack([perm % w, perm // w], dim=1))
    return torch.stack(out).float()


def stage_structure(B, n, h, w, m=8, nbhd=48, seed=0):
    """(pos [B,n,2] fp32 in cluster order, nbhd_idx int64 [B,n,M], mask uint8 [B,n,M] or None, uniq int64 [U], bias_idx int32
    [B,n,M]) on the GPU, for n tokens on an h x w stem grid (all of it when n == h * w, a random subset otherwise)."""
    from autofocusformermod_b200 import point_utils as pu
    pos = (grid_positions(B, h, w) if n == h * w else random_positions(B, n, h, w, seed)).cuda()
    spos, mean_pos, member, cmask, _ = pu.space_filling_cluster(pos, m, h, w)
    k = member.shape[1]
    nnc = min(int(round(nbhd / float(m))), k)
    nearest = pu.knn_keops(spos, mean_pos, nnc)
    member_idx, _, mask8, uniq, bias_idx = pu.stage_prepare(spos, nearest, member, cmask, want_mask64=False)
    return spos, member_idx, mask8, uniq, bias_idx
