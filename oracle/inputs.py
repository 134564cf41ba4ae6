"""Seeded synthetic inputs for the parity tests and the bench (TEST INFRASTRUCTURE).

SURVEY.md section 8(d): everything comes from ``torch.Generator().manual_seed(seed)`` on CPU.
"""
import torch

from . import point_ops as pt


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


def structured_neighbourhood(B, n, h, w, m=8, nbhd=48, seed=0):
    """Positions -> reference-style clustering + nearest clusters -> nbhd_idx [B,n,M] (+mask, pe_idx).
    This is what the CLUSTEN ops receive inside the backbone (aff.py:469-485)."""
    pos = grid_positions(B, h, w) if n == h * w else random_positions(B, n, h, w, seed)
    pos, mean_pos, member, cmask, _ = pt.space_filling_cluster(pos, m, h, w)
    k = member.shape[1]
    nnc = min(int(round(nbhd / float(m))), k)
    nb, mask, pe_idx = pt.assemble_neighbourhood(pos, mean_pos, member, cmask, nnc)
    return pos, nb, mask, pe_idx


def random_neighbourhood(B, nq, nkv, M, seed=0):
    """Adversarial index tensor: uniform random with duplicates (stresses the scatter side)."""
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, nkv, (B, nq, M), generator=g)


def qkv_case(B=2, H=2, N=4096, C=32, M=48, seed=0, structured=True, dtype=torch.float32):
    """BASELINE config 1: q,k,v ~ N(0,1) [B,H,N,C]; attn = softmax(randn); upstream grads ~ N(0,1)."""
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(B, H, N, C, generator=g)
    k = torch.randn(B, H, N, C, generator=g)
    v = torch.randn(B, H, N, C, generator=g)
    if structured:
        side = int(round(N ** 0.5))
        assert side * side == N or True
        h = w = max(side * 2, 8)
        idx = structured_neighbourhood(B, N, h, w, 8, M, seed)[1]
        M = idx.shape[-1]
    else:
        idx = random_neighbourhood(B, N, N, M, seed)
    attn = torch.randn(B, H, N, M, generator=g).softmax(-1)
    d_attn = torch.randn(B, H, N, M, generator=g)
    d_feat = torch.randn(B, H, N, C, generator=g)
    cast = lambda t: t.to(dtype).to(torch.float32) if dtype != torch.float32 else t
    return dict(q=cast(q), k=cast(k), v=cast(v), idx=idx, attn=cast(attn), d_attn=cast(d_attn), d_feat=cast(d_feat))


def wf_case(B=2, Nq=1024, N=4096, C=64, M=48, IC=4, seed=0, dtype=torch.float32):
    """BASELINE config 1, WF part: w ~ N(0,1) [B,Nq,M,IC], f [B,N,C], idx [B,Nq,M]."""
    g = torch.Generator().manual_seed(seed)
    w = torch.randn(B, Nq, M, IC, generator=g)
    f = torch.randn(B, N, C, generator=g)
    idx = torch.randint(0, N, (B, Nq, M), generator=g)
    d_out = torch.randn(B, Nq, IC, C, generator=g)
    cast = lambda t: t.to(dtype).to(torch.float32) if dtype != torch.float32 else t
    return dict(w=cast(w), f=cast(f), idx=idx, d_out=cast(d_out))
