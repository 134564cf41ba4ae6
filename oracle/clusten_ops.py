"""CPU restatement of the four neighbour-indexed CLUSTEN ops (TEST INFRASTRUCTURE, see oracle/__init__).

Each forward follows the reference kernel body cited next to it, written in the reference's own
"gather, then multiply-and-sum" style (clusten/test_wg_kernel.py:38-39, point_utils.py:116-117).
Backward is PyTorch autograd of these forwards; its semantics were cross-checked against
clustenqk_cuda_kernel.cu:118-128, clustenav_cuda_kernel.cu:117-123,152-156,
clustenwf_cuda_kernel.cu:120-131,161-165 and weighted_gather_cuda_kernel.cu:110-149.

All functions accept any float dtype, compute in that dtype (use float64/float32 inputs to get a
high-precision answer) and run on whatever device the inputs are on (CPU in practice).
"""
import torch


def _gather_rows(x, idx):
    """x: [B, (H,) N, C], idx: [B, Nq, M] int64 -> rows x[b, (h,) idx[b, i, j], :] as [B, (H,) Nq, M, C]."""
    B, Nq, M = idx.shape
    if x.dim() == 3:
        C = x.shape[-1]
        flat = idx.reshape(B, Nq * M, 1).expand(-1, -1, C)
        return x.gather(1, flat).reshape(B, Nq, M, C)
    H, C = x.shape[1], x.shape[-1]
    flat = idx.reshape(B, 1, Nq * M, 1).expand(-1, H, -1, C)
    return x.gather(2, flat).reshape(B, H, Nq, M, C)


def qk_forward(query, key, nbhd_idx):
    """attn[b,h,i,j] = sum_c query[b,h,i,c] * key[b,h,nbhd_idx[b,i,j],c]
    (clustenqk_cuda_kernel.cu:38-45; dtype rule clusten.py:27-28: key is cast to query's dtype)."""
    key = key.to(query.dtype)
    kg = _gather_rows(key, nbhd_idx)                       # B H N M C
    return (query.unsqueeze(3) * kg).sum(-1)               # B H N M


def av_forward(attn, v, nbhd_idx):
    """feat[b,h,i,c] = sum_j attn[b,h,i,j] * v[b,h,nbhd_idx[b,i,j],c]
    (clustenav_cuda_kernel.cu:40-46; clusten.py:54-55: v is cast to attn's dtype)."""
    v = v.to(attn.dtype)
    vg = _gather_rows(v, nbhd_idx)                         # B H N M C
    return (attn.unsqueeze(-1) * vg).sum(3)                # B H N C


def wf_forward(weights, feat, nbhd_idx):
    """feat_new[b,i,ic,c] = sum_j weights[b,i,j,ic] * feat[b,nbhd_idx[b,i,j],c]
    (clustenwf_cuda_kernel.cu:41-49; clusten.py:80-81: feat is cast to weights' dtype)."""
    feat = feat.to(weights.dtype)
    fg = _gather_rows(feat, nbhd_idx)                      # B N' M C
    return torch.einsum("bnmi,bnmc->bnic", weights, fg)    # B N' IC C


def wg_forward(nbhd_idx, weights, feat):
    """feat_new[b,i,c] = sum_k weights[b,i,k] * feat[b,nbhd_idx[b,i,k],c]
    (weighted_gather_cuda_kernel.cu:38-45, restated exactly as point_utils.py:116-117;
    clusten.py:106-107: weights are cast to feat's dtype)."""
    weights = weights.to(feat.dtype)
    fg = _gather_rows(feat, nbhd_idx)                      # B N K C
    return fg.mul(weights.unsqueeze(3)).sum(dim=2)


class _Apply:
    """Tiny stand-in with the ``.apply`` calling convention of the reference autograd Functions
    (clusten.py:19-120) so reference/oracle model code can be driven by the CPU restatements."""

    def __init__(self, fn):
        self.apply = fn


CLUSTENQKFunction = _Apply(qk_forward)
CLUSTENAVFunction = _Apply(av_forward)
CLUSTENWFFunction = _Apply(wf_forward)
WEIGHTEDGATHERFunction = _Apply(wg_forward)


def msdetrpc_forward(nn_idx, nn_weight, attn, val):
    """feat[b,i,c] = sum_m attn[b,i,m] * sum_k nn_weight[b,i,m,k] * val[b,nn_idx[b,i,m,k],c]
    (msdetrpc_cuda_kernel.cu:18-54, torch form of test_msdetrpc_kernel.py:41-42). "next" row f-1."""
    B, N, M, K = nn_idx.shape
    g = _gather_rows(val, nn_idx.reshape(B, N, M * K)).reshape(B, N, M, K, -1)
    return ((g * nn_weight.unsqueeze(-1)).sum(3) * attn.unsqueeze(-1)).sum(2)


MSDETRPCFunction = _Apply(msdetrpc_forward)


def fwd_bwd(fn, tensors, grad_out):
    """Run ``fn(*tensors)`` and autograd-backprop ``grad_out``; returns (out, [grads of float inputs])."""
    leaves = []
    args = []
    for t in tensors:
        if t.is_floating_point():
            t = t.detach().clone().requires_grad_(True)
            leaves.append(t)
        args.append(t)
    out = fn(*args)
    out.backward(grad_out.to(out.dtype))
    return out.detach(), [t.grad for t in leaves]
