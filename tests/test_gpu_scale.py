"""GPU parity AT THE SHAPES THE BENCH RUNS: the fused attention entry points (clusten_attn_fwd / clusten_attn_bwd + scatter +
table gradient, and the in-kernel-bias variant) and the WF merge (CLUSTENWFFunction) at AFF-Small stage 0 (B = 8, N = 16 384,
M = 48) and AFF-Base stage 0 (N = 32 768, m = 24, M = 144: padded last cluster -> mask), fp32 and bf16.

The CPU oracle would take minutes at these sizes, so the checker here is the reference's gather formulation
(clustenqk_cuda_kernel.cu:38-45, clustenav_cuda_kernel.cu:40-46, clustenwf_cuda_kernel.cu:41-49 and the glue of
backbone/aff.py:114-155) written in plain torch and evaluated in FLOAT64 on the same device, forward and autograd backward, over
the WHOLE tensors (not a sample of rows) -- the same composition tests/test_gpu_ops.py pins against oracle/ at small sizes.
Index tensors come from the product's own clustering / kNN / stage-prepare pipeline (bit-exact against the oracle in
tests/test_gpu_integer.py), so they have the real run structure, the real unions and the real padded clusters.
Tolerances: 1e-5 fp32, 1e-2 bf16 (max |a-b| / max |b|)."""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu

SHAPES = {
    # name: (B, H, C, grid h, grid w, m, nbhd)
    "small_s0": (8, 3, 32, 128, 128, 8, 48),
    "base_s0": (2, 4, 32, 128, 256, 24, 144),
}


def _structure(B, h, w, m, nbhd):
    from autofocusformermod_b200 import point_utils as pu
    ys, xs = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    pos = torch.stack([xs, ys], dim=2).reshape(1, -1, 2).float().cuda()
    spos, mean_pos, member, cmask, _ = pu.space_filling_cluster(pos, m, h, w)
    nnc = min(int(round(nbhd / float(m))), member.shape[1])
    nearest = pu.knn_keops(spos, mean_pos, nnc)
    idx, _, mask8, uniq, bias_idx = pu.stage_prepare(spos, nearest, member, cmask, want_mask64=False)
    ex = lambda t: None if t is None else t.expand(B, -1, -1).contiguous()
    return ex(spos), ex(idx), ex(mask8), uniq, ex(bias_idx)


def _attention_reference(q, kv, tab, bk, bv, idx, bias_idx, mask):
    """aff.py:114-155 in float64 torch: q [B,N,H,C], kv [B,N,H,2,C] -> out [B,N,H*C] (chunked over the batch to bound memory)."""
    B, N, H, C = q.shape
    outs = []
    for b in range(B):
        k = kv[b, :, :, 0].permute(1, 0, 2)                                  # H N C
        v = kv[b, :, :, 1].permute(1, 0, 2)
        qq = q[b].permute(1, 0, 2)
        kn = k[:, idx[b]]                                                    # H N M C
        attn = (qq.unsqueeze(2) * kn).sum(-1)                                # clustenqk_cuda_kernel.cu:38-45
        attn = attn + tab[bias_idx[b].long()].permute(2, 0, 1)               # aff.py:129-134
        if mask is not None:
            attn = attn + (1 - mask[b].double()).unsqueeze(0) * (-100)       # aff.py:137
        blank = (qq * bk.view(H, 1, C)).sum(-1, keepdim=True)                # aff.py:140
        p = torch.cat([attn, blank], dim=-1).softmax(-1)
        o = (p[..., :-1].unsqueeze(-1) * v[:, idx[b]]).sum(2) + p[..., -1:] * bv.view(H, 1, C)   # clustenav_cuda_kernel.cu:40-46, aff.py:148-155
        outs.append(o.permute(1, 0, 2).reshape(N, H * C))
    return torch.stack(outs)


def _inputs(name, dtype):
    B, H, C, h, w, m, nbhd = SHAPES[name]
    pos, idx, mask8, uniq, bias_idx = _structure(B, h, w, m, nbhd)
    N = idx.shape[1]
    g = torch.Generator(device="cuda").manual_seed(11)
    rnd = lambda *s: torch.randn(*s, device="cuda", generator=g)
    q = (rnd(B, N, H, C) * C ** -0.5).to(dtype)
    kv = rnd(B, N, H, 2, C).to(dtype)
    tab = rnd(uniq.numel(), H)
    bk, bv = rnd(H * C).to(dtype), rnd(H * C).to(dtype)
    go = rnd(B, N, H * C).to(dtype)
    return pos, idx, mask8, bias_idx, q, kv, tab, bk, bv, go


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
@pytest.mark.parametrize("name", list(SHAPES))
def test_fused_attention_forward_at_bench_scale(name, dtype, attn_fwd_kernel):
    """clusten_attn_fwd through the inference entry (ops.cluster_attention_fused, what AFF.forward calls under no_grad)."""
    from autofocusformermod_b200 import ops
    pos, idx, mask8, bias_idx, q, kv, tab, bk, bv, _ = _inputs(name, dtype)
    kvp = kv.permute(3, 0, 2, 1, 4)
    out = ops.cluster_attention_fused(q.permute(0, 2, 1, 3), kvp[0], kvp[1], idx, tab, bias_idx, mask8, bk, bv)
    ref = _attention_reference(q.double(), kv.double(), tab.double(), bk.double(), bv.double(), idx, bias_idx, mask8)
    e = rel_err(out, ref)
    print(name, dtype, "fused fwd rel err", f"{e:.2e}", "pack", ops.pack_flags(idx, idx.shape[1], mask=mask8)[:7])
    assert e <= (1e-5 if dtype == torch.float32 else 1e-2)
    assert ops.pack_flags(idx, idx.shape[1], mask=mask8)[0] == 0          # the tile-union path, not the generic kernel, was measured


@pytest.mark.parametrize("name", list(SHAPES))
def test_fused_attention_training_at_bench_scale(name):
    """ClusterAttentionCoreFunction (clusten_attn_fwd + clusten_attn_bwd + clusten_scatter_rows x2 + clusten_table_grad +
    clusten_blank_grad), bf16: out and the gradients of q, kv, the bias table and the blank-token parameters."""
    from autofocusformermod_b200 import ops
    dtype = torch.bfloat16
    pos, idx, mask8, bias_idx, q, kv, tab, bk, bv, go = _inputs(name, dtype)
    leaves = [t.detach().clone().requires_grad_(True) for t in (q, kv, tab, bk, bv)]
    out = ops.cluster_attention_core(*leaves, idx, bias_idx, mask8)
    out.backward(go)
    refs = [t.detach().double().requires_grad_(True) for t in (q, kv, tab, bk, bv)]
    ref = _attention_reference(*refs, idx, bias_idx, mask8)
    ref.backward(go.double())
    errs = {"out": rel_err(out, ref)}
    for nm, a, r in zip(("d_q", "d_kv", "d_bias_tab", "d_blank_k", "d_blank_v"), leaves, refs):
        errs[nm] = rel_err(a.grad, r.grad)
    print(name, {k: f"{v:.2e}" for k, v in errs.items()})
    assert errs["out"] <= 1e-2 and max(errs.values()) <= 2e-2, errs


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
@pytest.mark.parametrize("name", list(SHAPES))
def test_fused_attention_inkernel_bias_at_bench_scale(name, dtype, attn_fwd_kernel):
    """clusten_attn_pos_fwd (bias from positions, posbias.cuh) against the table formulation evaluated in float64: the bias the
    reference gathers is pos_embed(pre_table)[pe_idx] (aff.py:17-31,129-132,481-485)."""
    from autofocusformermod_b200 import ops
    from autofocusformermod_b200.aff import rel_pos_features, TABLE_WIDTH, REL_POS_WIDTH
    pos, idx, mask8, bias_idx, q, kv, _, bk, bv, _ = _inputs(name, dtype)
    B, N, H, C = q.shape
    g = torch.Generator(device="cuda").manual_seed(5)
    pe_w = torch.randn(H, 5, device="cuda", generator=g) * 0.2
    pe_b = torch.randn(H, device="cuda", generator=g)
    kvp = kv.permute(3, 0, 2, 1, 4)
    out = ops.cluster_attention_fused_pos(q.permute(0, 2, 1, 3), kvp[0], kvp[1], idx, pos, pe_w, pe_b, mask8, bk, bv)
    # aff.py:481-485: table row of every (token, neighbour), then the reference's table features through Linear(5, H)
    rel = pos.gather(1, idx.reshape(B, -1, 1).expand(-1, -1, 2)).reshape(B, N, -1, 2) - (pos.unsqueeze(2) - REL_POS_WIDTH)
    rel = rel.clamp(0, TABLE_WIDTH - 1).long()
    pe_idx = rel[..., 1] * TABLE_WIDTH + rel[..., 0]
    uniq, inv = torch.unique(pe_idx, return_inverse=True)
    tab = rel_pos_features(uniq).double() @ pe_w.double().t() + pe_b.double()
    ref = _attention_reference(q.double(), kv.double(), tab, bk.double(), bv.double(), idx, inv.int(), mask8)
    e = rel_err(out, ref)
    print(name, dtype, "in-kernel bias fwd rel err", f"{e:.2e}")
    assert e <= (1e-5 if dtype == torch.float32 else 1e-2)


MERGES = {
    # name: (B, C, grid h, grid w, m, nbhd, keep)  -- ClusterMerging after stage 0 (aff.py:332-361)
    "small_merge0": (8, 96, 128, 128, 8, 48, 4096),
    "base_merge0": (2, 128, 128, 256, 24, 144, 8192),
}


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
@pytest.mark.parametrize("name", list(MERGES))
def test_wf_merge_at_bench_scale(name, dtype):
    """CLUSTENWFFunction forward + both gradients at the merge shapes: kept tokens in selection order (no curve locality),
    their neighbourhoods keep the cluster run structure (aff.py:335)."""
    import autofocusformermod_b200 as P
    B, C, h, w, m, nbhd, keep = MERGES[name]
    _, nb, mask8, _, _ = _structure(B, h, w, m, nbhd)
    N, M = nb.shape[1], nb.shape[2]
    g = torch.Generator(device="cuda").manual_seed(3)
    sel = torch.stack([torch.randperm(N, device="cuda", generator=g)[:keep] for _ in range(B)])
    idx = nb.gather(1, sel.unsqueeze(2).expand(-1, -1, M)).contiguous()
    wts = torch.randn(B, keep, M, 4, device="cuda", generator=g).to(dtype)
    if mask8 is not None:                                  # padded slots carry zero weight (aff.py:352-358)
        wts = wts * mask8.gather(1, sel.unsqueeze(2).expand(-1, -1, M)).unsqueeze(3).to(dtype)
    feat = torch.randn(B, N, C, device="cuda", generator=g).to(dtype)
    go = torch.randn(B, keep, 4, C, device="cuda", generator=g).to(dtype)
    w1, f1 = wts.clone().requires_grad_(True), feat.clone().requires_grad_(True)
    out = P.CLUSTENWFFunction.apply(w1, f1, idx)
    out.backward(go)
    w2, f2 = wts.double().requires_grad_(True), feat.double().requires_grad_(True)
    refs = []
    for b in range(B):                                     # clustenwf_cuda_kernel.cu:41-49: out[i,ic,c] = sum_j w[i,j,ic] f[idx[i,j],c]
        refs.append(torch.einsum("ijk,ijc->ikc", w2[b], f2[b][idx[b]]))
    ref = torch.stack(refs)
    ref.backward(go.double())
    errs = {"out": rel_err(out, ref), "d_w": rel_err(w1.grad, w2.grad), "d_f": rel_err(f1.grad, f2.grad)}
    print(name, dtype, {k: f"{v:.2e}" for k, v in errs.items()})
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    assert max(errs.values()) <= tol, errs
