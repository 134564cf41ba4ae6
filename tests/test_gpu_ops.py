"""GPU parity: the four CLUSTEN autograd Functions (ctypes -> C ABI -> sm_100a kernels) against the CPU oracle on the
same seeded inputs.  Tolerances (BASELINE north_star): max|a-b|/max|b| <= 1e-5 fp32, <= 1e-2 bf16/fp16 (vs the fp32
oracle on inputs rounded to the low-precision type)."""
import os

import pytest
import torch

from oracle import clusten_ops as co
from oracle import inputs

from conftest import rel_err

pytestmark = pytest.mark.gpu

TOL = {torch.float32: 1e-5, torch.bfloat16: 1e-2, torch.float16: 1e-2}


def _ops():
    import autofocusformermod_b200 as pkg
    return pkg


def _run(fn, tensors, grad_out, dtype):
    """our Function on cuda in `dtype`; returns fp32 cpu (out, grads of float inputs)."""
    args, leaves = [], []
    for t in tensors:
        if t.is_floating_point():
            t = t.to("cuda", dtype).requires_grad_(True)
            leaves.append(t)
        else:
            t = t.cuda()
        args.append(t)
    out = fn(*args)
    out.backward(grad_out.to("cuda", dtype))
    torch.cuda.synchronize()
    return out.detach().float().cpu(), [l.grad.float().cpu() for l in leaves]


def _check(got, ref, dtype, what):
    out, grads = got
    rout, rgrads = ref
    assert out.shape == rout.shape, what
    e = rel_err(out, rout)
    assert e <= TOL[dtype], f"{what} forward rel err {e:.3e}"
    for n, (g, r) in enumerate(zip(grads, rgrads)):
        assert g.shape == r.shape
        e = rel_err(g, r)
        assert e <= TOL[dtype], f"{what} grad[{n}] rel err {e:.3e}"


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16], ids=["f32", "bf16", "f16"])
@pytest.mark.parametrize("structured", [True, False], ids=["clustered-idx", "random-idx"])
def test_qk_av_config1(dtype, structured):
    """BASELINE configs[0]: B=2, N=4096, heads=2, C=32, 48 neighbours."""
    P = _ops()
    c = inputs.qkv_case(B=2, H=2, N=4096, C=32, M=48, seed=0, structured=structured, dtype=dtype)
    ref = co.fwd_bwd(co.qk_forward, [c["q"], c["k"], c["idx"]], c["d_attn"])
    _check(_run(P.CLUSTENQKFunction.apply, [c["q"], c["k"], c["idx"]], c["d_attn"], dtype), ref, dtype, "QK")
    ref = co.fwd_bwd(co.av_forward, [c["attn"], c["v"], c["idx"]], c["d_feat"])
    _check(_run(P.CLUSTENAVFunction.apply, [c["attn"], c["v"], c["idx"]], c["d_feat"], dtype), ref, dtype, "AV")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
def test_wf_config1(dtype):
    P = _ops()
    c = inputs.wf_case(B=2, Nq=1024, N=4096, C=64, M=48, IC=4, seed=0, dtype=dtype)
    ref = co.fwd_bwd(co.wf_forward, [c["w"], c["f"], c["idx"]], c["d_out"])
    _check(_run(P.CLUSTENWFFunction.apply, [c["w"], c["f"], c["idx"]], c["d_out"], dtype), ref, dtype, "WF")


SHAPES_QK = [
    # B, H, N, C, M      what it covers
    (1, 1, 1, 4, 1),     # smallest
    (2, 3, 37, 24, 48),  # C=24 (AFF-Mini stage 3), ragged N
    (1, 4, 300, 16, 48), # C=16 (AFF-Mini stage 0)
    (2, 2, 257, 32, 144),# M=144 (AFF-Base)
    (1, 2, 50, 20, 7),   # C not a multiple of 8 -> bf16 scalar path, fp32 vector path
    (1, 2, 33, 5, 9),    # odd C -> scalar path
    (1, 1, 64, 128, 48), # C=128: full-warp rows
    (1, 1, 40, 256, 12), # C=256 > 32 chunks in fp32 -> scalar path
]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
@pytest.mark.parametrize("B,H,N,C,M", SHAPES_QK)
def test_qk_av_shapes(B, H, N, C, M, dtype):
    P = _ops()
    g = torch.Generator().manual_seed(B * 1000 + N + C + M)
    rnd = lambda *s: torch.randn(*s, generator=g).to(dtype).float()
    q, k, v = rnd(B, H, N, C), rnd(B, H, N, C), rnd(B, H, N, C)
    idx = torch.randint(0, N, (B, N, M), generator=g)
    attn, d_attn, d_feat = rnd(B, H, N, M).softmax(-1).to(dtype).float(), rnd(B, H, N, M), rnd(B, H, N, C)
    _check(_run(P.CLUSTENQKFunction.apply, [q, k, idx], d_attn, dtype), co.fwd_bwd(co.qk_forward, [q, k, idx], d_attn), dtype, "QK")
    _check(_run(P.CLUSTENAVFunction.apply, [attn, v, idx], d_feat, dtype), co.fwd_bwd(co.av_forward, [attn, v, idx], d_feat), dtype, "AV")


def test_qk_av_cross_sizes_and_unreferenced_rows():
    """Nk != Nq; key rows nobody points at must get exactly zero gradient (the reference zero-fills d_key)."""
    P = _ops()
    g = torch.Generator().manual_seed(4)
    B, H, Nq, Nk, C, M = 2, 2, 100, 260, 32, 16
    q, k = torch.randn(B, H, Nq, C, generator=g), torch.randn(B, H, Nk, C, generator=g)
    idx = torch.randint(0, 200, (B, Nq, M), generator=g)          # rows 200..259 unreferenced
    d_attn = torch.randn(B, H, Nq, M, generator=g)
    got = _run(P.CLUSTENQKFunction.apply, [q, k, idx], d_attn, torch.float32)
    _check(got, co.fwd_bwd(co.qk_forward, [q, k, idx], d_attn), torch.float32, "QK cross")
    assert float(got[1][1][:, :, 200:].abs().max()) == 0.0


def test_reference_model_views_are_consumed_without_copies():
    """q / key / v exactly as aff.py:103-113 builds them (permuted, non-contiguous views) and attn[..., :-1] slice
    (aff.py:146) -- results must equal the contiguous call bit for bit, and grads flow back to the packed kv."""
    P = _ops()
    torch.manual_seed(0)
    b, n, h, c_ = 2, 512, 4, 32
    c = inputs.qkv_case(B=b, H=h, N=n, C=c_, M=48, seed=2, structured=False)
    idx = c["idx"].cuda()
    qf = torch.randn(b, n, h * c_, device="cuda", requires_grad=True)
    kvf = torch.randn(b, n, 2 * h * c_, device="cuda", requires_grad=True)
    q = qf.reshape(b, n, h, c_).permute(0, 2, 1, 3)
    kv = kvf.view(b, n, h, 2, c_).permute(3, 0, 2, 1, 4)
    key, v = kv[0], kv[1]
    assert not q.is_contiguous() and not key.is_contiguous()
    attn = P.CLUSTENQKFunction.apply(q, key, idx)
    attn_c = P.CLUSTENQKFunction.apply(q.contiguous(), key.contiguous(), idx)
    assert torch.equal(attn, attn_c)
    sm = torch.cat([attn, torch.zeros_like(attn[..., :1])], -1).softmax(-1)
    feat = P.CLUSTENAVFunction.apply(sm[..., :-1], v, idx)
    feat_c = P.CLUSTENAVFunction.apply(sm[..., :-1].contiguous(), v.contiguous(), idx)
    assert torch.equal(feat, feat_c)
    out = feat.permute(0, 2, 1, 3).reshape(b, n, h * c_)
    out.square().sum().backward()
    # oracle on CPU, same graph
    qf2, kvf2 = qf.detach().cpu().requires_grad_(True), kvf.detach().cpu().requires_grad_(True)
    q2 = qf2.reshape(b, n, h, c_).permute(0, 2, 1, 3)
    kv2 = kvf2.view(b, n, h, 2, c_).permute(3, 0, 2, 1, 4)
    a2 = co.qk_forward(q2, kv2[0], c["idx"])
    sm2 = torch.cat([a2, torch.zeros_like(a2[..., :1])], -1).softmax(-1)
    f2 = co.av_forward(sm2[..., :-1], kv2[1], c["idx"])
    f2.permute(0, 2, 1, 3).reshape(b, n, h * c_).square().sum().backward()
    assert rel_err(qf.grad, qf2.grad) <= 2e-5
    assert rel_err(kvf.grad, kvf2.grad) <= 2e-5


def test_backward_is_deterministic():
    P = _ops()
    c = inputs.qkv_case(B=2, H=2, N=2048, C=32, M=48, seed=5, structured=False)
    a = _run(P.CLUSTENQKFunction.apply, [c["q"], c["k"], c["idx"]], c["d_attn"], torch.float32)
    b = _run(P.CLUSTENQKFunction.apply, [c["q"], c["k"], c["idx"]], c["d_attn"], torch.float32)
    assert all(torch.equal(x, y) for x, y in zip(a[1], b[1]))


SHAPES_WF = [
    # B, Nq, N, C, M, IC
    (1, 1, 1, 4, 1, 1),
    (2, 100, 400, 32, 48, 4),    # merge 0 of AFF-Mini
    (2, 77, 300, 128, 48, 4),
    (1, 64, 64, 256, 9, 4),      # PointConv (msdeformattn_pc.py:309)
    (1, 50, 200, 384, 48, 4),    # C=384 (Mini stage 3 width)
    (2, 31, 90, 20, 5, 3),       # IC=3 -> scalar path
    (1, 40, 80, 7, 4, 2),        # odd C -> scalar path
    (1, 40, 300, 96, 144, 4),    # Base M=144
]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
@pytest.mark.parametrize("B,Nq,N,C,M,IC", SHAPES_WF)
def test_wf_shapes(B, Nq, N, C, M, IC, dtype):
    P = _ops()
    c = inputs.wf_case(B=B, Nq=Nq, N=N, C=C, M=M, IC=IC, seed=Nq + C, dtype=dtype)
    ref = co.fwd_bwd(co.wf_forward, [c["w"], c["f"], c["idx"]], c["d_out"])
    _check(_run(P.CLUSTENWFFunction.apply, [c["w"], c["f"], c["idx"]], c["d_out"], dtype), ref, dtype, "WF")


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32], ids=["bf16", "f16", "f32"])
@pytest.mark.parametrize("C", [32, 96, 128, 192, 48, 512])
@pytest.mark.parametrize("n,m,nbhd,hw", [(4096, 8, 48, 64), (2003, 8, 48, 64), (655, 8, 48, 128), (1540, 24, 144, 64)])
def test_wf_merge_structured(n, m, nbhd, hw, C, dtype):
    """The PointConv merge as the backbone runs it (aff.py:332-361): neighbourhoods of the reference clustering,
    gathered for a shuffled quarter of the tokens (top-k order).  16-bit types take the tensor-core kernels and the
    WF plan (perm + per-octet lists); padded last clusters (n % m != 0) exercise the impure-slot fix-up."""
    P = _ops()
    from autofocusformermod_b200 import ops
    B = 2
    _, nb, _, _ = inputs.structured_neighbourhood(B, n, hw, hw, m, nbhd, seed=n)
    g = torch.Generator().manual_seed(n + C)
    Nq = n // 4
    sel = torch.stack([torch.randperm(n, generator=g)[:Nq] for _ in range(B)])
    if n % m:                                       # make sure tokens next to the padded cluster are kept
        sel[:, :8] = torch.arange(n - 8, n)
    idx = nb.gather(1, sel.unsqueeze(-1).expand(-1, -1, nb.shape[-1])).contiguous()
    M = idx.shape[-1]
    cast = lambda t: t.to(dtype).float()
    w, f = cast(torch.randn(B, Nq, M, 4, generator=g)), cast(torch.randn(B, n, C, generator=g))
    d_out = cast(torch.randn(B, Nq, 4, C, generator=g))
    ref = co.fwd_bwd(co.wf_forward, [w, f, idx], d_out)
    _check(_run(P.CLUSTENWFFunction.apply, [w, f, idx], d_out, dtype), ref, dtype, "WF merge")
    flags = ops.wf_plan_flags(idx.cuda(), n)
    assert flags[0] == 0, flags                      # octet path, not the generic one
    assert (flags[1] > 0) == (n % m != 0), flags     # impure slots exactly when the last cluster is padded


def test_wf_plan_routes_unstructured_idx_to_generic_path():
    from autofocusformermod_b200 import ops
    idx = inputs.random_neighbourhood(2, 300, 1000, 48, seed=1).cuda()
    assert ops.wf_plan_flags(idx, 1000)[0] == 1
    assert ops.wf_plan(torch.zeros(1, 4, 9, dtype=torch.int64, device="cuda"), 10) is None      # M % 8 != 0: no plan


def test_wf_backward_is_deterministic():
    P = _ops()
    _, nb, _, _ = inputs.structured_neighbourhood(2, 2003, 64, 64, 8, 48, seed=5)
    g = torch.Generator().manual_seed(0)
    sel = torch.stack([torch.randperm(2003, generator=g)[:500] for _ in range(2)])
    idx = nb.gather(1, sel.unsqueeze(-1).expand(-1, -1, 48)).contiguous().cuda()
    w = torch.randn(2, 500, 48, 4, generator=g).cuda().bfloat16()
    f = torch.randn(2, 2003, 96, generator=g).cuda().bfloat16()
    go = torch.randn(2, 500, 4, 96, generator=g).cuda().bfloat16()
    res = []
    for _ in range(3):
        idx_i = idx.clone()                          # fresh plan every time (the list order must not depend on atomics)
        wi, fi = w.clone().requires_grad_(True), f.clone().requires_grad_(True)
        P.CLUSTENWFFunction.apply(wi, fi, idx_i).backward(go)
        res.append((wi.grad.clone(), fi.grad.clone()))
    for a, b in res[1:]:
        assert torch.equal(a, res[0][0]) and torch.equal(b, res[0][1])


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
@pytest.mark.parametrize("B,Nq,N,C,K", [(3, 50, 100, 32, 4), (2, 1024, 256, 256, 4), (1, 300, 70, 100, 4), (1, 9, 5, 3, 2)])
def test_weighted_gather_shapes(B, Nq, N, C, K, dtype):
    """includes the reference's own test shape family (clusten/test_wg_kernel.py: n=50, n_=100, k=4, c=32)."""
    P = _ops()
    g = torch.Generator().manual_seed(Nq)
    idx = torch.randint(0, N, (B, Nq, K), generator=g)
    w = torch.rand(B, Nq, K, generator=g).to(dtype).float()
    f = torch.rand(B, N, C, generator=g).to(dtype).float()
    d_out = torch.randn(B, Nq, C, generator=g).to(dtype).float()
    ref = co.fwd_bwd(co.wg_forward, [idx, w, f], d_out)
    _check(_run(P.WEIGHTEDGATHERFunction.apply, [idx, w, f], d_out, dtype), ref, dtype, "WG")


def test_weighted_gather_golden(golden_dir):
    """The reference's seeded WEIGHTEDGATHER check (tests/golden/wg_refcheck.npz)."""
    import os
    import numpy as np
    P = _ops()
    g = np.load(os.path.join(golden_dir, "wg_refcheck.npz"))
    idx = torch.from_numpy(g["idx"].astype(np.int64))
    w, f, up = torch.from_numpy(g["w"]), torch.from_numpy(g["f"]), torch.from_numpy(g["up"])
    out, (d_w, d_f) = _run(P.WEIGHTEDGATHERFunction.apply, [idx, w, f], torch.full_like(up, 1.0 / up.numel()), torch.float32)
    assert rel_err(out, up) <= 1e-5
    assert rel_err(d_w, torch.from_numpy(g["d_w"])) <= 1e-5
    assert rel_err(d_f, torch.from_numpy(g["d_f"])) <= 1e-5


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
@pytest.mark.parametrize("B,N,Nk,M,K,C", [(100, 50, 100, 8, 4, 32), (2, 300, 1200, 12, 4, 256), (1, 7, 5, 3, 2, 20)])
def test_msdetrpc(B, N, Nk, M, K, C, dtype):
    """MSDETRPCFunction against the torch formulation of the reference's own check (clusten/test_msdetrpc_kernel.py:14-42,
    first case = its sizes, seeded): forward and the three gradients."""
    P = _ops()
    g = torch.Generator().manual_seed(B + N)
    cast = lambda t: t.to(dtype).float()
    idx = torch.randint(Nk, (B, N, M, K), generator=g)
    w, attn, val = cast(torch.rand(B, N, M, K, generator=g)), cast(torch.rand(B, N, M, generator=g)), cast(torch.rand(B, Nk, C, generator=g))
    go = cast(torch.randn(B, N, C, generator=g))
    ref = co.fwd_bwd(co.msdetrpc_forward, [idx, w, attn, val], go)
    _check(_run(P.MSDETRPCFunction.apply, [idx, w, attn, val], go, dtype), ref, dtype, "MSDETRPC")


@pytest.mark.parametrize("amp", [False, True], ids=["fp32", "bf16-autocast"])
@pytest.mark.parametrize("shape,cin,cout", [((2, 500, 64), 64, 192), ((3, 77, 96), 96, 96), ((1000, 128), 128, 1536), ((5, 40), 40, 7)])
def test_linear_function_matches_torch(shape, cin, cout, amp):
    """ops.linear (F.linear with the column-sum bias gradient) against nn.functional.linear + autograd, eager and under bf16
    autocast: output, grad_input, grad_weight, grad_bias (dtypes included)."""
    from autofocusformermod_b200 import ops
    g = torch.Generator().manual_seed(cin + cout)
    x = torch.randn(*shape, generator=g).cuda()
    w = (torch.randn(cout, cin, generator=g) * cin ** -0.5).cuda()
    b = torch.randn(cout, generator=g).cuda()
    go = torch.randn(*shape[:-1], cout, generator=g).cuda()
    res = []
    for fn in (torch.nn.functional.linear, ops.linear):
        xi, wi, bi = (t.clone().requires_grad_(True) for t in (x, w, b))
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
            y = fn(xi, wi, bi)
        y.backward(go.to(y.dtype))
        res.append((y.detach(), xi.grad, wi.grad, bi.grad))
    tol = 2e-2 if amp else 1e-5
    for a, r in zip(res[1], res[0]):
        assert a.dtype == r.dtype and a.shape == r.shape
        assert rel_err(a.float().cpu(), r.float().cpu()) <= tol


def test_dtype_cast_rules_and_none_grads():
    """clusten.py:27-28,54-55,80-81,106-107: second operand follows the first; idx gets no gradient."""
    P = _ops()
    q = torch.randn(1, 2, 16, 32, device="cuda", dtype=torch.bfloat16)
    k = torch.randn(1, 2, 16, 32, device="cuda", dtype=torch.float32)
    idx = torch.randint(0, 16, (1, 16, 8), device="cuda")
    assert P.CLUSTENQKFunction.apply(q, k, idx).dtype == torch.bfloat16
    a = torch.rand(1, 2, 16, 8, device="cuda")
    assert P.CLUSTENAVFunction.apply(a, k.bfloat16(), idx).dtype == torch.float32
    w = torch.rand(1, 16, 8, device="cuda", dtype=torch.float32)
    f = torch.rand(1, 16, 24, device="cuda", dtype=torch.bfloat16)
    assert P.WEIGHTEDGATHERFunction.apply(idx, w, f).dtype == torch.bfloat16
    w4 = torch.rand(1, 16, 8, 4, device="cuda", dtype=torch.bfloat16)
    assert P.CLUSTENWFFunction.apply(w4, f.float(), idx).dtype == torch.bfloat16
    with pytest.raises(RuntimeError):
        P.CLUSTENQKFunction.apply(q.double(), k.double(), idx)            # fp64 is not supported (documented)


def test_empty_inputs():
    P = _ops()
    q = torch.randn(0, 2, 16, 32, device="cuda")
    idx = torch.zeros(0, 16, 8, dtype=torch.int64, device="cuda")
    assert P.CLUSTENQKFunction.apply(q, q, idx).shape == (0, 2, 16, 8)


@pytest.mark.parametrize("shape", ["small_stage0", "base_stage0"])
def test_backbone_scale_properties(shape):
    """Full BASELINE-size shapes, checked through size-independent properties (the CPU oracle would take minutes):
    linearity in each operand, and the adjoint identity <QK(q,k), g> == <q, d_q> == <k, d_k> evaluated in fp64."""
    P = _ops()
    if shape == "small_stage0":
        B, H, N, C, M, m = 8, 3, 16384, 32, 48, 8
    else:
        B, H, N, C, M, m = 2, 4, 32768, 32, 144, 24
    g = torch.Generator(device="cuda").manual_seed(0)
    q = torch.randn(B, H, N, C, device="cuda", generator=g, requires_grad=True)
    k = torch.randn(B, H, N, C, device="cuda", generator=g, requires_grad=True)
    # run-structured neighbourhoods: M/m runs of m consecutive rows
    runs = torch.randint(0, N // m, (B, N, M // m), device="cuda", generator=g)
    idx = (runs.unsqueeze(-1) * m + torch.arange(m, device="cuda")).reshape(B, N, M)
    go = torch.randn(B, H, N, M, device="cuda", generator=g)
    attn = P.CLUSTENQKFunction.apply(q, k, idx)
    attn2 = P.CLUSTENQKFunction.apply(2.0 * q.detach(), k.detach(), idx)
    assert rel_err(attn2, 2.0 * attn) <= 1e-6
    attn.backward(go)
    lhs = float((attn.detach().double() * go.double()).sum())
    for t in (q, k):
        rhs = float((t.detach().double() * t.grad.double()).sum())
        assert abs(lhs - rhs) <= 1e-5 * abs(lhs) + 1e-3
    # spot-check 64 random tokens against direct dot products
    bi = torch.randint(0, B, (64,), device="cuda", generator=g)
    ti = torch.randint(0, N, (64,), device="cuda", generator=g)
    rows = k.detach()[bi[:, None, None], torch.arange(H, device="cuda")[None, :, None], idx[bi, ti][:, None, :]]   # 64 H M C
    ref = (q.detach()[bi, :, ti].unsqueeze(2) * rows).sum(-1)
    assert rel_err(attn.detach()[bi, :, ti], ref) <= 1e-5
    # AV adjoint identity
    a = torch.rand(B, H, N, M, device="cuda", generator=g, requires_grad=True)
    v = torch.randn(B, H, N, C, device="cuda", generator=g, requires_grad=True)
    gf = torch.randn(B, H, N, C, device="cuda", generator=g)
    feat = P.CLUSTENAVFunction.apply(a, v, idx)
    feat.backward(gf)
    lhs = float((feat.detach().double() * gf.double()).sum())
    for t in (a, v):
        rhs = float((t.detach().double() * t.grad.double()).sum())
        assert abs(lhs - rhs) <= 1e-5 * abs(lhs) + 1e-3


def test_tile_path_is_taken_for_clustered_idx_and_refused_for_random_idx():
    """The device-side dispatch flag of the tile pack: curve-ordered octet neighbourhoods run on the tensor-core
    kernels (flag 0, union <= U_MAX, no impure slot); a random index tensor falls back to the generic kernels."""
    from autofocusformermod_b200 import ops
    c = inputs.qkv_case(B=2, H=2, N=4096, C=32, M=48, seed=0, structured=True)
    generic, max_u, impure, over = ops.pack_flags(c["idx"].cuda(), 4096)[:4]
    assert generic == 0 and impure == 0 and over == 0 and 6 <= max_u <= 48
    r = inputs.random_neighbourhood(2, 512, 512, 48, seed=1).cuda()
    assert ops.pack_flags(r, 512)[0] == 1


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16], ids=["f32", "bf16", "f16"])
@pytest.mark.parametrize("H,C", [(2, 16), (3, 32), (16, 24), (4, 8)])
@pytest.mark.parametrize("n,m,nbhd", [(1024, 8, 48), (2000, 8, 48), (1536, 24, 144), (2003, 8, 48), (655, 8, 48), (1540, 24, 144)])
def test_tile_kernels_match_generic_and_oracle(n, m, nbhd, H, C, dtype):
    """Tensor-core tile path vs the generic row-gather path vs the oracle on the model's own neighbourhood structure
    (m = 8 / M = 48 and AFF-Base's m = 24 / M = 144), per-head dims 8/16/24/32, ragged last tile (n = 2000)."""
    from autofocusformermod_b200 import ops
    P = _ops()
    B = 2
    _, idx, _, _ = inputs.structured_neighbourhood(B, n, 64, 64, m, nbhd, seed=n)
    M = idx.shape[-1]
    g = torch.Generator().manual_seed(n + C)
    rnd = lambda *s: torch.randn(*s, generator=g).to(dtype).float()
    q, k, v = rnd(B, H, n, C), rnd(B, H, n, C), rnd(B, H, n, C)
    attn, d_attn, d_feat = rnd(B, H, n, M).softmax(-1).to(dtype).float(), rnd(B, H, n, M), rnd(B, H, n, C)
    flags = ops.pack_flags(idx.cuda(), n)
    assert flags[0] == 0, flags                                   # tensor-core path, also with a padded last cluster ...
    assert (flags[2] > 0) == (n % m != 0), flags                  # ... whose tokens take the slow in-kernel path
    tile_qk = _run(P.CLUSTENQKFunction.apply, [q, k, idx], d_attn, dtype)
    tile_av = _run(P.CLUSTENAVFunction.apply, [attn, v, idx], d_feat, dtype)
    ops.USE_TILE_KERNELS = False
    try:
        gen_qk = _run(P.CLUSTENQKFunction.apply, [q, k, idx], d_attn, dtype)
        gen_av = _run(P.CLUSTENAVFunction.apply, [attn, v, idx], d_feat, dtype)
    finally:
        ops.USE_TILE_KERNELS = True
    ref_qk = co.fwd_bwd(co.qk_forward, [q, k, idx], d_attn)
    ref_av = co.fwd_bwd(co.av_forward, [attn, v, idx], d_feat)
    _check(tile_qk, ref_qk, dtype, "QK tile")
    _check(tile_av, ref_av, dtype, "AV tile")
    _check(gen_qk, ref_qk, dtype, "QK generic")
    _check(gen_av, ref_av, dtype, "AV generic")


def _fused_reference(q, k, v, idx, bias_tab, bias_idx, mask, blank_k, blank_v):
    """aff.py:114-155 with the oracle ops (CPU, fp32): returns (out [B,N,H*C], probs [B,H,N,M+1])."""
    B, H, N, C = q.shape
    attn = co.qk_forward(q, k, idx)
    attn = attn + bias_tab[bias_idx.long()].permute(0, 3, 1, 2)
    if mask is not None:
        attn = attn + (1 - mask.long().reshape(B, 1, N, -1)) * (-100)
    blank = (q * blank_k.reshape(1, H, 1, C)).sum(-1, keepdim=True)
    p = torch.cat([attn, blank], dim=-1).softmax(-1)
    out = co.av_forward(p[..., :-1], v, idx) + p[..., -1:] * blank_v.reshape(1, H, 1, C)
    return out.permute(0, 2, 1, 3).reshape(B, N, H * C), p


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
@pytest.mark.parametrize("H,C", [(2, 16), (3, 32), (16, 24)])
@pytest.mark.parametrize("n,m,nbhd,kind", [(1024, 8, 48, "clustered"), (2003, 8, 48, "clustered"), (1540, 24, 144, "clustered"),
                                            (300, 8, 48, "random")])
def test_fused_attention_core(n, m, nbhd, kind, H, C, dtype, attn_fwd_kernel):
    """clusten_attn_fwd (QK + bias + mask + blank + softmax + AV fused) against the op-by-op oracle composition, on the
    tensor-core tile path (clustered idx, incl. padded last clusters -> cluster mask + impure tokens, and AFF-Base's
    M = 144) and on the generic kernel (random idx)."""
    from autofocusformermod_b200 import ops
    B = 2
    if kind == "clustered":
        _, idx, mask, _ = inputs.structured_neighbourhood(B, n, 64, 64, m, nbhd, seed=n)
    else:
        idx, mask = inputs.random_neighbourhood(B, n, n, nbhd, seed=n), None
    M = idx.shape[-1]
    g = torch.Generator().manual_seed(n + H)
    rnd = lambda *s: torch.randn(*s, generator=g).to(dtype).float()
    q, k, v = rnd(B, H, n, C) * C ** -0.5, rnd(B, H, n, C), rnd(B, H, n, C)
    q = q.to(dtype).float()
    R = 37
    bias_tab = torch.randn(R, H, generator=g)
    bias_idx = torch.randint(0, R, (B, n, M), generator=g, dtype=torch.int32)
    blank_k, blank_v = rnd(H * C), rnd(H * C)
    ref_out, ref_p = _fused_reference(q, k, v, idx, bias_tab, bias_idx, mask, blank_k, blank_v)
    cu = lambda t: None if t is None else t.cuda()
    out, probs = ops.cluster_attention_fused(q.cuda().to(dtype), k.cuda().to(dtype), v.cuda().to(dtype), idx.cuda(), bias_tab.cuda(),
                                             bias_idx.cuda(), None if mask is None else mask.to(torch.uint8).cuda(),
                                             blank_k.cuda().to(dtype), blank_v.cuda().to(dtype), need_probs=True)
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    assert rel_err(out.float(), ref_out) <= tol
    assert rel_err(probs, ref_p) <= tol
    if kind == "clustered":
        assert ops.pack_flags(idx.cuda(), n)[0] == 0


@pytest.mark.parametrize("n,m,nbhd", [(2003, 8, 48), (655, 8, 48), (131, 8, 48), (1540, 24, 144)])
def test_mask_aware_pack_keeps_padded_clusters_on_the_tile_path(n, m, nbhd):
    """Padded last cluster (point_utils.py:282-283): impure tokens for the literal pack (QK / AV ops), none for the mask-aware
    pack of the fused attention path, which treats masked slots as wildcards."""
    from autofocusformermod_b200 import ops
    _, idx, mask, _ = inputs.structured_neighbourhood(2, n, 64, 64, m, nbhd, seed=n)
    idx, mask8 = idx.cuda(), mask.to(torch.uint8).cuda()
    plain, masked = ops.pack_flags(idx, n), ops.pack_flags(idx, n, mask=mask8)
    assert plain[2] > 0 and masked[2] == 0 and masked[0] == 0, (plain, masked)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
@pytest.mark.parametrize("U,CH,shape,permuted", [(700, 3, (2, 500, 48), True), (700, 3, (2, 500, 48), False), (5000, 16, (1, 300, 48), True),
                                                 (37, 4, (3, 200, 48), False), (1, 2, (1, 5, 8), False), (9000, 2, (1, 400, 48), False)])
def test_table_lookup(U, CH, shape, permuted, dtype):
    """tab[inverse] (aff.py:129-132 restricted to the referenced table rows) and its segment-sum gradient against torch
    indexing in fp32; the permuted case is the [B,H,N,M] gradient layout of the attention bias."""
    from autofocusformermod_b200 import ops
    g = torch.Generator().manual_seed(3)
    tab = torch.randn(U, CH, generator=g).to(dtype)
    inv = torch.randint(0, U, shape, generator=g)
    d_out = torch.randn(*shape, CH, generator=g).to(dtype)
    t_ref = tab.float().requires_grad_(True)
    ref = t_ref[inv]
    ref.backward(d_out.float())
    for idt in (torch.int64, torch.int32):
        t = tab.cuda().detach().requires_grad_(True)
        out = ops.table_lookup(t, inv.cuda().to(idt))
        assert out.dtype == dtype and tuple(out.shape) == (*shape, CH)
        assert torch.equal(out.detach().float().cpu(), ref.detach())
        go = d_out.cuda()
        if permuted:                                   # gradient arrives as a permuted view of [B,CH,N,M] memory
            go = go.permute(0, 3, 1, 2).contiguous().permute(0, 2, 3, 1)
        out.backward(go)
        torch.cuda.synchronize()
        assert rel_err(t.grad.float().cpu(), t_ref.grad) <= TOL[dtype]
        # table sized by an upper bound, the referenced row count known only on the device
        t2 = torch.cat([tab, torch.zeros(3 * U + 5, CH, dtype=dtype)]).cuda().detach().requires_grad_(True)
        cnt = torch.tensor([U], dtype=torch.int32, device="cuda")
        out2 = ops.table_lookup(t2, inv.cuda().to(idt), cnt)
        assert torch.equal(out2.detach(), out.detach())
        out2.backward(go)
        assert rel_err(t2.grad[:U].float().cpu(), t_ref.grad) <= TOL[dtype] and float(t2.grad[U:].abs().max()) == 0.0


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "f16"])
@pytest.mark.parametrize("H,C", [(2, 16), (3, 32), (16, 24)])
@pytest.mark.parametrize("n,m,nbhd,kind", [(1024, 8, 48, "clustered"), (2003, 8, 48, "clustered"), (1540, 24, 144, "clustered"),
                                            (300, 8, 48, "random")])
def test_fused_attention_core_backward(n, m, nbhd, kind, H, C, dtype):
    """ClusterAttentionCoreFunction (clusten_attn_fwd + clusten_attn_bwd + scatter + table grad) against autograd of the
    op-by-op oracle composition (aff.py:114-155) in fp32 on the same 16-bit-rounded inputs: out and the gradients of q,
    kv, the bias table, blank_k and blank_v."""
    from autofocusformermod_b200 import ops
    B = 2
    if kind == "clustered":
        _, idx, mask, _ = inputs.structured_neighbourhood(B, n, 64, 64, m, nbhd, seed=n)
    else:
        idx, mask = inputs.random_neighbourhood(B, n, n, nbhd, seed=n), None
    M = idx.shape[-1]
    g = torch.Generator().manual_seed(n + H)
    rnd = lambda *s: torch.randn(*s, generator=g).to(dtype).float()
    q = (rnd(B, n, H, C) * C ** -0.5).to(dtype).float()
    kv = rnd(B, n, H, 2, C)
    R = 37
    bias_tab = torch.randn(R, H, generator=g)
    bias_idx = torch.randint(0, R, (B, n, M), generator=g, dtype=torch.int32)
    blank_k, blank_v = rnd(H * C), rnd(H * C)
    d_out = rnd(B, n, H * C)
    leaves = [t.clone().requires_grad_(True) for t in (q, kv, bias_tab, blank_k, blank_v)]
    rq, rkv, rtab, rbk, rbv = leaves
    ref_out, _ = _fused_reference(rq.permute(0, 2, 1, 3), rkv[:, :, :, 0].permute(0, 2, 1, 3), rkv[:, :, :, 1].permute(0, 2, 1, 3), idx,
                                  rtab, bias_idx, mask, rbk, rbv)
    ref_out.backward(d_out)
    cq, ckv = q.cuda().to(dtype).requires_grad_(True), kv.cuda().to(dtype).requires_grad_(True)
    ctab = bias_tab.cuda().requires_grad_(True)
    cbk, cbv = blank_k.cuda().requires_grad_(True), blank_v.cuda().requires_grad_(True)
    out = ops.cluster_attention_core(cq, ckv, ctab, cbk, cbv, idx.cuda(), bias_idx.cuda(),
                                     None if mask is None else mask.to(torch.uint8).cuda())
    out.backward(d_out.cuda().to(dtype))
    torch.cuda.synchronize()
    tol = 1e-2
    assert rel_err(out.float().cpu(), ref_out.detach()) <= tol
    for name, got, ref in (("d_q", cq, rq), ("d_kv", ckv, rkv), ("d_bias_tab", ctab, rtab), ("d_blank_k", cbk, rbk), ("d_blank_v", cbv, rbv)):
        e = rel_err(got.grad.float().cpu(), ref.grad)
        assert e <= 2 * tol, f"{name} rel err {e:.3e}"


TC_LINEAR = os.environ.get("CLUSTEN_TC_LINEAR") == "1"


@pytest.mark.parametrize("R,K,N", [(1000, 32, 32), (4173, 128, 256), (300, 96, 288), (129, 768, 2304), (1, 32, 2), (16384, 64, 32)])
def test_linear_f32_tensor_core(R, K, N):
    """clusten_linear_f32 (3xTF32 split on the tensor cores) against a float64 matmul: fp32-level accuracy, ragged R and N,
    a strided input, no bias."""
    from autofocusformermod_b200 import ops
    g = torch.Generator().manual_seed(R + K + N)
    x = torch.randn(R, K, generator=g)
    w = torch.randn(N, K, generator=g) * K ** -0.5
    b = torch.randn(N, generator=g)
    ref = x.double() @ w.double().t() + b.double()
    xc, wc, bc = x.cuda(), w.cuda(), b.cuda()
    assert ops.linear_f32_supported(xc, wc, bc)
    y = ops.linear_f32(xc, wc, bc)
    assert y.shape == (R, N) and y.dtype == torch.float32
    assert rel_err(y.cpu().double(), ref) <= 2e-6
    y3 = ops.linear_f32(xc.view(1, R, K), wc, None)                      # leading dims, no bias
    assert y3.shape == (1, R, N)
    assert rel_err(y3[0].cpu().double(), ref - b.double()) <= 2e-6
    wide = torch.randn(R, K + 32, generator=g).cuda()                    # row stride K + 32: consumed in place
    ys = ops.linear_f32(wide[:, :K], wc, bc)
    assert rel_err(ys.cpu().double(), wide[:, :K].cpu().double() @ w.double().t() + b.double()) <= 2e-6


@pytest.mark.parametrize("R,K,N", [(1000, 32, 32), (4173, 128, 256), (300, 96, 288), (129, 768, 2304), (1, 32, 4), (16384, 64, 32),
                                   (40000, 32, 96), (5000, 256, 100), (777, 1536, 384), (20000, 32, 36),
                                   (16384, 256, 768), (20000, 128, 384), (40000, 96, 288), (16500, 256, 200)])
@pytest.mark.parametrize("epilogue", ["bias", "gelu", "residual"])
@pytest.mark.parametrize("split", ["tf32", "f16"])
def test_linear_tc_tcgen05(R, K, N, epilogue, split):
    """clusten_linear_tc_f32 (tcgen05.mma kind::tf32 through TMA and tensor memory, 3xTF32 split) against a float64 product for the
    three epilogues the block uses (aff.py:107-108 q * scale, aff.py:45-46 GELU(fc1), aff.py:230,236 shortcut + gamma * x): fp32-level
    accuracy; ragged R (not a multiple of the 128-row tile), N that is not a multiple of the column tile, several tiles per CTA
    (R = 40 000 -> 313 row tiles on 148 SMs), long K (48 chunks, 12 accumulator chains), strided input, no bias / no gamma."""
    from autofocusformermod_b200 import ops
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(R + K + N)
    x = torch.randn(R, K, generator=g)
    w = torch.randn(N, K, generator=g) * K ** -0.5
    b = torch.randn(N, generator=g)
    res = torch.randn(R, N, generator=g)
    gam = torch.rand(N, generator=g) + 0.5
    ncol = (N // 2) // 4 * 4

    def reference(xd, bias, gamma):
        y = xd.double() @ w.double().t()
        if bias is not None:
            y = y + bias.double()
        if epilogue == "gelu":
            return F.gelu(y)
        if epilogue == "residual":
            return res.double() + (y if gamma is None else gamma.double() * y)
        y[:, :ncol] *= 0.3
        return y

    xc, wc, bc, rc, gc = x.cuda(), w.cuda(), b.cuda(), res.cuda(), gam.cuda()
    assert ops.linear_tc_supported(xc, wc, bc, rc, gc)
    kw = dict(res=rc, gamma=gc, alpha=0.3, alpha_cols=ncol, split=split)
    y = ops.linear_tc(xc, wc, bc, epilogue, **kw)
    assert y.shape == (R, N) and y.dtype == torch.float32
    assert rel_err(y.cpu().double(), reference(x, b, gam)) <= 2e-6
    y1 = ops.linear_tc(xc, wc, bc, epilogue, chain=1, **kw)              # every K chunk summed in fp32 registers
    assert rel_err(y1.cpu().double(), reference(x, b, gam)) <= 1e-6
    y3 = ops.linear_tc(xc.view(1, R, K), wc, None, epilogue, res=rc.view(1, R, N), gamma=None, alpha=0.3, alpha_cols=ncol, split=split)
    assert y3.shape == (1, R, N)                                         # leading dims, no bias, no gamma
    assert rel_err(y3[0].cpu().double(), reference(x, None, None)) <= 2e-6
    wide = torch.randn(R, K + 32, generator=g).cuda()                    # row stride K + 32: consumed in place
    ys = ops.linear_tc(wide[:, :K], wc, bc, epilogue, **kw)
    assert rel_err(ys.cpu().double(), reference(wide[:, :K].cpu(), b, gam)) <= 2e-6


@pytest.mark.parametrize("R,K,N", [(1000, 32, 96), (4173, 128, 256), (300, 96, 288), (2049, 1024, 128)])
@pytest.mark.parametrize("epilogue", ["bias", "gelu"])
@pytest.mark.parametrize("split", ["tf32", "f16", None])
def test_linear_tc_layer_norm_in_gemm(R, K, N, epilogue, split):
    """``linear_tc(..., ln=...)``: the LayerNorm in front of the layer (aff.py:196-199 norm1 / norm2, aff.py:258 merge norm) applied to
    the rows while the GEMM stages them, from the statistics of the LayerNorm kernel itself (clusten_layer_norm_fwd with y = NULL).
    Equal to LayerNorm-then-linear_tc BIT FOR BIT (same statistics, same rounding of the normalised value), and to float64 at 2e-6."""
    from autofocusformermod_b200 import ops
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(R + K + N)
    x = (torch.randn(R, K, generator=g) * 1.7 + 0.4).cuda()
    w = (torch.randn(N, K, generator=g) * K ** -0.5).cuda()
    b = torch.randn(N, generator=g).cuda()
    lw, lb = (torch.rand(K, generator=g) + 0.5).cuda(), torch.randn(K, generator=g).cuda()
    mean, rstd = ops.layer_norm_stats(x, lw, lb, 1e-5)
    xd = x.double()
    m64, v64 = xd.mean(1), xd.var(1, unbiased=False)
    assert rel_err(mean.double(), m64) <= 1e-6 and rel_err(rstd.double(), (v64 + 1e-5).rsqrt()) <= 1e-6
    y = ops.linear_tc(x, w, b, epilogue, ln=(mean, rstd, lw, lb), split=split, ln_fold=False)     # split None: the default, fp16 for normalised rows
    y_two = ops.linear_tc(ops.layer_norm(x, lw, lb, 1e-5), w, b, epilogue, split=split or "f16")
    assert torch.equal(y, y_two)
    ref = F.layer_norm(xd, (K,), lw.double(), lb.double(), 1e-5) @ w.double().t() + b.double()
    if epilogue == "gelu":
        ref = F.gelu(ref)
    assert rel_err(y.double(), ref) <= 2e-6
    # default: gamma / beta folded into the layer (W * gamma, b + W beta), the kernel only normalises -- same result to fp32 rounding
    y_f = ops.linear_tc(x, w, b, epilogue, ln=(mean, rstd, lw, lb), split=split)
    assert rel_err(y_f.double(), ref) <= 2e-6 and rel_err(y_f, y) <= 2e-6
    wf, bf = ops.ln_folded_layer(w, b, lw, lb)
    assert ops.ln_folded_layer(w, b, lw, lb)[0] is wf                   # cached ...
    lw.mul_(1.5)
    wf2, _ = ops.ln_folded_layer(w, b, lw, lb)                           # ... until a parameter changes
    assert wf2 is not wf and rel_err(wf2, 1.5 * wf) <= 1e-6
    y_n = ops.linear_tc(x, w, None, epilogue, ln=(mean, rstd, lw, lb), split=split)      # no bias: b' = W beta alone
    ref_n = F.layer_norm(xd, (K,), lw.double(), lb.double(), 1e-5) @ w.double().t()
    assert rel_err(y_n.double(), F.gelu(ref_n) if epilogue == "gelu" else ref_n) <= 2e-6


@pytest.mark.parametrize("R,K,N", [(16384, 256, 768), (20000, 128, 384), (40000, 96, 288), (19000, 64, 160), (16500, 256, 200)])
@pytest.mark.parametrize("epilogue", ["bias", "gelu", "residual"])
@pytest.mark.parametrize("split", ["tf32", "f16"])
def test_linear_tc_resident_rows_equal_streamed(R, K, N, epilogue, split, monkeypatch):
    """The resident tile walk (a CTA owns row tiles, splits their rows once into tensor memory and walks the column tiles; taken when
    K fits the A stages -- 8 chunks fp16, 4 chunks TF32 -- and the row tiles fill the SMs) performs the same MMAs in the same order
    as the streamed walk: BIT-identical results, incl. more row tiles than SMs, K chunk counts that wrap the stage ring mid-tile
    (K = 96), ragged R and ragged N, and the LayerNorm-on-load form."""
    from autofocusformermod_b200 import ops
    g = torch.Generator().manual_seed(R + K + N)
    x = (torch.randn(R, K, generator=g) * 1.3).cuda()
    w = (torch.randn(N, K, generator=g) * K ** -0.5).cuda()
    b = torch.randn(N, generator=g).cuda()
    res = torch.randn(R, N, generator=g).cuda()
    gam = (torch.rand(N, generator=g) + 0.5).cuda()
    lw, lb = (torch.rand(K, generator=g) + 0.5).cuda(), torch.randn(K, generator=g).cuda()
    kw = dict(res=res, gamma=gam, alpha=0.3, alpha_cols=(N // 2) // 4 * 4, split=split)
    stats = ops.layer_norm_stats(x, lw, lb, 1e-5)
    out = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("CLUSTEN_TC_RESIDENT", mode)
        out[mode] = (ops.linear_tc(x, w, b, epilogue, **kw), ops.linear_tc(x, w, b, epilogue, ln=stats + (lw, lb), **kw))
    assert torch.equal(out["1"][0], out["0"][0]) and torch.equal(out["1"][1], out["0"][1])
    ref = x.double() @ w.double().t() + b.double()
    if epilogue == "gelu":
        ref = torch.nn.functional.gelu(ref)
    elif epilogue == "residual":
        ref = res.double() + gam.double() * ref
    else:
        ref[:, :kw["alpha_cols"]] *= 0.3
    assert rel_err(out["1"][0].double(), ref) <= 2e-6


def test_linear_tc_f16_split_scales():
    """The fp16 form of the split (kind::f16, half the MMAs): weight rows of very different magnitude each get their own power-of-two
    scale, undone per output column in the epilogue; activations up to the fp16 range are exact to 22 bits, beyond it they saturate
    instead of turning into inf / NaN."""
    from autofocusformermod_b200 import ops
    g = torch.Generator().manual_seed(11)
    R, K, N = 3000, 256, 160
    x = (torch.randn(R, K, generator=g) * 40.0).cuda()
    w = torch.randn(N, K, generator=g) * K ** -0.5
    w = (w * torch.logspace(-7, 3, N).unsqueeze(1)).cuda()               # output rows from 1e-7 to 1e3
    b = torch.zeros(N).cuda()
    y = ops.linear_tc(x, w, b, split="f16")
    ref = x.double() @ w.double().t()
    col_err = (y.double() - ref).abs().amax(0) / ref.abs().amax(0)       # every output column on its own scale
    assert float(col_err.max()) <= 2e-6, float(col_err.max())
    hi, lo, inv = ops.f16_split(w)
    assert hi.dtype == torch.float16 and inv.shape == (N,)
    scaled = (hi.float() + lo.float()) * inv.unsqueeze(1)
    assert float(((scaled - w).abs().amax(1) / w.abs().amax(1)).max()) <= 2.0 ** -21
    big = x.clone()
    big[0, 0] = 3.0e5                                                    # beyond fp16: saturates at 65504, finite result
    yb = ops.linear_tc(big, w, b, split="f16")
    assert bool(torch.isfinite(yb).all())
    assert rel_err(yb[1:], y[1:]) == 0.0                                 # the other rows are untouched
    assert rel_err(ops.linear_tc(big, w, b, split="tf32").double()[0], (big.double() @ w.double().t())[0]) <= 2e-6   # TF32 form: full range


def test_linear_tc_weight_split_follows_the_weight():
    """The (hi, lo) operands are cached per weight and rebuilt when the weight is updated in place."""
    from autofocusformermod_b200 import ops
    g = torch.Generator().manual_seed(5)
    x, w = torch.randn(300, 64, generator=g).cuda(), torch.randn(32, 64, generator=g).cuda()
    y0 = ops.linear_tc(x, w)
    hi, lo = ops.tf32_split(w)
    assert torch.equal((hi.view(torch.int32) & 0x1fff), torch.zeros_like(hi, dtype=torch.int32))       # TF32: 13 low bits clear
    assert float((hi + lo - w).abs().max()) <= float(w.abs().max()) * 2.0 ** -21
    assert ops.tf32_split(w)[0] is hi                                    # cached
    w.mul_(2.0)
    y1 = ops.linear_tc(x, w)
    assert ops.tf32_split(w)[0] is not hi
    assert rel_err(y1, 2.0 * y0) <= 1e-6


INKERNEL_BIAS = os.environ.get("CLUSTEN_INKERNEL_BIAS") == "1"


def _pos_bias_case(n, m, nbhd, kind, H, B=2, hw=64):
    """Positions, neighbourhoods, mask and the table formulation of the bias (aff.py:481-485 + 17-31 + 129) for one stage."""
    from autofocusformermod_b200.aff import rel_pos_features
    pos, idx, mask, _ = inputs.structured_neighbourhood(B, n, hw, hw, m, nbhd, seed=n)
    if kind == "random":
        idx, mask = inputs.random_neighbourhood(B, n, n, nbhd, seed=n), None
    M = idx.shape[-1]
    rel = (pos.gather(1, idx.reshape(B, -1, 1).expand(-1, -1, 2)).reshape(B, n, M, 2) - (pos.unsqueeze(2) - 511)).clamp(0, 1022).long()
    pe_idx = rel[..., 1] * 1023 + rel[..., 0]
    uniq, inverse = torch.unique(pe_idx.reshape(-1), return_inverse=True)
    g = torch.Generator().manual_seed(n + 3 * H)
    W, bvec = torch.randn(H, 5, generator=g) * 0.2, torch.randn(H, generator=g)
    return pos.float(), idx, mask, rel_pos_features(uniq), inverse.reshape(B, n, M).to(torch.int32), W, bvec


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
@pytest.mark.parametrize("H,C", [(2, 16), (3, 32)])
@pytest.mark.parametrize("n,m,nbhd,kind", [(1024, 8, 48, "clustered"), (2003, 8, 48, "clustered"), (1540, 24, 144, "clustered"),
                                            (300, 8, 48, "random")])
def test_inkernel_bias_forward_matches_table_variant(n, m, nbhd, kind, H, C, dtype):
    """clusten_attn_pos_fwd (bias computed from positions) against clusten_attn_fwd (bias gathered from pos_embed's table)."""
    from autofocusformermod_b200 import ops
    B = 2
    pos, idx, mask, feats, bias_idx, W, bvec = _pos_bias_case(n, m, nbhd, kind, H, B)
    g = torch.Generator().manual_seed(n + H)
    rnd = lambda *s: torch.randn(*s, generator=g).to(dtype).cuda()
    q, k, v = rnd(B, H, n, C) * C ** -0.5, rnd(B, H, n, C), rnd(B, H, n, C)
    bk, bv = rnd(H * C), rnd(H * C)
    m8 = None if mask is None else mask.to(torch.uint8).cuda()
    tab = (feats @ W.t() + bvec).cuda()
    ref = ops.cluster_attention_fused(q, k, v, idx.cuda(), tab, bias_idx.cuda(), m8, bk, bv)
    out = ops.cluster_attention_fused_pos(q, k, v, idx.cuda(), pos.cuda(), W.cuda(), bvec.cuda(), m8, bk, bv)
    torch.cuda.synchronize()
    assert rel_err(out.float(), ref.float()) <= (2e-5 if dtype == torch.float32 else 1e-2)


@pytest.mark.parametrize("H,C", [(2, 32), (4, 16)])
@pytest.mark.parametrize("n,m,nbhd,kind", [(1024, 8, 48, "clustered"), (2003, 8, 48, "clustered"), (1540, 24, 144, "clustered"),
                                            (300, 8, 48, "random")])
def test_inkernel_bias_training_matches_table_variant(n, m, nbhd, kind, H, C):
    """ClusterAttentionPosFunction against ClusterAttentionCoreFunction behind table_linear: output and every gradient
    (q, kv, pos_embed weight / bias, blank tokens)."""
    from autofocusformermod_b200 import ops
    B, dt = 2, torch.bfloat16
    pos, idx, mask, feats, bias_idx, W, bvec = _pos_bias_case(n, m, nbhd, kind, H, B)
    g = torch.Generator().manual_seed(n + H)
    rnd = lambda *s: torch.randn(*s, generator=g)
    q0, kv0 = (rnd(B, n, H, C) * C ** -0.5).to(dt), rnd(B, n, H, 2, C).to(dt)
    bk0, bv0, go = rnd(H * C), rnd(H * C), rnd(B, n, H * C).to(dt).cuda()
    m8 = None if mask is None else mask.to(torch.uint8).cuda()
    res = []
    for variant in ("table", "pos"):
        leaf = lambda t: t.detach().clone().cuda().requires_grad_(True)
        q, kv, w, b_, bk, bv = leaf(q0), leaf(kv0), leaf(W), leaf(bvec), leaf(bk0), leaf(bv0)
        if variant == "table":
            out = ops.cluster_attention_core(q, kv, ops.table_linear(feats.cuda(), w, b_), bk, bv, idx.cuda(), bias_idx.cuda(), m8)
        else:
            out = ops.cluster_attention_core_pos(q, kv, w, b_, bk, bv, idx.cuda(), pos.cuda(), m8)
        out.backward(go)
        res.append([out] + [t.grad for t in (q, kv, w, b_, bk, bv)])
    torch.cuda.synchronize()
    for name, a, r in zip(("out", "d_q", "d_kv", "d_pe_weight", "d_pe_bias", "d_blank_k", "d_blank_v"), res[1], res[0]):
        assert rel_err(a.float(), r.float()) <= 2e-2, name


@pytest.mark.parametrize("xdt,ydt", [(torch.float32, torch.float32), (torch.bfloat16, torch.float32), (torch.bfloat16, torch.bfloat16)],
                         ids=["f32", "bf16-f32", "bf16-bf16"])
@pytest.mark.parametrize("R,C", [(1000, 32), (513, 96), (64, 384), (7, 1024), (300, 100), (5, 8), (65025, 4)])
def test_layer_norm(R, C, xdt, ydt):
    """clusten_layer_norm_fwd / _bwd (one warp per token row) against torch.nn.functional.layer_norm in fp32."""
    from autofocusformermod_b200 import ops
    g = torch.Generator().manual_seed(R + C)
    x = (torch.randn(R, C, generator=g) * 2 + 0.5).to(xdt)
    w, b = torch.randn(C, generator=g), torch.randn(C, generator=g)
    dy = torch.randn(R, C, generator=g).to(ydt)
    xr, wr, br = x.float().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xr, (C,), wr, br, 1e-5)
    ref.backward(dy.float())
    xc, wc, bc = (t.detach().cuda().requires_grad_(True) for t in (x, w, b))
    y = ops.layer_norm(xc, wc, bc, 1e-5, ydt)
    assert y.dtype == ydt
    y.backward(dy.cuda())
    torch.cuda.synchronize()
    tol = 1e-5 if (xdt, ydt) == (torch.float32, torch.float32) else 1e-2
    assert rel_err(y.float().cpu(), ref.detach()) <= tol
    assert rel_err(xc.grad.float().cpu(), xr.grad) <= tol
    assert rel_err(wc.grad.cpu(), wr.grad) <= max(tol, 2e-5)
    assert rel_err(bc.grad.cpu(), br.grad) <= max(tol, 2e-5)


@pytest.mark.parametrize("use_gamma,use_scale", [(True, True), (True, False), (False, True)], ids=["gamma+drop", "gamma", "drop"])
@pytest.mark.parametrize("rdt,xdt", [(torch.float32, torch.float32), (torch.float32, torch.bfloat16), (torch.bfloat16, torch.bfloat16),
                                     (torch.float16, torch.float16)], ids=["f32", "f32+bf16", "bf16", "f16"])
@pytest.mark.parametrize("B,rows,C", [(3, 77, 96), (2, 1000, 32), (1, 5, 1024), (5, 333, 64)])
def test_scale_residual(B, rows, C, rdt, xdt, use_gamma, use_scale):
    """clusten_scale_residual_fwd / _bwd (res + x * gamma * sample_scale, aff.py:230,236 with timm's DropPath folded in) against
    the op-by-op torch formulation: bit-exact in fp32, 16-bit within 1e-2; gradients of res, x and gamma."""
    from autofocusformermod_b200 import ops
    g = torch.Generator().manual_seed(B * 1000 + rows + C)
    res = torch.randn(B, rows, C, generator=g).to(rdt)
    x = torch.randn(B, rows, C, generator=g).to(xdt)
    gamma = (torch.randn(C, generator=g) * 0.5) if use_gamma else None
    scale = (torch.bernoulli(torch.full((B,), 0.6), generator=g) / 0.6) if use_scale else None
    if scale is not None and B > 1:
        scale[0], scale[-1] = 0.0, 1 / 0.6                      # both outcomes present
    up = torch.randn(B, rows, C, generator=g)

    def leaf(t):
        return None if t is None else t.detach().cuda().requires_grad_(t.is_floating_point())

    rc, xc, gc = leaf(res), leaf(x), leaf(gamma)
    sc = None if scale is None else scale.cuda()
    assert ops.scale_residual_supported(rc, xc, gc, sc)
    out = ops.scale_residual(rc, xc, gc, sc)
    want_dtype = torch.float32 if (use_gamma or rdt == torch.float32) else rdt
    assert out.dtype == want_dtype and out.shape == res.shape
    out.backward(up.cuda().to(out.dtype))
    # reference: the same formula in fp32 on the (rounded) inputs
    rr, xr = res.float().requires_grad_(True), x.float().requires_grad_(True)
    gr = None if gamma is None else gamma.clone().requires_grad_(True)
    t = xr if gr is None else gr * xr
    if scale is not None:
        t = t * scale.view(B, 1, 1)
    ref = rr + t
    ref.backward(up.to(out.dtype).float())
    torch.cuda.synchronize()
    if (rdt, xdt) == (torch.float32, torch.float32):
        assert torch.equal(out.cpu(), ref.detach()), "fp32 forward must equal the op-by-op result bit for bit"
        assert torch.equal(xc.grad.cpu(), xr.grad)
    tol = 1e-6 if (rdt, xdt) == (torch.float32, torch.float32) else 1e-2
    assert rel_err(out.float().cpu(), ref.detach()) <= tol
    assert rel_err(rc.grad.float().cpu(), rr.grad) <= tol
    assert rel_err(xc.grad.float().cpu(), xr.grad) <= tol
    if gr is not None:
        assert gc.grad.dtype == torch.float32
        assert rel_err(gc.grad.cpu(), gr.grad) <= max(tol, 2e-5)


@pytest.mark.parametrize("use_count", [False, True], ids=["all-rows", "device-count"])
@pytest.mark.parametrize("R,F,H", [(1000, 5, 2), (65025, 5, 16), (333, 5, 32), (77, 8, 3), (1, 5, 4)])
def test_table_linear(R, F, H, use_count):
    """clusten_table_linear_fwd / _bwd (pos_embed = Linear(5, heads) on the referenced table rows, aff.py:101,129) against
    F.linear in fp64 on the first ``count`` rows; rows past the count are zeros and take no part in the gradients."""
    from autofocusformermod_b200 import ops
    g = torch.Generator().manual_seed(R + 7 * H)
    feat = torch.randn(R, F, generator=g) * 3
    w, b = torch.randn(H, F, generator=g), torch.randn(H, generator=g)
    up = torch.randn(R, H, generator=g)
    U = max(1, (R * 2) // 3) if use_count else R
    count = torch.tensor(U, dtype=torch.int32).cuda() if use_count else None
    wc, bc = w.clone().cuda().requires_grad_(True), b.clone().cuda().requires_grad_(True)
    assert ops.table_linear_supported(feat.cuda(), wc, bc)
    out = ops.table_linear(feat.cuda(), wc, bc, count)
    out.backward(up.cuda())
    torch.cuda.synchronize()
    wr, br = w.double().requires_grad_(True), b.double().requires_grad_(True)
    ref = torch.nn.functional.linear(feat[:U].double(), wr, br)
    ref.backward(up[:U].double())
    assert out.dtype == torch.float32 and out.shape == (R, H)
    assert rel_err(out[:U].cpu().double(), ref.detach()) <= 1e-6
    assert not out[U:].any()
    assert rel_err(wc.grad.cpu().double(), wr.grad) <= 2e-5
    assert rel_err(bc.grad.cpu().double(), br.grad) <= 2e-5
    # no bias
    out2 = ops.table_linear(feat.cuda(), wc.detach(), None, count)
    assert rel_err(out2[:U].cpu().double(), torch.nn.functional.linear(feat[:U].double(), w.double())) <= 1e-6


def test_scale_residual_falls_back_for_shapes_the_kernel_does_not_take():
    """Non-contiguous operands and odd channel counts go through the torch formulation with the same result."""
    from autofocusformermod_b200 import ops
    g = torch.Generator().manual_seed(5)
    res = torch.randn(2, 50, 30, generator=g).cuda()
    x = torch.randn(2, 50, 30, generator=g).cuda()
    gamma = torch.randn(30, generator=g).cuda()
    scale = torch.tensor([0.0, 2.0]).cuda()
    assert not ops.scale_residual_supported(res, x, gamma, scale)               # C % 4 != 0
    out = ops.scale_residual(res, x, gamma, scale)
    assert torch.equal(out, res + (gamma * x) * scale.view(2, 1, 1))
    res2 = torch.randn(2, 64, 50, generator=g).cuda().transpose(1, 2)           # [2, 50, 64] non-contiguous
    x2 = torch.randn(2, 50, 64, generator=g).cuda()
    assert not ops.scale_residual_supported(res2, x2, None, scale)
    assert torch.equal(ops.scale_residual(res2, x2, None, scale), res2 + x2 * scale.view(2, 1, 1))
    assert torch.equal(ops.scale_residual(res2, x2), res2 + x2)


@pytest.mark.parametrize("bitmap", ["1", "0"])
@pytest.mark.parametrize("n,m,nbhd,hw,B", [(4096, 8, 48, 64, 2), (2003, 8, 48, 64, 2), (1540, 24, 144, 64, 2), (655, 8, 48, 128, 2),
                                          (40000, 8, 48, 256, 4), (30001, 8, 48, 512, 3), (1540, 6, 36, 64, 2)])
def test_stage_prepare_matches_torch_formulation(n, m, nbhd, hw, B, bitmap, monkeypatch):
    """clusten_stage_prepare against the op-by-op torch formulation of aff.py:475-485 + torch.unique (bit-exact integers); both
    marking schemes (shared-memory bitmap per CTA = default, global byte map), incl. sizes where every CTA loops and m % 4 != 0."""
    import math
    from autofocusformermod_b200 import point_utils as pu
    monkeypatch.setenv("CLUSTEN_PREPARE_BITMAP", bitmap)
    g = torch.Generator().manual_seed(n)
    pos = torch.stack([torch.randperm(hw * hw, generator=g)[:n] for _ in range(B)])
    pos = torch.stack([pos % hw, pos // hw], dim=-1).float().cuda()
    pos, mean_pos, member, cmask, _ = pu.space_filling_cluster(pos, m, hw, hw)
    k = member.shape[1]
    nnc = min(int(round(nbhd / float(m))), k)
    nearest = pu.knn_keops(pos, mean_pos, nnc)
    gi = nearest.view(B, -1, 1).expand(-1, -1, m)
    ref_member = member.gather(index=gi, dim=1).reshape(B, n, nnc * m)
    ref_mask = None if cmask is None else cmask.gather(index=gi, dim=1).reshape(B, n, nnc * m)
    pos_nb = pos.gather(index=ref_member.view(B, -1, 1).expand(-1, -1, 2), dim=1).reshape(B, n, nnc * m, 2)
    rel = (pos_nb - (pos.unsqueeze(2) - 511)).clamp(0, 1022)
    pe = (rel[..., 1] * 1023 + rel[..., 0]).long()
    ref_uniq, ref_inv = torch.unique(pe.reshape(-1), return_inverse=True)
    member_idx, mask64, mask8, uniq, bias_idx = pu.stage_prepare(pos, nearest, member, cmask)
    assert torch.equal(member_idx, ref_member)
    if cmask is None:
        assert mask64 is None and mask8 is None
    else:
        assert torch.equal(mask64, ref_mask) and torch.equal(mask8.long(), ref_mask)
    assert torch.equal(uniq, ref_uniq)
    assert torch.equal(bias_idx.reshape(-1).long(), ref_inv)
    # no-host-read variant: uniq at its upper bound (2h-1)(2w-1), the count stays on the device
    r2 = pu.stage_prepare(pos, nearest, member, cmask, extent=(hw, hw))
    U = int(r2[5].item())
    assert U == ref_uniq.numel() and r2[3].numel() == min((2 * hw - 1) ** 2, B * n * nnc * m) >= U
    assert torch.equal(r2[3][:U], ref_uniq) and bool((r2[3][U:] == 0).all()) and torch.equal(r2[4], bias_idx)


@pytest.mark.parametrize("dtype,c", [(torch.float32, 32), (torch.float32, 96), (torch.float32, 2), (torch.float32, 1), (torch.int64, 48),
                                     (torch.int64, 1), (torch.uint8, 48), (torch.uint8, 3), (torch.bfloat16, 5), (torch.float16, 384)])
@pytest.mark.parametrize("B,n,k", [(2, 1000, 1000), (3, 4099, 517), (1, 7, 20)])
def test_gather_rows_equals_torch_gather(dtype, c, B, n, k):
    """clusten_gather_rows against ``src.gather(1, idx.expand(-1, -1, c))`` (aff.py:332,335,340,471): bit-identical for every
    row width / alignment class (16-, 8-, 4-, 1-byte pieces), permutations, subsets and repeated rows."""
    from autofocusformermod_b200 import ops
    g = torch.Generator().manual_seed(1000 * c + n)
    if dtype.is_floating_point:
        src = torch.randn(B, n, c, generator=g).to(dtype).cuda()
    else:
        src = torch.randint(0, 200, (B, n, c), generator=g).to(dtype).cuda()
    if k == n:
        idx = torch.stack([torch.randperm(n, generator=g) for _ in range(B)]).unsqueeze(2).cuda()
    else:
        idx = torch.randint(0, n, (B, k, 1), generator=g).cuda()
    before = ops.launch_count()
    out = ops.gather_rows(src, idx)
    assert ops.launch_count() == before + 1, "the native row gather did not run"
    assert out.dtype == src.dtype and torch.equal(out, src.gather(1, idx.expand(-1, -1, c)))
    # a view whose rows are not contiguous is copied first; a source that needs a gradient keeps autograd (torch.gather)
    wide = torch.cat([src, src], dim=2)[:, :, :c]
    assert torch.equal(ops.gather_rows(wide, idx), out)
    if dtype == torch.float32:
        leaf = src.clone().requires_grad_(True)
        o2 = ops.gather_rows(leaf, idx)
        assert o2.requires_grad and torch.equal(o2.detach(), out)
        with torch.no_grad():
            assert not ops.gather_rows(leaf, idx).requires_grad


@pytest.mark.parametrize("oc", [16, 24, 32, 48, 64])
@pytest.mark.parametrize("B,H,W", [(2, 64, 96), (1, 37, 50), (3, 128, 128)])
def test_stem_conv_bn_gelu_matches_torch(oc, B, H, W):
    """clusten_stem_conv_bn_gelu against ``act1(bn(proj1(x)))`` of PatchEmbed.forward (aff.py:527-529,549) evaluated in float64:
    <= 1e-5 (north-star tolerance; the fp32 cuDNN / ATen chain itself is ~1e-6 from float64), odd sizes included."""
    from torch import nn
    from autofocusformermod_b200 import ops
    torch.manual_seed(oc + H)
    conv = nn.Conv2d(3, oc, 3, stride=2, padding=1).cuda()
    bn = nn.BatchNorm2d(oc).cuda()
    with torch.no_grad():
        bn.running_mean.normal_(0, 0.5)
        bn.running_var.uniform_(0.3, 2.0)
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.normal_(0, 0.3)
    bn.eval()
    x = torch.randn(B, 3, H, W, device="cuda") * 1.7
    with torch.no_grad():
        assert ops.stem_conv_bn_gelu_supported(x, conv, bn)
        before = ops.launch_count()
        y = ops.stem_conv_bn_gelu(x, conv, bn)
        assert ops.launch_count() == before + 1
        c64, b64 = conv.double(), bn.double()
        ref = torch.nn.functional.gelu(b64(c64(x.double())))
    assert y.shape == ref.shape and y.dtype == torch.float32
    assert rel_err(y, ref) <= 1e-5
    assert float((y.double() - ref).abs().max()) <= 1e-5 * max(1.0, float(ref.abs().max()))
    # not supported -> the caller keeps the four-pass formulation
    bn.train()
    assert not ops.stem_conv_bn_gelu_supported(x, conv.float(), bn.float())
    assert not ops.stem_conv_bn_gelu_supported(x.half(), conv, bn.eval())


@pytest.mark.parametrize("E", [32, 64, 96, 128])
@pytest.mark.parametrize("B,H,W", [(2, 64, 96), (1, 50, 37), (2, 128, 128)])
def test_stem_as_gemm_matches_torch(E, B, H, W, monkeypatch):
    """PatchEmbed.forward (aff.py:537-565) with conv1 + BatchNorm + GELU in one pass, the second convolution as im2col + tcgen05 GEMM
    (ops.stem_tokens) and the patch norm, against the same module on cuDNN / ATen evaluated in float64: tokens <= 1e-5, positions equal;
    the im2col rows against F.unfold (bit-exact); odd sizes (padded to a multiple of 4 first) included."""
    import torch.nn.functional as F
    from autofocusformermod_b200 import aff, ops
    torch.manual_seed(E + H)
    pe = aff.PatchEmbed(in_chans=3, embed_dim=E, norm_layer=aff.LayerNorm).cuda().eval()
    with torch.no_grad():
        pe.bn.running_mean.normal_(0, 0.5)
        pe.bn.running_var.uniform_(0.3, 2.0)
        pe.bn.weight.uniform_(0.5, 1.5)
        pe.bn.bias.normal_(0, 0.3)
        x = torch.randn(B, 3, H, W, device="cuda") * 1.3
        before = ops.launch_count()
        pos, tok, h, w = pe(x)
        assert ops.launch_count() >= before + 3, "the stem did not take the native path"
        monkeypatch.setattr(aff, "STEM_GEMM", False)
        monkeypatch.setattr(aff, "FUSED_STEM", False)
        pos_t, tok_t, h_t, w_t = pe(x)                                   # cuDNN / ATen, fp32
        pe64 = aff.PatchEmbed(in_chans=3, embed_dim=E, norm_layer=torch.nn.LayerNorm).cuda().double().eval()
        pe64.load_state_dict({k: v.double() for k, v in pe.state_dict().items()})
        _, tok64, _, _ = pe64(x.double())
    assert (h, w) == (h_t, w_t) and torch.equal(pos, pos_t) and tok.shape == tok_t.shape == tok64.shape
    assert rel_err(tok, tok64) <= 1e-5 and rel_err(tok_t, tok64) <= 1e-5
    # the im2col rows alone
    mid = torch.randn(B, 24, 30, 16, device="cuda")                      # pixel-major [B, H, W, C]
    A = ops.stem_im2col(mid, 160)
    ref = F.unfold(mid.permute(0, 3, 1, 2), kernel_size=3, stride=2, padding=1)          # [B, C * 9, L], row index c * 9 + tap
    ref = ref.view(B, 16, 9, -1).permute(0, 3, 2, 1).reshape(-1, 144)                  # [(b, py, px), tap * C + c]
    assert torch.equal(A[:, :144], ref) and bool((A[:, 144:] == 0).all())
