"""Host-side logic of the Python layer that needs no kernel (this suite has no GPU): the dispatch predicates of the glue ops,
their dtype-promotion rule, the op-by-op torch formulation used for operands outside a kernel's shape / dtype set (it is also
what the GPU parity tests compare the kernels with), and the module wrappers that keep the reference's parameter names.
None of this is a CPU path of the product: the CLUSTEN ops refuse non-CUDA tensors (tests/test_abi.py)."""
import torch
import torch.nn as nn

from autofocusformermod_b200 import aff, ops


def test_scale_residual_formulation_is_the_reference_block_line():
    """`x = shortcut + drop_path(gamma * x)` (aff.py:230,236 with timm's DropPath): the torch formulation scale_residual uses for
    operands outside the kernel's set equals the reference line for the same per-sample draw (evaluated on host tensors)."""
    g = torch.Generator().manual_seed(0)
    B, n, C = 4, 7, 12
    feat, x, gamma = torch.randn(B, n, C, generator=g), torch.randn(B, n, C, generator=g), torch.randn(C, generator=g)
    keep = 0.7
    mask = torch.bernoulli(torch.full((B, 1, 1), keep), generator=g)
    want = feat + (gamma * x) * (mask / keep)                                         # timm 0.6.12 drop_path, scale_by_keep
    got = ops.scale_residual(feat, x, gamma, (mask / keep).reshape(B))
    assert torch.equal(got, want)
    assert torch.equal(ops.scale_residual(feat, x), feat + x)
    assert torch.equal(ops.scale_residual(feat, x, gamma), feat + gamma * x)
    assert not ops.scale_residual_supported(feat, x, gamma, None)                     # host tensors are outside the kernel's set


def test_scale_residual_dtype_promotion_follows_aten():
    """The kernel's output dtype rule (_residual_out_dtype) is ATen's promotion of `res + x * gamma`: an fp32 layer scale lifts a
    16-bit residual stream to fp32 (otherwise a 1e-5 layer scale would vanish in bf16), 16-bit + 16-bit stays 16-bit."""
    bf, f32 = torch.bfloat16, torch.float32
    for rdt, xdt, gdt in [(f32, f32, None), (f32, bf, None), (bf, bf, None), (bf, bf, f32), (f32, bf, f32), (bf, bf, bf)]:
        res, x = torch.zeros(2, 3, 4, dtype=rdt), torch.zeros(2, 3, 4, dtype=xdt)
        gamma = None if gdt is None else torch.ones(4, dtype=gdt)
        want = (res + (x if gamma is None else gamma * x)).dtype
        assert ops._residual_out_dtype(res, x, gamma) == want, (rdt, xdt, gdt)


def test_drop_path_sample_scale():
    dp = aff.DropPath(0.25)
    x = torch.zeros(4000, 2, 3)
    dp.train()
    s = dp.sample_scale(x)
    assert s.shape == (4000,) and s.dtype == torch.float32
    vals = torch.unique(s)
    assert vals.numel() == 2 and vals[0] == 0 and abs(float(vals[1]) - 1 / 0.75) < 1e-6
    assert abs(float((s > 0).float().mean()) - 0.75) < 0.03
    dp.eval()
    assert dp.sample_scale(x) is None and dp(x) is x
    assert aff.DropPath(0.0).train().sample_scale(x) is None


def test_module_wrappers_keep_reference_parameter_names_and_cpu_semantics():
    """Linear / TableLinear / LayerNorm subclass the torch modules: same state_dict keys (reference checkpoints load) and the
    torch forward for operands outside the kernels' set."""
    torch.manual_seed(0)
    lin, ref = aff.Linear(8, 6), nn.Linear(8, 6)
    ref.load_state_dict(lin.state_dict())
    x = torch.randn(5, 8)
    assert torch.equal(lin(x), ref(x))
    tl, tref = aff.TableLinear(5, 3), nn.Linear(5, 3)
    tref.load_state_dict(tl.state_dict())
    f = torch.randn(11, 5)
    assert torch.equal(tl(f, None), tref(f))
    assert not ops.table_linear_supported(f, tl.weight, tl.bias)
    ln, lref = aff.LayerNorm(8), nn.LayerNorm(8)
    lref.load_state_dict(ln.state_dict())
    assert torch.equal(ln(x), lref(x))
    blk = aff.ClusterTransformerBlock(32, 2, layer_scale=1e-5, drop_path=0.1)
    keys = set(blk.state_dict())
    assert {"gamma1", "gamma2", "attn.pos_embed.weight", "attn.pos_embed.bias", "attn.q.weight", "attn.kv.bias", "attn.blank_k",
            "mlp.fc1.weight", "norm1.weight"} <= keys
    merge = aff.ClusterMerging(32, 64)
    assert {"weight_net.0.weight", "weight_net.1.weight", "weight_net.1.bias", "norm.weight", "linear.weight"} <= set(merge.state_dict())


def test_layernorm_linear_bound_holds_for_any_input():
    """aff._ln_linear_bound: what lets proj / fc2 use the fp16 form of the tcgen05 split.  The bound comes from
    the parameters alone and must dominate |Linear(LayerNorm(x))| for every x -- including rows with huge, tiny and constant values."""
    g = torch.Generator().manual_seed(3)
    ln, lin = aff.LayerNorm(64), aff.Linear(64, 96)
    with torch.no_grad():
        ln.weight.copy_(torch.randn(64, generator=g) * 2.0)
        ln.bias.copy_(torch.randn(64, generator=g))
        lin.weight.copy_(torch.randn(96, 64, generator=g))
        lin.bias.copy_(torch.randn(96, generator=g) * 3.0)
    extra = torch.nn.Parameter(torch.tensor([0.5, -7.0]))
    bound = aff._ln_linear_bound(lin, ln)
    x = torch.cat([torch.randn(200, 64, generator=g) * s for s in (1e-6, 1.0, 1e6)] + [torch.full((1, 64), 3.0)])
    x[5, 7] = 1e9                                                                     # one dominating element: the normalised row is ~ sqrt(K) e_7
    y = lin(torch.nn.functional.layer_norm(x, (64,), ln.weight, ln.bias, ln.eps))
    assert float(y.abs().max()) <= bound
    assert aff._ln_linear_bound(lin, ln, extra) == max(bound, 7.0)
    assert aff._ln_linear_bound(lin, torch.nn.Identity()) is None                     # no LayerNorm in front: no bound
    with torch.no_grad():
        lin.weight.mul_(2.0)                                                          # in-place update -> new version -> recomputed
    assert aff._ln_linear_bound(lin, ln) > 1.5 * bound


def test_kernel_dispatch_predicates_reject_what_the_kernels_do_not_take():
    x = torch.zeros(4, 48)
    w = torch.zeros(6, 48)
    assert not ops.linear_tc_supported(x, w, None)                                    # not a CUDA tensor
    assert not ops.linear_f32_supported(x, w, None)                                   # not a CUDA tensor
    assert not ops.linear_f32_supported(x.to(torch.bfloat16), w, None)
    assert not ops.table_linear_supported(torch.zeros(4, 9), torch.zeros(3, 9), None)
    res = torch.zeros(2, 5, 30)
    assert not ops.scale_residual_supported(res, res, None, torch.ones(2))


def test_layer_norm_fold_is_the_same_layer():
    """ops.ln_folded_layer: LayerNorm(gamma, beta) followed by Linear(W, b) equals plain normalisation followed by Linear(W * gamma,
    b + W beta) -- the identity behind CLUSTEN_LN_FOLD (the tcgen05 Linear's split warps then only compute (x - mean) * rstd).
    Evaluated in float64 on host tensors; the folded operands are cached until a parameter changes."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(1)
    K, N = 96, 40
    x = torch.randn(50, K, generator=g, dtype=torch.float64) * 2.0 + 0.7
    w = torch.randn(N, K, generator=g) * K ** -0.5
    b = torch.randn(N, generator=g)
    lw, lb = torch.rand(K, generator=g) + 0.5, torch.randn(K, generator=g)
    want = F.linear(F.layer_norm(x, (K,), lw.double(), lb.double(), 1e-5), w.double(), b.double())
    wf, bf = ops.ln_folded_layer(w, b, lw, lb)
    norm = (x - x.mean(1, keepdim=True)) * (x.var(1, unbiased=False, keepdim=True) + 1e-5).rsqrt()
    got = F.linear(norm, wf.double(), bf.double())
    assert wf.dtype == torch.float32 and bf.dtype == torch.float32
    assert float((got - want).abs().max() / want.abs().max()) <= 5e-7                # fp32 rounding of W' and b' only
    assert ops.ln_folded_layer(w, b, lw, lb)[0] is wf                                # cached
    lb.add_(1.0)
    wf2, bf2 = ops.ln_folded_layer(w, b, lw, lb)                                     # beta changed: b' follows, W' is rebuilt equal
    assert wf2 is not wf and torch.equal(wf2, wf) and not torch.equal(bf2, bf)
    wn, bn_ = ops.ln_folded_layer(w, None, lw, lb)                                   # a layer without bias still gets b' = W beta
    assert float((bn_.double() - w.double() @ lb.double()).abs().max()) <= 1e-6


def test_stem_gemm_operands_are_the_second_convolution():
    """ops.stem_proj2_weight: the [E, C, 3, 3] convolution weight as [E, (ky, kx, c)] zero-padded to a multiple of 32 columns.  With
    the im2col rows A[(b, py, px), (ky * 3 + kx) * C + c] = mid[b, 2 py - 1 + ky, 2 px - 1 + kx, c] (what clusten_stem_im2col writes;
    built here with F.unfold) the GEMM A W'^T + bias IS conv2d(stride 2, padding 1) in token-major order (aff.py:549-552)."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(2)
    B, C, E, H, W = 2, 16, 32, 10, 14
    conv = nn.Conv2d(C, E, 3, stride=2, padding=1).double()
    mid = torch.randn(B, C, H, W, generator=g, dtype=torch.float64)
    want = conv(mid).flatten(2).transpose(1, 2)                                      # [B, h * w, E]
    conv32 = nn.Conv2d(C, E, 3, stride=2, padding=1)
    conv32.load_state_dict({k: v.float() for k, v in conv.state_dict().items()})
    w2 = ops.stem_proj2_weight(conv32)
    assert w2.shape == (E, 160) and bool((w2[:, 144:] == 0).all()) and ops.stem_proj2_weight(conv32) is w2
    cols = F.unfold(mid, kernel_size=3, stride=2, padding=1)                         # [B, C * 9, L], row c * 9 + tap
    A = cols.view(B, C, 9, -1).permute(0, 3, 2, 1).reshape(B, -1, 9 * C)             # [B, L, tap * C + c]
    got = A @ w2[:, :144].double().t() + conv32.bias.double()
    assert got.shape == want.shape and float((got - want).abs().max()) <= 1e-6


def test_new_fast_paths_keep_torch_semantics_outside_their_operand_set():
    """The row gather, the stem kernels and the merge-score / table-feature kernels only take CUDA operands without a gradient;
    everything else goes through the torch formulation with the same values (evaluated on host tensors here)."""
    g = torch.Generator().manual_seed(3)
    src = torch.randn(2, 9, 5, generator=g)
    idx = torch.randint(0, 9, (2, 4, 1), generator=g)
    assert torch.equal(ops.gather_rows(src, idx), src.gather(1, idx.expand(-1, -1, 5)))          # host tensors: torch.gather
    assert torch.equal(aff._gather(src, idx), src.gather(1, idx.expand(-1, -1, 5)))
    leaf = src.clone().requires_grad_(True)
    out = ops.gather_rows(leaf, idx)                                                             # a gradient flows: autograd's gather
    out.sum().backward()
    assert leaf.grad is not None and float(leaf.grad.sum()) == 2 * 4 * 5
    conv1, bn, conv2 = nn.Conv2d(3, 16, 3, stride=2, padding=1), nn.BatchNorm2d(16).eval(), nn.Conv2d(16, 32, 3, stride=2, padding=1)
    x = torch.randn(1, 3, 16, 16, generator=g)
    with torch.no_grad():
        assert not ops.stem_conv_bn_gelu_supported(x, conv1, bn)                                 # host tensor
        assert not ops.stem_gemm_supported(x, conv1, bn, conv2)
    rows = torch.tensor([0, 511 * 1023 + 511, 1023 * 1023 - 1, 511 * 1023 + 512])
    f = aff.rel_pos_features(rows)                                                               # host tensor: the torch formulation
    assert f.shape == (4, 5) and torch.equal(f[1], torch.zeros(5))                               # the 0 / 0 centre is zeroed
    assert torch.equal(f[3], torch.tensor([1.0, 0.0, 1.0, 0.0, 1.0]))                            # dx = 1, dy = 0
    assert torch.equal(f[0, :2], torch.tensor([-511.0, -511.0])) and torch.equal(f[2, :2], torch.tensor([511.0, 511.0]))
