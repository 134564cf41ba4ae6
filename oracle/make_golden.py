"""Generate the committed golden vectors in tests/golden (TEST INFRASTRUCTURE).

Run ONLY in the authoring container (needs /root/reference):   python -m oracle.make_golden
Every vector is an OUTPUT OF THE REFERENCE'S OWN PYTHON (point_utils.py / aff.py imported through
oracle/ref_loader.py with the canonical tie rules forced), never of the oracle restatement; the
oracle and the CUDA path are both tested against them.
"""
import os

import numpy as np
import torch

from . import aff_oracle as ao
from . import inputs, ref_loader

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _np(t):
    return None if t is None else t.detach().cpu().numpy()


def sfc_cases():
    return {
        "grid32x32_m8": dict(h=32, w=32, n=1024, m=8),          # k*m == n, heavy key ties on the grid
        "rand655_64x64_m8": dict(h=64, w=64, n=655, m=8),       # padded last cluster -> mask
        "rand512_32x64_m24": dict(h=32, w=64, n=512, m=24),     # Base-style m=24, non-square
        "rand1021_48x40_m8": dict(h=48, w=40, n=1021, m=8),     # odd sizes
    }


def make_sfc(pu):
    for name, c in sfc_cases().items():
        pos = (inputs.grid_positions(2, c["h"], c["w"]) if c["n"] == c["h"] * c["w"]
               else inputs.random_positions(2, c["n"], c["h"], c["w"], seed=7))
        with ref_loader.canonical_ties():
            p, mean, member, mask, rank = pu.space_filling_cluster(pos, c["m"], c["h"], c["w"])
        np.savez_compressed(os.path.join(OUT, f"sfc_{name}.npz"), pos_in=_np(pos).astype(np.int16),
                            h=c["h"], w=c["w"], m=c["m"], pos=_np(p), mean=_np(mean),
                            member=_np(member).astype(np.int32),
                            mask=(np.zeros(0, np.int8) if mask is None else _np(mask).astype(np.int8)),
                            rank=_np(rank).astype(np.int32))
        print("sfc", name)


def make_shepard(pu):
    """upsample_feature_shepard through the reference's PyTorch path (custom_kernel=False,
    point_utils.py:116-117) and shepard_decay_weights (point_utils.py:63-75)."""
    g = torch.Generator().manual_seed(3)
    query = inputs.grid_positions(2, 16, 16)
    database = inputs.random_positions(2, 64, 16, 16, seed=11)
    feature = torch.randn(2, 64, 24, generator=g)
    with ref_loader.canonical_ties():
        up = pu.upsample_feature_shepard(query, database, feature, custom_kernel=False)
        w = pu.upsample_feature_shepard(query, database, feature, return_weight_only=True)
    np.savez_compressed(os.path.join(OUT, "shepard_16x16_from64.npz"), query=_np(query), database=_np(database),
                        feature=_np(feature), up=_np(up), weights=_np(w))
    print("shepard")


def make_wg():
    """The reference's own check of WEIGHTEDGATHER (clusten/test_wg_kernel.py:14-44), seeded and
    shrunk: forward and both gradients of the PyTorch gather formulation."""
    g = torch.Generator().manual_seed(5)
    b, n, n_, k, c = 3, 50, 100, 4, 32
    nn_idx = torch.randint(n_, (b, n, k), generator=g)
    w = torch.rand(b, n, k, generator=g).requires_grad_(True)
    f = torch.rand(b, n_, c, generator=g).requires_grad_(True)
    nn_features = f.gather(index=nn_idx.view(b, -1).unsqueeze(2).expand(-1, -1, c), dim=1).reshape(b, n, k, c)
    up = nn_features.mul(w.unsqueeze(3).expand(-1, -1, -1, c)).sum(dim=2)
    up.mean().backward()
    np.savez_compressed(os.path.join(OUT, "wg_refcheck.npz"), idx=_np(nn_idx).astype(np.int32), w=_np(w), f=_np(f),
                        up=_np(up), d_w=_np(w.grad), d_f=_np(f.grad))
    print("wg")


# (preset, B, H, W) of the reference-class goldens: the toy preset plus the BASELINE presets at the shapes the bench runs
AFF_GOLDENS = {
    "aff_test_256": ("test", 2, 256, 256),
    "aff_mini_512": ("mini", 2, 512, 512),                 # configs[1] backbone: 16 384 tokens, ds 0.25
    "aff_tiny_1_5_512": ("tiny_1_5", 2, 512, 512),         # configs[2]: ds 0.2 -> 3276 / 655 / 131 tokens, padded clusters, masks
    "aff_base_256x512": ("base", 1, 256, 512),             # configs[4]'s model: m = 24, M = 144 (every stage padded)
}


def make_aff(aff, name):
    """Reference AFF class forward (aff.py:568-686), closed-form weights and images."""
    preset, B, H, Wd = AFF_GOLDENS[name]
    cfg = ao.PRESETS[preset]
    W = ao.synthetic_state(cfg)
    m = aff.AFF(embed_dim=cfg["embed_dim"], cluster_size=cfg["cluster_size"], nbhd_size=list(cfg["nbhd_size"]),
                alpha=cfg["alpha"], ds_rate=cfg["ds_rate"], depths=cfg["depths"], num_heads=cfg["num_heads"],
                mlp_ratio=cfg["mlp_ratio"], drop_path_rate=0.0, layer_scale=cfg["layer_scale"])
    sd = m.state_dict()
    assert set(sd) - set(W) == {"patch_embed.bn.num_batches_tracked"} and not (set(W) - set(sd)), "param names drifted"
    W2 = dict(W)
    W2["patch_embed.bn.num_batches_tracked"] = sd["patch_embed.bn.num_batches_tracked"]
    m.load_state_dict(W2)
    m.eval()
    x = ao.synthetic_images(B, H, Wd)
    with torch.no_grad(), ref_loader.canonical_ties():
        out = m(x)
    save = {}
    for i in range(2, 6):
        f = out[f"res{i}"]
        save[f"res{i}_pos"] = _np(out[f"res{i}_pos"]).astype(np.int16)
        # keep fixtures small: every stride-th token, all channels (+ global checksums)
        stride = max(1, f.shape[1] // 64)
        save[f"res{i}_stride"] = stride
        save[f"res{i}_sub"] = _np(f[:, ::stride])
        save[f"res{i}_sum"] = _np(f.double().sum())
        save[f"res{i}_abs"] = _np(f.double().abs().sum())
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **save)
    print("aff", name, {k: tuple(v.shape) for k, v in out.items() if torch.is_tensor(v) and not k.endswith("pos")})


def msdeform_case():
    """Inputs of the MSDeformAttnPc golden: three levels of tokens on a 24 x 32 stem grid (AFF reports the stem grid as the spatial
    shape of every level), seeded weights with non-zero sampling offsets."""
    g = torch.Generator().manual_seed(17)
    b, c, heads, levels, points = 2, 32, 4, 3, 4
    hw = (24, 32)
    ns = [40, 150, 500]                                    # res5, res4, res3 token counts
    poss = [inputs.random_positions(b, n, hw[0], hw[1], seed=20 + i) for i, n in enumerate(ns)]
    querys = [torch.randn(b, n, c, generator=g) for n in ns]
    values = [torch.randn(b, n, c, generator=g) for n in ns]
    params = {"sampling_offsets.weight": torch.randn(heads * levels * points * 2, c, generator=g) * 0.3,
              "attention_weights.weight": torch.randn(heads * levels * points, c, generator=g) * 0.5,
              "attention_weights.bias": torch.randn(heads * levels * points, generator=g) * 0.5}
    return dict(b=b, c=c, heads=heads, levels=levels, points=points, hw=hw, ns=ns, poss=poss, querys=querys, values=values, params=params)


def make_msdeform():
    """The reference's MSDeformAttnPc class (msdeformattn_pc.py:107-205) with its lookup tables built as forward_features does
    (:486-502), forward and the gradients of a quadratic loss."""
    cls, scale_pos = ref_loader.load_msdeformattn_pc()
    pu, _ = ref_loader.load()
    c = msdeform_case()
    torch.manual_seed(3)
    m = cls(c["c"], c["levels"], c["heads"], c["points"], 4.0, True)
    sd = m.state_dict()
    sd.update(c["params"])
    m.load_state_dict(sd)
    hw = c["hw"]
    ys, xs = torch.meshgrid(torch.arange(hw[0]), torch.arange(hw[1]), indexing="ij")
    grid_pos = torch.stack([xs, ys], dim=2).reshape(1, -1, 2).expand(c["b"], -1, -1).float()
    ss = [hw] * (c["levels"] + 1)
    with ref_loader.canonical_ties():
        nb_idx = [pu.knn_keops(grid_pos, scale_pos(p, hw, hw, no_bias=True), 4) for p in c["poss"]]
    qs = [q.clone().requires_grad_(True) for q in c["querys"]]
    vs = [v.clone().requires_grad_(True) for v in c["values"]]
    outs = m(qs, c["poss"], vs, ss, nb_idx)
    sum(o.square().mean() for o in outs).backward()
    save = {"state." + k: _np(v) for k, v in m.state_dict().items()}
    for i in range(c["levels"]):
        save[f"out{i}"], save[f"d_query{i}"], save[f"d_value{i}"] = _np(outs[i]), _np(qs[i].grad), _np(vs[i].grad)
        save[f"nb_idx{i}"] = _np(nb_idx[i]).astype(np.int32)
    for k, p in m.named_parameters():
        save["grad." + k] = _np(p.grad)
    np.savez_compressed(os.path.join(OUT, "msdeformattn_pc.npz"), **save)
    print("msdeformattn_pc", [tuple(o.shape) for o in outs])


def main():
    os.makedirs(OUT, exist_ok=True)
    pu, aff = ref_loader.load()
    import sys
    if len(sys.argv) < 2:
        make_sfc(pu)
        make_shepard(pu)
        make_wg()
    if len(sys.argv) < 2 or "msdeformattn_pc" in sys.argv[1:]:
        make_msdeform()
    for name in AFF_GOLDENS:
        if len(sys.argv) < 2 or name in sys.argv[1:]:
            make_aff(aff, name)


if __name__ == "__main__":
    main()
