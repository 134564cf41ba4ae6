"""CPU restatement of the AFF backbone forward (TEST INFRASTRUCTURE, see oracle/__init__).

Functional form over a reference-named ``state_dict`` (``layers.{i}.blocks.{j}.attn.q.weight`` ...):
the same numbers as ``AFF.forward`` (mask2former/modeling/backbone/aff.py:662-686) computed with
plain torch functional ops and the oracle's clustering / kNN / CLUSTEN restatements.  Validated in
the authoring container against the reference's own ``AFF`` class loaded through
oracle/ref_loader.py (tests/test_oracle_vs_reference.py, golden vectors in tests/golden).

Stochastic depth (timm DropPath) is the identity here (eval mode / DROP_PATH_RATE=0).
"""
import math

import torch
import torch.nn.functional as F

from . import clusten_ops as ops
from . import point_ops as pt

# cfg.MODEL.AFF.* of the reference yaml files (configs/**/aff/*.yaml, defaults config.py:87-104)
PRESETS = {
    "mini":     dict(embed_dim=[32, 128, 256, 384], depths=[2, 2, 6, 2], num_heads=[2, 4, 8, 16], mlp_ratio=2.0,
                     cluster_size=8, nbhd_size=[48] * 4, layer_scale=0.0, alpha=4.0, ds_rate=0.25),
    "tiny_1_5": dict(embed_dim=[64, 128, 256, 512], depths=[3, 4, 18, 5], num_heads=[2, 4, 8, 16], mlp_ratio=3.0,
                     cluster_size=8, nbhd_size=[48] * 4, layer_scale=0.0, alpha=4.0, ds_rate=0.2),
    "small":    dict(embed_dim=[96, 192, 384, 768], depths=[3, 4, 18, 2], num_heads=[3, 6, 12, 24], mlp_ratio=3.0,
                     cluster_size=8, nbhd_size=[48] * 4, layer_scale=1e-5, alpha=8.0, ds_rate=0.25),
    "base":     dict(embed_dim=[128, 256, 512, 1024], depths=[3, 4, 18, 2], num_heads=[4, 8, 16, 32], mlp_ratio=3.0,
                     cluster_size=24, nbhd_size=[144] * 4, layer_scale=1e-5, alpha=8.0, ds_rate=0.25),
    # small test model: Mini's dims with fewer blocks (covers per-head dims 16, 32, 24)
    "test":     dict(embed_dim=[32, 128, 256, 384], depths=[1, 1, 2, 1], num_heads=[2, 4, 8, 16], mlp_ratio=2.0,
                     cluster_size=8, nbhd_size=[48] * 4, layer_scale=1e-5, alpha=4.0, ds_rate=0.25),
}

_pre_table = None


def pre_table():
    global _pre_table
    if _pre_table is None:
        _pre_table = pt.build_pre_table()
    return _pre_table


def param_shapes(cfg):
    """Reference parameter/buffer names and shapes of ``AFF(**cfg)`` (aff.py:53-686)."""
    E, D, Hh = cfg["embed_dim"], cfg["depths"], cfg["num_heads"]
    s = {"patch_embed.proj1.weight": (E[0] // 2, 3, 3, 3), "patch_embed.proj1.bias": (E[0] // 2,),
         "patch_embed.bn.weight": (E[0] // 2,), "patch_embed.bn.bias": (E[0] // 2,),
         "patch_embed.bn.running_mean": (E[0] // 2,), "patch_embed.bn.running_var": (E[0] // 2,),
         "patch_embed.proj2.weight": (E[0], E[0] // 2, 3, 3), "patch_embed.proj2.bias": (E[0],),
         "patch_embed.norm.weight": (E[0],), "patch_embed.norm.bias": (E[0],)}
    for i in range(4):
        c = E[i]
        hid = int(c * cfg["mlp_ratio"])
        for j in range(D[i]):
            p = f"layers.{i}.blocks.{j}."
            s.update({p + "norm1.weight": (c,), p + "norm1.bias": (c,),
                      p + "attn.q.weight": (c, c), p + "attn.q.bias": (c,),
                      p + "attn.kv.weight": (2 * c, c), p + "attn.kv.bias": (2 * c,),
                      p + "attn.blank_k": (c,), p + "attn.blank_v": (c,),
                      p + "attn.pos_embed.weight": (Hh[i], 5), p + "attn.pos_embed.bias": (Hh[i],),
                      p + "attn.proj.weight": (c, c), p + "attn.proj.bias": (c,),
                      p + "norm2.weight": (c,), p + "norm2.bias": (c,),
                      p + "mlp.fc1.weight": (hid, c), p + "mlp.fc1.bias": (hid,),
                      p + "mlp.fc2.weight": (c, hid), p + "mlp.fc2.bias": (c,)})
            if cfg["layer_scale"] and cfg["layer_scale"] > 0:
                s.update({p + "gamma1": (c,), p + "gamma2": (c,)})
        if i < 3:
            p = f"layers.{i}."
            s.update({p + "downsample.weight_net.0.weight": (4, 5), p + "downsample.weight_net.0.bias": (4,),
                      p + "downsample.weight_net.1.weight": (4,), p + "downsample.weight_net.1.bias": (4,),
                      p + "downsample.norm.weight": (4 * c,), p + "downsample.norm.bias": (4 * c,),
                      p + "downsample.linear.weight": (E[i + 1], 4 * c), p + "downsample.linear.bias": (E[i + 1],),
                      p + "prob_net.weight": (1, c), p + "prob_net.bias": (1,)})
        s.update({f"norm{i}.weight": (c,), f"norm{i}.bias": (c,)})
    return s


def synthetic_state(cfg, seed=0):
    """Closed-form pseudo-random weights (no RNG-version dependence): same dict in every process."""
    out = {}
    for n_i, (name, shape) in enumerate(param_shapes(cfg).items()):
        numel = int(math.prod(shape))
        t = torch.arange(numel, dtype=torch.float64)
        v = torch.sin(t * 0.7390851 + 1.6180339 * (n_i + 1) + 0.5 * seed) * math.cos(0.37 * n_i + 1.0)
        fan_in = shape[-1] if len(shape) == 2 else (int(math.prod(shape[1:])) if len(shape) == 4 else 1)
        if name.endswith("running_var"):
            v = 1.0 + 0.25 * v.abs()
        elif ".norm" in name and name.endswith("weight") or name.endswith("bn.weight") or "weight_net.1.weight" in name \
                or (name.startswith("norm") and name.endswith("weight")):
            v = 1.0 + 0.1 * v
        elif "gamma" in name:
            v = 0.5 + 0.1 * v          # layer-scale large enough that the attention branch matters in tests
        elif len(shape) >= 2:
            v = v * (1.5 / math.sqrt(fan_in))
        else:
            v = 0.2 * v
        out[name] = v.to(torch.float32).reshape(shape)
    return out


def synthetic_images(B, H, W, seed=0):
    t = torch.arange(B * 3 * H * W, dtype=torch.float64)
    return torch.sin(t * 0.0123457 + seed).mul(1.5).add(torch.cos(t * 0.41 + 2.0 * seed)).to(torch.float32).reshape(B, 3, H, W)


def _ln(x, W, p):
    return F.layer_norm(x, (x.shape[-1],), W[p + ".weight"], W[p + ".bias"])


def _lin(x, W, p):
    return F.linear(x, W[p + ".weight"], W[p + ".bias"])


def patch_embed(x, W, training=False):
    """aff.py:538-565"""
    _, _, H, Wd = x.shape
    if Wd % 4:
        x = F.pad(x, (0, 4 - Wd % 4))
    if H % 4:
        x = F.pad(x, (0, 0, 0, 4 - H % 4))
    x = F.conv2d(x, W["patch_embed.proj1.weight"], W["patch_embed.proj1.bias"], stride=2, padding=1)
    x = F.batch_norm(x, None if training else W["patch_embed.bn.running_mean"],
                     None if training else W["patch_embed.bn.running_var"],
                     W["patch_embed.bn.weight"], W["patch_embed.bn.bias"], training=training)
    x = F.conv2d(F.gelu(x), W["patch_embed.proj2.weight"], W["patch_embed.proj2.bias"], stride=2, padding=1)
    b, c, h, w = x.shape
    x = _ln(x.flatten(2).transpose(1, 2), W, "patch_embed.norm")
    ys, xs = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    pos = torch.stack([xs, ys], dim=2).unsqueeze(0).expand(b, -1, -1, -1).reshape(b, -1, 2).to(x.dtype)
    return pos, x, h, w


def cluster_attention(feat, W, p, heads, member_idx, mask, pe_idx, global_attn):
    """aff.py:86-160"""
    b, n, c = feat.shape
    c_ = c // heads
    q = _lin(feat, W, p + ".q") * (c_ ** -0.5)
    kv = _lin(feat, W, p + ".kv")
    q = q.reshape(b, n, heads, c_).permute(0, 2, 1, 3)
    kv = kv.view(b, n, heads, 2, c_).permute(3, 0, 2, 1, 4)
    key, v = kv[0], kv[1]
    if global_attn:
        attn = q @ key.transpose(-1, -2)
        mask = None
    else:
        attn = ops.qk_forward(q, key, member_idx)                                   # :114
        if mask is not None:
            mask = mask.reshape(b, 1, n, -1)
    pe_table = _lin(pre_table(), W, p + ".pos_embed")                               # :129
    pos_embed = pe_table[pe_idx.reshape(-1)].reshape(*pe_idx.shape, heads).permute(0, 3, 1, 2)
    attn = attn + pos_embed
    if mask is not None:
        attn = attn + (1 - mask) * (-100)                                           # :137
    blank_attn = (q * W[p + ".blank_k"].reshape(1, heads, 1, c_)).sum(-1, keepdim=True)
    attn = torch.cat([attn, blank_attn], dim=-1).softmax(dim=-1)                    # :140-142
    blank_attn, attn = attn[..., -1:], attn[..., :-1]
    blank_v = blank_attn * W[p + ".blank_v"].reshape(1, heads, 1, c_)
    if global_attn:
        out = (attn @ v)
    else:
        out = ops.av_forward(attn, v, member_idx)                                   # :154
    out = (out + blank_v).permute(0, 2, 1, 3).reshape(b, n, c)
    return _lin(out, W, p + ".proj")


def block(feat, W, p, heads, layer_scale, member_idx, mask, pe_idx, global_attn):
    """aff.py:205-238 (DropPath = identity)"""
    a = cluster_attention(_ln(feat, W, p + ".norm1"), W, p + ".attn", heads, member_idx, mask, pe_idx, global_attn)
    feat = feat + (W[p + ".gamma1"] * a if layer_scale else a)
    mlp = _lin(F.gelu(_lin(_ln(feat, W, p + ".norm2"), W, p + ".mlp.fc1")), W, p + ".mlp.fc2")
    return feat + (W[p + ".gamma2"] * mlp if layer_scale else mlp)


def cluster_merging(pos, feat, W, p, member_idx, mask, learned_prob, stride, pe_idx, reserve_num, alpha, ds_rate):
    """aff.py:276-365"""
    b, n, c = feat.shape
    M = member_idx.shape[-1]
    idx = pt.merge_select(pos, learned_prob, stride, alpha, ds_rate, reserve_num)   # :292-329
    n2 = idx.shape[1]
    pos = pos.gather(1, idx.expand(-1, -1, 2))
    member_idx = member_idx.gather(1, idx.expand(-1, -1, M))
    pe_idx = pe_idx.gather(1, idx.expand(-1, -1, M))
    if mask is not None:
        mask = mask.gather(1, idx.expand(-1, -1, M))
    lp = learned_prob.gather(1, member_idx.reshape(b, -1, 1)).reshape(b, n2, M, 1)  # :340
    wt = _lin(pre_table(), W, p + ".weight_net.0")
    wt = F.gelu(F.layer_norm(wt, (4,), W[p + ".weight_net.1.weight"], W[p + ".weight_net.1.bias"]))
    weights = wt[pe_idx.reshape(-1)].reshape(b, n2, M, 4)                           # :349
    if mask is not None:
        lp = lp * mask.unsqueeze(3)
    weights = weights * lp                                                          # :352-355
    out = ops.wf_forward(weights, feat, member_idx).reshape(b, n2, -1)              # :361
    out = _lin(_ln(out, W, p + ".norm"), W, p + ".linear")
    return pos, out, idx


def basic_layer(pos, feat, W, i, cfg, h, w, stride, trace=None):
    """aff.py:427-507 (non-cached clustering branch)"""
    b, n, d = pos.shape
    m = cfg["cluster_size"]
    heads = cfg["num_heads"][i]
    p = f"layers.{i}"
    if cfg["nbhd_size"][i] >= n:
        global_attn, member_idx, mask = True, None, None
        rel = (pos[:, None, :, :] + pt.REL_POS_WIDTH) - pos[:, :, None, :]
        rel = rel.clamp(0, pt.TABLE_WIDTH - 1)
        pe_idx = (rel[..., 1] * pt.TABLE_WIDTH + rel[..., 0]).long()
    else:
        global_attn = False
        k = int(math.ceil(n / float(m)))
        nnc = min(int(round(cfg["nbhd_size"][i] / float(m))), k)
        pos, mean_pos, member, cmask, reorder = pt.space_filling_cluster(pos, m, h, w)          # :469
        feat = feat.gather(1, reorder.expand(-1, -1, feat.shape[2]))                             # :471
        member_idx, mask, pe_idx = pt.assemble_neighbourhood(pos, mean_pos, member, cmask, nnc)  # :475-485
        if trace is not None:
            trace[f"stage{i}.reorder"] = reorder
            trace[f"stage{i}.member_idx"] = member_idx
    for j in range(cfg["depths"][i]):
        feat = block(feat, W, f"{p}.blocks.{j}", heads, bool(cfg["layer_scale"]), member_idx, mask, pe_idx, global_attn)
    if i < 3:
        learned_prob = _lin(feat, W, p + ".prob_net").sigmoid()                                  # :496
        reserve_num = math.ceil(h / (stride * 2)) * math.ceil(w / (stride * 2))                  # :497
        pos_down, feat_down, sel = cluster_merging(pos, feat, W, p + ".downsample", member_idx, mask, learned_prob,
                                                   stride, pe_idx, reserve_num, cfg["alpha"], cfg["ds_rate"])
        if trace is not None:
            trace[f"stage{i}.select"] = sel
        return pos, feat, pos_down, feat_down
    return pos, feat, pos, feat


def aff_forward(x, W, cfg, training=False, trace=None):
    """aff.py:662-686: returns {res2..res5, res*_pos, res*_spatial_shape}."""
    pos, x, h, w = patch_embed(x, W, training)
    outs = {}
    for i in range(4):
        pos_out, x_out, pos, x = basic_layer(pos, x, W, i, cfg, h, w, 2 ** (i + 1), trace)
        outs[f"res{i + 2}"] = _ln(x_out, W, f"norm{i}")
        outs[f"res{i + 2}_pos"] = pos_out
        outs[f"res{i + 2}_spatial_shape"] = (h, w)
    return outs


def point_conv(x, pos, W, p):
    """PointConv.forward of the point-cloud pixel decoder (pixel_decoder/msdeformattn_pc.py:285-314): kNN-9 among the points
    themselves, weight_net over the relative-position table rows, CLUSTENWF, LayerNorm, Linear.  ``W`` = reference-named
    state dict with prefix ``p`` (weight_net.0 / weight_net.1 / norm / linear)."""
    b, n, c = x.shape
    nn_idx = pt.knn(pos, pos, 9)
    nn_pos = pos.gather(index=nn_idx.view(b, -1, 1).expand(-1, -1, 2), dim=1).reshape(b, n, 9, 2)
    rel_pos = pos.unsqueeze(2) - nn_pos
    rel = (rel_pos.long() + pt.REL_POS_WIDTH).clamp(0, pt.TABLE_WIDTH - 1)
    pe_idx = rel[..., 1] * pt.TABLE_WIDTH + rel[..., 0]
    t = F.gelu(_ln(_lin(pre_table(), W, p + "weight_net.0"), W, p + "weight_net.1"))
    weights = t[pe_idx.reshape(-1)].reshape(b, n, 9, -1)
    feat = ops.wf_forward(weights, x, nn_idx).reshape(b, n, -1)
    return _lin(_ln(feat, W, p + "norm"), W, p + "linear")
