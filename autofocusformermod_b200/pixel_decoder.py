"""Point-cloud pixel-decoder pieces that sit on the CLUSTEN path (mask2former/modeling/pixel_decoder/msdeformattn_pc.py):

    PointConv   (msdeformattn_pc.py:271-314)   kNN-9 self neighbourhood -> relative-position weights -> CLUSTENWF -> LN -> Linear
    point2img   (transformer_decoder/mask2former_transformer_decoder.py:20-39)   per-point masks -> dense [b,q,h,w] image
    point_attn_mask  (mask2former_transformer_decoder.py:484-486)   Shepard upsampling of the mask logits to a decoder level
                     (kNN-4 + WEIGHTEDGATHER on libclusten_b200) -> boolean attention mask

Same class name, constructor arguments, parameter names (``weight_net``, ``norm``, ``linear``) and call convention
(``forward((x, pos))``) as the reference, so a reference state_dict loads unchanged.  The kNN, the weighted-feature merge
and the table lookup run on libclusten_b200; ``weight_net`` is evaluated on the table rows the batch references
(the reference evaluates it on all 1023^2 rows per call, msdeformattn_pc.py:302).
"""
import math

import torch
from torch import nn
from torch.nn import functional as F

from .aff import REL_POS_WIDTH, TABLE_WIDTH, LayerNorm, Linear, _TableLookup
from .ops import CLUSTENWFFunction, MSDETRPCFunction
from .point_utils import knn_keops, shepard_decay_weights, table_rank, upsample_feature_shepard


class PointConv(nn.Module):
    def __init__(self, dim, out_dim, bias):
        super().__init__()
        inner_ch = 4
        self.weight_net = nn.Sequential(nn.Linear(5, inner_ch, bias=True), nn.LayerNorm(inner_ch), nn.GELU())
        # (aff.LayerNorm / aff.Linear: nn.LayerNorm / nn.Linear with the same parameters; in fp32 inference the pair runs as one
        # tcgen05 GEMM with the norm applied while the rows are staged)
        self.norm = LayerNorm(inner_ch * dim)
        self.linear = Linear(dim * inner_ch, out_dim, bias=bias)

    def forward(self, inp):
        """inp = (x [b,n,c] point features, pos [b,n,2] point positions) -> [b,n,out_dim]"""
        x, pos = inp
        b, n, c = x.shape
        nn_idx = knn_keops(pos, pos, 9)                                                        # msdeformattn_pc.py:295
        nn_pos = pos.gather(index=nn_idx.view(b, -1, 1).expand(-1, -1, 2), dim=1).reshape(b, n, 9, 2)
        rel_pos = pos.unsqueeze(2) - nn_pos                                                    # :297 (query minus neighbour)
        rel = (rel_pos.long() + REL_POS_WIDTH).clamp(0, TABLE_WIDTH - 1)                       # :305
        pe_idx = rel[..., 1] * TABLE_WIDTH + rel[..., 0]                                       # :306
        # :302-308 on the referenced table rows, ranked on the device (no sort, no host read: the decoder stays capturable)
        uniq, inverse, count = table_rank(pe_idx)
        weights = _TableLookup(uniq=uniq, inverse=inverse, count=count)(self.weight_net)
        feat = CLUSTENWFFunction.apply(weights, x, nn_idx).reshape(b, n, -1)                   # :309
        return self.linear.fused(feat, "bias", norm=self.norm)                                  # :311-313


def scale_pos(last_pos, last_ss, cur_ss, no_bias=False):
    """Positions of a (h, w) = last_ss grid expressed on a cur_ss grid (msdeformattn_pc.py:28-53); ``no_bias`` scales about the
    cell centres.  AFF reports the stem grid as the spatial shape of every stage (aff.py:679-685), so inside the AFF pixel decoder
    this is the identity; other backbones of the reference tree use it with real ratios."""
    if last_ss[0] == cur_ss[0] and last_ss[1] == cur_ss[1]:
        return last_pos
    w_ratio, h_ratio = cur_ss[1] / last_ss[1], cur_ss[0] / last_ss[0]
    ret = last_pos + 0.5 if no_bias else last_pos
    ret = torch.stack((ret[..., 0] * w_ratio, ret[..., 1] * h_ratio), dim=-1)
    return ret - 0.5 if no_bias else ret


def grid_lookup_tables(poss, spatial_shapes, grid_hw):
    """The kNN-4 lookup tables of ``MSDeformAttnPixelDecoder.forward_features`` (msdeformattn_pc.py:486-502): for every cell of the
    finest grid, the 4 nearest tokens of each level (positions rescaled to that grid).  -> list of int64 [b, h*w, 4]"""
    b, dev = poss[0].shape[0], poss[0].device
    ys, xs = torch.meshgrid(torch.arange(grid_hw[0], device=dev), torch.arange(grid_hw[1], device=dev), indexing="ij")
    grid_pos = torch.stack([xs, ys], dim=2).reshape(1, -1, 2).expand(b, -1, -1).float().contiguous()
    return [knn_keops(grid_pos, scale_pos(pos.float(), ss, grid_hw, no_bias=True).contiguous(), 4) for pos, ss in zip(poss, spatial_shapes)]


class MSDeformAttnPc(nn.Module):
    """Multi-scale deformable attention over point clouds (msdeformattn_pc.py:107-205): every query predicts ``n_points`` sampling
    locations per level and head; a location is resolved to the 4 nearest tokens of that level through the grid lookup table
    (``nb_idx``, no kNN at run time), interpolated with inverse-distance weights, and the ``n_levels * n_points`` samples are mixed
    with softmax weights -- the two-level weighted gather MSDETRPCFunction (clusten_msdetrpc_fwd / _bwd).  Same constructor
    arguments, parameter names and initialisation as the reference, so its state_dict loads unchanged."""

    def __init__(self, d_model, n_levels, n_heads, n_points, shepard_power, shepard_power_learnable):
        super().__init__()
        self.d_model, self.n_levels, self.n_heads, self.n_points = d_model, n_levels, n_heads, n_points
        self.sampling_offsets = Linear(d_model, n_heads * n_levels * n_points * 2)
        self.attention_weights = Linear(d_model, n_heads * n_levels * n_points)
        self.value_proj = Linear(d_model, d_model)
        self.output_proj = Linear(d_model, d_model)
        self.shepard_power = nn.Parameter(shepard_power * torch.ones(1)) if shepard_power_learnable else shepard_power
        self._reset_parameters()

    def _reset_parameters(self):                                                                   # msdeformattn_pc.py:127-141
        nn.init.constant_(self.sampling_offsets.weight.data, 0.0)
        thetas = torch.arange(self.n_heads, dtype=torch.float32) * (2.0 * math.pi / self.n_heads)
        grid = torch.stack([thetas.cos(), thetas.sin()], -1)
        grid = (grid / grid.abs().max(-1, keepdim=True)[0]).view(self.n_heads, 1, 1, 2).repeat(1, self.n_levels, self.n_points, 1)
        grid = grid * torch.arange(1, self.n_points + 1, dtype=torch.float32).view(1, 1, -1, 1)
        with torch.no_grad():
            self.sampling_offsets.bias = nn.Parameter(grid.reshape(-1))
        nn.init.constant_(self.attention_weights.weight.data, 0.0)
        nn.init.constant_(self.attention_weights.bias.data, 0.0)
        nn.init.xavier_uniform_(self.value_proj.weight.data)
        nn.init.constant_(self.value_proj.bias.data, 0.0)
        nn.init.xavier_uniform_(self.output_proj.weight.data)
        nn.init.constant_(self.output_proj.bias.data, 0.0)

    def forward(self, querys, poss, values, spatial_shapes, nb_idx):
        """querys / values: per level [b, n_l, c]; poss: per level [b, n_l, 2]; spatial_shapes: n_levels + 1 (h, w) pairs, the last
        one the lookup grid; nb_idx: per level int64 [b, h*w, 4] (grid_lookup_tables).  Returns the per-level outputs [b, n_l, c]."""
        b, _, c = querys[0].shape
        h, l, k = self.n_heads, self.n_levels, self.n_points
        c_ = c // h
        grid_hw = spatial_shapes[-1]
        # values of all levels, one [b*h, sum n_l, c_] operand: the second-level indices address it with per-level offsets (:161,191)
        val = self.value_proj(torch.cat(values, dim=1)).reshape(b, -1, h, c_).permute(0, 2, 1, 3).reshape(b * h, -1, c_)
        starts = [0]
        for q in querys:
            starts.append(starts[-1] + q.shape[1])
        pos_h = [p.unsqueeze(1).expand(-1, h, -1, -1).reshape(b * h, -1, 2).contiguous() for p in poss]      # level positions per head
        outputs = []
        for i, (query, pos) in enumerate(zip(querys, poss)):
            n = query.shape[1]
            offsets = self.sampling_offsets(query).view(b, n, h, l, k, 2)                            # :163
            attn = F.softmax(self.attention_weights(query).view(b, n, h, l * k), -1).view(b, n, h, l, k)    # :164-165
            idx_levels, w_levels = [], []
            for j in range(l):
                ref = scale_pos(pos, spatial_shapes[i], spatial_shapes[j], no_bias=True)             # :169
                loc = (ref[:, :, None, None, :] + offsets[:, :, :, j]).permute(0, 2, 1, 3, 4).reshape(b * h, n * k, 2)    # :174,183
                cell = scale_pos(loc, spatial_shapes[j], grid_hw, no_bias=True).round().long()       # :186-187
                gather_idx = cell[..., 0].clamp(0, grid_hw[1] - 1) + cell[..., 1].clamp(0, grid_hw[0] - 1) * grid_hw[1]
                nb = nb_idx[j].gather(index=gather_idx.view(b, -1, 1).expand(-1, -1, 4), dim=1).reshape(b * h, n * k, 4)   # :191
                nn_pos = pos_h[j].gather(index=nb.reshape(b * h, -1, 1).expand(-1, -1, 2), dim=1).reshape(b * h, n * k, 4, 2)
                dist = (loc.unsqueeze(2) - nn_pos).pow(2).sum(-1)                                    # point_utils.py:104-105 (squared)
                w_levels.append(shepard_decay_weights(dist, power=self.shepard_power))               # :194
                idx_levels.append(nb + starts[j])
            nn_idx = torch.stack(idx_levels, dim=2).reshape(b * h, n, k * l, 4)                      # :198
            nn_w = torch.stack(w_levels, dim=2).reshape(b * h, n, k * l, 4)
            a = attn.permute(0, 2, 1, 4, 3).reshape(b * h, n, k * l)                                 # :200 (point-major, level-minor)
            out = MSDETRPCFunction.apply(nn_idx, nn_w, a, val)                                       # :201
            outputs.append(self.output_proj(out.reshape(b, h, n, c_).permute(0, 2, 1, 3).reshape(b, n, c)))
        return outputs


def point2img(x, pos, mask_size=None):
    """x [b,q,n] per-point values, pos [b,n,2] integer grid positions -> [b,q,h,w] (mask2former_transformer_decoder.py:20-39).
    With ``mask_size = (h, w)`` nothing is read back to the host; ``None`` reproduces the reference's two ``.item()`` reads."""
    if x.shape[0] != pos.shape[0]:
        pos = pos.repeat(x.shape[0] // pos.shape[0], 1, 1)
    b, q, n = x.shape
    pos = pos.long()
    if mask_size is None:
        h = int(pos[:, :, 1].max().item()) + 1
        w = int(pos[:, :, 0].max().item()) + 1
    else:
        h, w = mask_size
    assert h * w == n, "h*w != n in point2img!"
    pos_idx = pos[:, :, 1] * w + pos[:, :, 0]
    ret = torch.zeros(b, q, h * w, device=x.device, dtype=x.dtype)
    ret.scatter_(index=pos_idx.unsqueeze(1).expand(-1, q, -1), dim=2, src=x)
    return ret.reshape(b, q, h, w)


def point_attn_mask(target_pos, mf_pos, outputs_mask, num_heads):
    """Decoder attention mask of one level (mask2former_transformer_decoder.py:484-486): the mask logits [b,q,n_mf] at the
    mask-feature points are interpolated to the level's points ``target_pos`` [b,n,2] (kNN-4 inverse-distance weights,
    WEIGHTEDGATHER kernel), thresholded at sigmoid < 0.5 and repeated over the heads -> bool [b*heads, q, n], detached."""
    up = upsample_feature_shepard(target_pos, mf_pos, outputs_mask.permute(0, 2, 1)).permute(0, 2, 1)
    mask = (up.sigmoid().unsqueeze(1).repeat(1, num_heads, 1, 1).flatten(0, 1) < 0.5).bool()
    return mask.detach()
