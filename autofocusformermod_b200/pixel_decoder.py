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
import torch
from torch import nn

from .aff import REL_POS_WIDTH, TABLE_WIDTH, _TableLookup
from .ops import CLUSTENWFFunction
from .point_utils import knn_keops, upsample_feature_shepard


class PointConv(nn.Module):
    def __init__(self, dim, out_dim, bias):
        super().__init__()
        inner_ch = 4
        self.weight_net = nn.Sequential(nn.Linear(5, inner_ch, bias=True), nn.LayerNorm(inner_ch), nn.GELU())
        self.norm = nn.LayerNorm(inner_ch * dim)
        self.linear = nn.Linear(dim * inner_ch, out_dim, bias=bias)

    def forward(self, inp):
        """inp = (x [b,n,c] point features, pos [b,n,2] point positions) -> [b,n,out_dim]"""
        x, pos = inp
        b, n, c = x.shape
        nn_idx = knn_keops(pos, pos, 9)                                                        # msdeformattn_pc.py:295
        nn_pos = pos.gather(index=nn_idx.view(b, -1, 1).expand(-1, -1, 2), dim=1).reshape(b, n, 9, 2)
        rel_pos = pos.unsqueeze(2) - nn_pos                                                    # :297 (query minus neighbour)
        rel = (rel_pos.long() + REL_POS_WIDTH).clamp(0, TABLE_WIDTH - 1)                       # :305
        pe_idx = rel[..., 1] * TABLE_WIDTH + rel[..., 0]                                       # :306
        weights = _TableLookup(pe_idx)(self.weight_net)                                        # :302-308 on the referenced rows
        feat = CLUSTENWFFunction.apply(weights, x, nn_idx).reshape(b, n, -1)                   # :309
        return self.linear(self.norm(feat))                                                    # :311-313


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
