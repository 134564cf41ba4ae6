"""Point-cloud pixel-decoder pieces that sit on the CLUSTEN path (mask2former/modeling/pixel_decoder/msdeformattn_pc.py):

    PointConv   (msdeformattn_pc.py:271-314)   kNN-9 self neighbourhood -> relative-position weights -> CLUSTENWF -> LN -> Linear

Same class name, constructor arguments, parameter names (``weight_net``, ``norm``, ``linear``) and call convention
(``forward((x, pos))``) as the reference, so a reference state_dict loads unchanged.  The kNN, the weighted-feature merge
and the table lookup run on libclusten_b200; ``weight_net`` is evaluated on the table rows the batch references
(the reference evaluates it on all 1023^2 rows per call, msdeformattn_pc.py:302).
"""
import torch
from torch import nn

from .aff import REL_POS_WIDTH, TABLE_WIDTH, _TableLookup
from .ops import CLUSTENWFFunction
from .point_utils import knn_keops


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
