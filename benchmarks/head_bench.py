"""The point-cloud pixel-decoder pieces of BASELINE configs[1] that sit on the CLUSTEN path, at the res2 scale of a 512x512 image
(N0 = 16 384 stem-grid tokens, conv dim 256, fp32 -- the decoder runs under autocast(enabled=False), msdeformattn_pc.py:464):

  fpn_upsample   upsample_feature_shepard(pos, last_pos, out[-1])  (msdeformattn_pc.py:527): kNN-4 of the 16 384 tokens among the
                 4096 tokens of res3 + inverse-distance weights + WEIGHTEDGATHER
  point_conv     PointConv((y, pos))  (msdeformattn_pc.py:528 -> 285-314): self kNN-9 (16 384^2 distance evaluations per image),
                 relative-position weight table, CLUSTENWF (M = 9, IC = 4), LayerNorm, Linear
  msdeform_attn  one MSDeformAttnPc layer over res5 / res4 / res3 (msdeformattn_pc.py:143-205): 3 levels x 4 points x 8 heads,
                 lookup-table gathers + MSDETRPC

Times are CUDA-event medians in microseconds per call at the given batch; kNN also as distance evaluations per second.

    python benchmarks/head_bench.py [--batch 16] [--iters 5]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def _time(fn, iters):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return round(ts[len(ts) // 2], 1)


def head_path_us(batch=16, iters=5, h=128, w=128, dim=256):
    from _inputs import grid_positions, random_positions
    from autofocusformermod_b200 import point_utils as pu
    from autofocusformermod_b200.pixel_decoder import MSDeformAttnPc, PointConv, grid_lookup_tables
    B, n0 = batch, h * w
    g = torch.Generator(device="cuda").manual_seed(0)
    pos = grid_positions(B, h, w).cuda()
    ns = [n0 // 64, n0 // 16, n0 // 4]                                     # res5, res4, res3 token counts (ds 0.25)
    poss = [random_positions(B, n, h, w, seed=3 + i).cuda() for i, n in enumerate(ns)]
    last_pos, last = poss[-1], torch.randn(B, ns[-1], dim, device="cuda", generator=g)
    y = torch.randn(B, n0, dim, device="cuda", generator=g)
    conv = PointConv(dim, dim, bias=False).cuda().eval()
    attn = MSDeformAttnPc(dim, 3, 8, 4, 4.0, True).cuda().eval()
    ss = [(h, w)] * 4
    srcs = [torch.randn(B, n, dim, device="cuda", generator=g) for n in ns]
    out = {"batch": B, "n0": n0, "dim": dim}
    with torch.no_grad():
        nb_idx = grid_lookup_tables(poss, ss[:-1], (h, w))
        out["knn4_upsample"] = _time(lambda: pu.knn_keops(pos, last_pos, 4), iters)
        out["fpn_upsample"] = _time(lambda: pu.upsample_feature_shepard(pos, last_pos, last), iters)
        out["knn9_self"] = _time(lambda: pu.knn_keops(pos, pos, 9), iters)
        out["knn9_gdist_per_s"] = round(B * n0 * n0 / (out["knn9_self"] * 1e-6) / 1e9, 1)
        out["point_conv"] = _time(lambda: conv((y, pos)), iters)
        out["lookup_tables_3_levels"] = _time(lambda: grid_lookup_tables(poss, ss[:-1], (h, w)), iters)
        out["msdeform_attn_layer"] = _time(lambda: attn(srcs, poss, srcs, ss, nb_idx), iters)
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--iters", type=int, default=5)
    a = ap.parse_args()
    print(json.dumps(head_path_us(a.batch, a.iters)))
