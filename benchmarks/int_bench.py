"""Integer path of one AFF stage in microseconds (clustering sort / kNN / top-k are launch- and latency-bound at
these sizes, so they are reported as times, and kNN also as distance evaluations per second):

  space_filling_cluster (point_utils.py:135-287)      clusten_sfc_cluster
  knn_keops(tokens -> cluster centres, k = 6) (aff.py:475)   clusten_knn
  stage_prepare (aff.py:478-485)                       clusten_stage_prepare
  merge selection (aff.py:292-329)                     clusten_topk_select + clusten_mask_select
  tile pack (per index tensor)                         clusten_pack_build

on the stem grid of a 512x512 image (N = 16 384, per-GPU batch 16) and of a 1024x2048 image (N = 131 072, batch 1).

    python benchmarks/int_bench.py [--iters 5]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = {"n16384_b16": (16, 128, 128), "n131072_b1": (1, 256, 512)}


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


def integer_path_us(iters=5, m=8, nbhd=48, ds_rate=0.25):
    from autofocusformermod_b200 import ops
    from autofocusformermod_b200 import point_utils as pu
    out = {}
    for name, (B, h, w) in CASES.items():
        n = h * w
        ys, xs = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
        pos = torch.stack([xs, ys], dim=2).reshape(1, -1, 2).float().expand(B, -1, -1).contiguous().cuda()
        spos, mean_pos, member, cmask, _ = pu.space_filling_cluster(pos, m, h, w)
        nnc = nbhd // m
        nearest = pu.knn_keops(spos, mean_pos, nnc)
        prepared = pu.stage_prepare(spos, nearest, member, cmask, extent=(h, w))
        idx = prepared[0]
        score = torch.rand(B, n, device="cuda")
        reserve = ((spos.long() % 4) == 0).all(-1).float()
        keep, rnum = int(n * ds_rate), ((h + 3) // 4) * ((w + 3) // 4)
        score = score + reserve * (-100)
        r = {
            "sfc_cluster": _time(lambda: pu.space_filling_cluster(pos, m, h, w), iters),
            "knn_k6": _time(lambda: pu.knn_keops(spos, mean_pos, nnc), iters),
            "stage_prepare": _time(lambda: pu.stage_prepare(spos, nearest, member, cmask, extent=(h, w)), iters),
            "merge_select": _time(lambda: pu.merge_select(score, reserve, keep, rnum), iters),
        }

        def pack():
            if hasattr(idx, "_clusten_pack"):
                del idx._clusten_pack
            ops.neighbourhood_pack(idx, n)

        r["pack_build"] = _time(pack, iters)
        r["knn_gdist_per_s"] = round(B * n * mean_pos.shape[1] / (r["knn_k6"] * 1e-6) / 1e9, 1)
        out[name] = r
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=5)
    print(json.dumps(integer_path_us(ap.parse_args().iters)))
