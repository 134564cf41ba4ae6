"""fp32 Linear layers of the backbone (aff.py:62-70,103-106,181-189): clusten_linear_tc_f32 (tcgen05, 3xTF32) against cuBLAS
(``F.linear`` with TF32 off = what the reference runs) -- accuracy against a float64 product and time per call.

    python benchmarks/linear_bench.py [--model mini|small] [--batch 16] [--chain 0] [--iters 10]
"""
import argparse
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SHAPES = {   # (tokens per image at 512^2, C, mlp ratio) per stage
    "mini": [(16384, 32, 2), (4096, 128, 2), (1024, 256, 2), (256, 384, 2)],
    "small": [(16384, 96, 3), (4096, 192, 3), (1024, 384, 3), (256, 768, 3)],
}


def _time(fn, iters):
    """Median device time of one call, replayed from a CUDA graph (how the model runs it; no host time between the events)."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        fn()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        graph.replay()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def rel(a, b):
    return float((a.double() - b).abs().max() / b.abs().max())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="mini")
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--chain", type=int, default=0)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--check-only", action="store_true")
    ap.add_argument("--split", default=None, help="f16 | tf32 (default: the library default, CLUSTEN_TC_SPLIT)")
    ap.add_argument("--only", default="", help="comma-separated stage:layer picks, e.g. 0:proj,1:q+kv")
    a = ap.parse_args()
    from autofocusformermod_b200 import ops
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(0)
    rows = []
    only = set(a.only.split(",")) if a.only else None
    for si, (n, C, mlp) in enumerate(SHAPES[a.model]):
        R = a.batch * n
        for name, K, N, epi in (("q+kv", C, 3 * C, "bias"), ("proj", C, C, "residual"), ("fc1", C, mlp * C, "gelu"), ("fc2", mlp * C, C, "residual")):
            if only is not None and f"{si}:{name}" not in only:
                continue
            x = torch.randn(R, K, device="cuda", generator=g)
            w = torch.randn(N, K, device="cuda", generator=g) * K ** -0.5
            b = torch.randn(N, device="cuda", generator=g)
            res = torch.randn(R, N, device="cuda", generator=g)
            gam = torch.rand(N, device="cuda", generator=g) + 0.5

            def ref_fn():
                y = F.linear(x, w, b)
                if epi == "gelu":
                    return F.gelu(y)
                if epi == "residual":
                    return res + gam * y
                y[:, :C] *= 0.25
                return y

            def our_fn():
                return ops.linear_tc(x, w, b, epi, res=res, gamma=gam, alpha=0.25, alpha_cols=C, chain=a.chain, split=a.split)

            sub = slice(0, min(R, 8192))
            y64 = x[sub].double() @ w.double().t() + b.double()
            if epi == "gelu":
                y64 = F.gelu(y64)
            elif epi == "residual":
                y64 = res[sub].double() + gam.double() * y64
            else:
                y64[:, :C] *= 0.25
            yo, yr = our_fn(), ref_fn()
            torch.cuda.synchronize()
            row = {"layer": name, "R": R, "K": K, "N": N, "err_tc": rel(yo[sub], y64), "err_cublas": rel(yr[sub], y64),
                   "tail_equal": bool(torch.allclose(yo[-300:], yr[-300:], rtol=1e-4, atol=1e-4))}
            if not a.check_only:
                row["us_tc"] = round(_time(our_fn, a.iters), 1)
                row["us_cublas"] = round(_time(ref_fn, a.iters), 1)
                row["us_cublas_gemm_only"] = round(_time(lambda: F.linear(x, w, b), a.iters), 1)
                bytes_ = 4 * (R * (K + N * (2 if epi == "residual" else 1)) + 2 * N * K)
                row["gbps_tc"] = round(bytes_ / row["us_tc"] / 1e3, 1)
                row["tflops_tc"] = round(2 * R * K * N / row["us_tc"] / 1e6, 1)
            print(json.dumps(row), flush=True)
            rows.append(row)
    if not a.check_only:
        print(json.dumps({"model": a.model, "batch": a.batch, "chain": a.chain, "sum_us_tc": round(sum(r["us_tc"] for r in rows), 1),
                          "sum_us_cublas": round(sum(r["us_cublas"] for r in rows), 1)}))


if __name__ == "__main__":
    main()
