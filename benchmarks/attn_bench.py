"""Fused ClusterAttention core (clusten_attn_fwd / clusten_attn_bwd / clusten_scatter_rows / clusten_table_grad) at one AFF
stage shape: per-entry device time, algorithmic bytes, achieved GB/s.  Index tensors come from the product's own stage pipeline
(clustering + kNN + stage_prepare on a random token subset), so padded last clusters (n % m != 0) show their real cost.

    python benchmarks/attn_bench.py --n 655 --heads 8 --batch 32 [--m 8 --nbhd 48 --c 32 --dtype bf16|f16|f32 --iters 10 --inkernel-bias]   (f32: inference forward only)
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=655)
    ap.add_argument("--heads", type=int, default=8)
    ap.add_argument("--c", type=int, default=32)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--m", type=int, default=8)
    ap.add_argument("--nbhd", type=int, default=48)
    ap.add_argument("--grid", type=int, default=128)
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--fwd-only", action="store_true")
    ap.add_argument("--inkernel-bias", action="store_true", help="the clusten_attn_pos_* variant: bias computed from positions in the kernels")
    args = ap.parse_args()
    from autofocusformermod_b200 import ops
    from _inputs import stage_structure
    dt = {"bf16": torch.bfloat16, "f16": torch.float16, "f32": torch.float32}[args.dtype]
    if dt == torch.float32:
        args.fwd_only = True                              # fp32 runs the fused kernel on the inference path only
    B, N, H, C = args.batch, args.n, args.heads, args.c
    pos, nb, mask, uniq, inv = stage_structure(1, N, args.grid, args.grid, args.m, args.nbhd, seed=0)
    pos = pos.expand(B, -1, -1).contiguous()
    M = nb.shape[-1]
    idx = nb.expand(B, -1, -1).contiguous()
    mask8 = None if mask is None else mask.expand(B, -1, -1).contiguous()
    bias_idx = inv.expand(B, -1, -1).contiguous()
    g = torch.Generator(device="cuda").manual_seed(0)
    q = (torch.randn(B, N, H, C, device="cuda", generator=g) * C ** -0.5).to(dt).requires_grad_(True)
    kv = torch.randn(B, N, H, 2, C, device="cuda", generator=g).to(dt).requires_grad_(True)
    tab = torch.randn(uniq.numel(), H, device="cuda", generator=g).requires_grad_(True)
    pe_w = (torch.randn(H, 5, device="cuda", generator=g) * 0.2).requires_grad_(True)
    pe_b = torch.randn(H, device="cuda", generator=g).requires_grad_(True)
    bk = torch.randn(H * C, device="cuda", generator=g).to(dt).requires_grad_(True)
    bv = torch.randn(H * C, device="cuda", generator=g).to(dt).requires_grad_(True)
    go = torch.randn(B, N, H * C, device="cuda", generator=g).to(dt)
    flush = torch.zeros(256 << 20, dtype=torch.uint8, device="cuda")
    print(json.dumps({"shape": dict(B=B, N=N, H=H, C=C, M=M, U=int(uniq.numel())), "pack_flags(generic,maxU,impure,over)": ops.pack_flags(idx, N), "masked_pack_flags": ops.pack_flags(idx, N, mask=mask8)}), flush=True)
    per = {}
    for it in range(args.iters + 3):
        flush.add_(1)
        if it >= 3:
            ops.start_kernel_timer("*")
        if dt == torch.float32:
            with torch.no_grad():
                kvp = kv.permute(3, 0, 2, 1, 4)
                if args.inkernel_bias:
                    ops.cluster_attention_fused_pos(q.permute(0, 2, 1, 3), kvp[0], kvp[1], idx, pos, pe_w, pe_b, mask8, bk, bv)
                else:
                    ops.cluster_attention_fused(q.permute(0, 2, 1, 3), kvp[0], kvp[1], idx, tab, bias_idx, mask8, bk, bv)
        else:
            if args.inkernel_bias:
                out = ops.cluster_attention_core_pos(q, kv, pe_w, pe_b, bk, bv, idx, pos, mask8)
            else:
                out = ops.cluster_attention_core(q, kv, tab, bk, bv, idx, bias_idx, mask8)
            if not args.fwd_only:
                out.backward(go)
        if it >= 3:
            for name, ms, nbytes in ops.stop_kernel_timer():
                e = per.setdefault(name, [0.0, 0, 0])
                e[0] += ms; e[1] += nbytes; e[2] += 1
    peak = 6452.5
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    for name, (ms, nbytes, cnt) in sorted(per.items()):
        gbs = nbytes / ms / 1e6 if ms > 0 else 0
        print(json.dumps({"entry": name, "calls_per_iter": cnt / args.iters, "ms_per_iter": round(ms / args.iters, 4),
                          "algo_MB_per_iter": round(nbytes / args.iters / 1e6, 1), "GBs": round(gbs, 1), "frac": round(gbs / peak, 3)}), flush=True)


if __name__ == "__main__":
    main()
