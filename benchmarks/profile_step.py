"""torch.profiler view of one bench workload: GPU time per kernel over a few steps (eager or graphed), plus the GPU-busy share
of the wall-clock step.  usage: python benchmarks/profile_step.py [--workload NAME] [--steps 3] [--no-graph]"""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="aff_tiny15_train_b32_512_bf16")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--rows", type=int, default=45)
    args = ap.parse_args()
    import bench
    from autofocusformermod_b200.aff import build_aff, graphed_training_forward
    wl = bench.WORKLOADS[args.workload]
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    train = wl["mode"] == "train"
    model = build_aff(wl["preset"]).to(dev)
    model.train(train)
    amp = wl["dtype"] == "bf16"
    x = bench.make_images(wl["batch"], wl["H"], wl["W"], 0).to(dev)
    if train:
        opt = torch.optim.AdamW(model.parameters(), lr=1e-4, fused=True)
        f = None if args.no_graph else graphed_training_forward(model, x, torch.bfloat16 if amp else None)

        def step():
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp, cache_enabled=f is None):
                if f is not None:
                    feats = f(x)
                else:
                    out = model(x)
                    feats = [out[f"res{i}"] for i in range(2, 6)]
                loss = sum(t.float().mean() for t in feats)
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
    else:
        g = None if args.no_graph else model.graphed(x, torch.bfloat16 if amp else None)

        def step():
            if g is not None:
                return g(x)
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
                return model(x)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        for _ in range(args.steps):
            step()
        torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / args.steps * 1e3
    ev = prof.key_averages()
    tot = sum(e.device_time_total for e in ev) / args.steps / 1e3
    print(f"wall {wall:.2f} ms/step (under the profiler), GPU kernel time {tot:.2f} ms/step")
    rows = sorted(ev, key=lambda e: -e.device_time_total)[:args.rows]
    for e in rows:
        print(f"{e.device_time_total / args.steps / 1e3:8.3f} ms {e.count // args.steps:5d}x  {e.key[:110]}")


if __name__ == "__main__":
    main()
