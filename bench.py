#!/usr/bin/env python
"""bench.py -- AFF backbone images/s on the CLUSTEN B200 path (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A step = one pass of the hot path (AFF backbone: clustering -> kNN -> QK/AV blocks -> top-k merge/WF, x4 stages) over
one batch of synthetic images.  Default workload = BASELINE configs[1]'s backbone: AFF-Mini, 512x512, batch 16 per
GPU, fp32 forward, random-init weights (the Mask2Former head is out of scope).

  value         images/s, inputs resident in HBM, CUDA events, max over ranks (one process per GPU, batch-sharded:
                no collective on the forward path; training workloads add DDP's NCCL gradient all-reduce)
  e2e           same metric through the public API with HOST buffers: pinned-host images -> H2D -> AFF.forward ->
                all outputs D2H, every step, inside the timed region
  roofline      the dominant libclusten_b200 entry point of the step, timed live with CUDA events around each of its
                launches inside the timed region; achieved = its ALGORITHMIC bytes / its device time
  cpu_baseline  the CPU oracle port of the same backbone (oracle/aff_oracle.py) on the host cores, bounded sample
  --impl reference   times that CPU port alone (the reference's Python cannot travel to the GPU box: it needs
                detectron2 / pykeops / the reference tree), rank 0 only.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: preset, per-GPU batch, H, W, mode, dtype
    "aff_mini_fwd_b16_512": dict(preset="mini", batch=16, H=512, W=512, mode="fwd", dtype="f32",
                                 note="BASELINE configs[1] backbone (AFF-Mini ADE20K, 512x512, batch 16), eval forward"),
    "aff_tiny15_train_b32_512_bf16": dict(preset="tiny_1_5", batch=32, H=512, W=512, mode="train", dtype="bf16",
                                          note="BASELINE configs[2]: AFF-Tiny-1/5 fwd+bwd+AdamW, 512x512, batch 32, bf16 autocast"),
    "aff_small_fwd_b16_512": dict(preset="small", batch=16, H=512, W=512, mode="fwd", dtype="f32",
                                  note="AFF-Small backbone forward, 512x512, batch 16 (north_star scaling target)"),
    "aff_small_fwd_b1_1024x2048": dict(preset="small", batch=1, H=1024, W=2048, mode="fwd", dtype="f32",
                                       note="BASELINE configs[3] backbone: AFF-Small Cityscapes 1024x2048, 1 image per GPU"),
    "aff_base_train_b2_512x1024_bf16": dict(preset="base", batch=2, H=512, W=1024, mode="train", dtype="bf16",
                                            note="BASELINE configs[4] backbone: AFF-Base 512x1024 crop, 2 per GPU, bf16, DDP"),
    # tiny case for the CPU tests of the host logic (never a bench line)
    "aff_test_fwd_b2_128": dict(preset="test", batch=2, H=128, W=128, mode="fwd", dtype="f32", note="unit-test workload"),
}
DEFAULT_WORKLOAD = "aff_mini_fwd_b16_512"
CPU_SAMPLE_IMAGES = 2


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons every 200 ms while the timed region runs (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        busy = [x for x in sm if x > 0.5 * (mx[0] if mx else 1e9)] or sm
        return {"sm_mhz": busy[len(busy) // 2] if busy else None, "sm_max_mhz": mx[0] if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def pin_to_gpu_numa_node(gpu_index):
    """One process per GPU: bind the process to the CPU cores of the GPU's NUMA node (NVML's ideal affinity), so that its pinned
    host buffers are allocated there and the H2D / D2H copies of the end-to-end loop do not all cross one socket (round 1: e2e
    scaling 0.935 at 8 GPUs with every rank's buffers on node 0).  Best effort: returns the number of cores bound or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cores = [64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1]
        cores = sorted(set(cores) & set(os.sched_getaffinity(0)))
        if cores:
            os.sched_setaffinity(0, cores)
            return len(cores)
    except Exception:
        pass
    return None


def make_images(batch, H, W, seed):
    import torch
    g = torch.Generator().manual_seed(seed)
    return torch.randn(batch, 3, H, W, generator=g)


def max_over_ranks(t, world):
    """Step time of the job = the slowest rank's (t: 1-element tensor on the rank's device)."""
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def whole_job_images(batch_per_gpu, world, steps):
    """Weak scaling: every rank processes its own batch_per_gpu images per step; value = this / max-over-ranks time."""
    return batch_per_gpu * world * steps


def run_reference(args, wl, rank, world):
    """CPU arm: the oracle port of the backbone on the host cores (rank 0 only)."""
    if rank != 0:
        return
    import torch
    from oracle import aff_oracle as ao
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = ao.PRESETS[wl["preset"]]
    W = ao.synthetic_state(cfg)
    train = wl["mode"] == "train"
    if train:
        W = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in W.items()}
    x = make_images(CPU_SAMPLE_IMAGES, wl["H"], wl["W"], seed=0)

    def step():
        if train:
            out = ao.aff_forward(x, W, cfg, training=False)
            sum(out[f"res{i}"].float().mean() for i in range(2, 6)).backward()
        else:
            with torch.no_grad():
                ao.aff_forward(x, W, cfg)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = CPU_SAMPLE_IMAGES * args.steps / dt
    sample = f"{CPU_SAMPLE_IMAGES} images of {wl['H']}x{wl['W']} per step, fp32, {wl['mode']}"
    emit({
        "impl": "reference", "metric": "aff_backbone_images_per_sec", "value": round(val, 4), "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * dt / args.steps, 2),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "note": wl["note"], "batch_per_step": CPU_SAMPLE_IMAGES},
        "cpu_baseline": {"value": round(val, 4), "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": round(val, 4), "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


def contract_extras(dev, log):
    """The second half of BASELINE.json's metric and the numbers SURVEY.md 8(d) asks for beside it, measured in the same process on
    the same GPU after the timed regions (N = 1 only):

      roofline_ops      the CLUSTEN ops (QK / AV / WF forward + backward, weighted gather) at AFF-Small stage 0 (B = 32, N = 16 384,
                        M = 48; merge N' = 4096), bf16 and fp32, L2 flushed between iterations: ms, achieved GB/s, fraction of the
                        measured HBM peak on the INTERFACE bytes (``frac``; the int64 index counted as delivered) and on the data
                        operands + results alone (``frac_data``)
      ref_cuda_kernels  the reference's OWN kernels (oracle/_ref, built unmodified from /root/reference for sm_100a; checker leg, like
                        cpu_baseline) on the same fp32 inputs -- the "kernel to beat"
      integer_path_us   clustering / kNN / stage preparation / merge selection / tile pack at N = 16 384 and 131 072
      head_path_us      the pixel-decoder callers of the path (configs[1]'s head half): FPN Shepard upsample, PointConv, one
                        MSDeformAttnPc layer at N0 = 16 384, conv dim 256, batch 16 (benchmarks/head_bench.py)
      north_star_model  AFF-Small backbone forward, 512x512, batch 16, fp32 (the model north_star's scaling target names)"""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "benchmarks"))
    import int_bench
    import op_bench
    from oracle import ref_cuda
    quiet = lambda *a, **k: None
    ours, refk = {}, None
    for dt in ("bf16", "f32"):
        want_ref = dt == "f32" and ref_cuda.available()
        rows = op_bench.bench_ops(op_bench.default_args(shape="small_s0", dtype=dt, iters=5, ref=want_ref), log=quiet)
        keep = ("qk_fwd", "av_fwd", "qk_bwd", "av_bwd", "wf_fwd", "wf_bwd", "wg_fwd")
        ours[dt] = {r["op"]: {k: r[k] for k in ("ms", "algo_MB", "GBs", "frac", "frac_data")} for r in rows if r["op"] in keep}
        if want_ref:
            refk = {r["op"][4:].split(" ")[0]: {"ms": r["ms"], "GBs": r["GBs"], "frac": r["frac"]} for r in rows if r["op"].startswith("REF ")}
        torch.cuda.empty_cache()
    out = {"roofline_ops": {"shape": "AFF-Small stage 0: B=32 H=3 N=16384 C=32 M=48, merge N'=4096 C=96 IC=4, WG K=4 C=256; L2 flushed",
                            "peak": op_bench.peak_gbs()[0], **ours},
           "ref_cuda_kernels": ({"kind": "reference (oracle/_ref, unmodified sources, sm_100a)", "dtype": "f32", **refk} if refk else None),
           "integer_path_us": int_bench.integer_path_us(iters=5)}
    torch.cuda.empty_cache()
    import head_bench
    out["head_path_us"] = head_bench.head_path_us(batch=16, iters=3)     # PointConv / FPN upsample / MSDeformAttnPc at N0 = 16 384
    torch.cuda.empty_cache()
    # the north-star model on the same device: graph replay, inputs resident, CUDA events
    from autofocusformermod_b200.aff import build_aff
    torch.manual_seed(0)
    wl = WORKLOADS["aff_small_fwd_b16_512"]
    m = build_aff(wl["preset"]).to(dev).eval()
    x = make_images(wl["batch"], wl["H"], wl["W"], seed=0).to(dev)
    g = m.graphed(x)
    for _ in range(3):
        g(x)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        g(x)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    out["north_star_model"] = {"workload": "aff_small_fwd_b16_512", "images_per_s": round(wl["batch"] / ms * 1e3, 1), "ms_per_step": round(ms, 3),
                               "dtype": "f32", "steps": 5, "execution": "CUDA graph replay, inputs resident in HBM"}
    return out


_REAL_STDOUT = None


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Native libraries print there too (NCCL's version banner under torchrun), so
    file descriptor 1 is pointed at stderr for the whole run and the JSON line is written to the saved descriptor."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip roofline_ops / ref_cuda_kernels / integer_path_us / north_star_model")
    ap.add_argument("--no-graph", action="store_true", help="inference workloads: eager forward instead of the CUDA-graph replay")
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch of the workload (diagnostics; the line says so)")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.batch > 0:
        wl["batch"] = args.batch
        wl["note"] += f" [batch overridden to {args.batch} per GPU]"
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, wl, rank, world)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = pin_to_gpu_numa_node(local_rank) if world > 1 else None      # before any pinned allocation: first touch decides the node
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on STDOUT at NCCL_DEBUG=VERSION: keep stdout to the one JSON line of the contract
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    from autofocusformermod_b200 import ops
    from autofocusformermod_b200.aff import build_aff

    torch.manual_seed(0)
    train = wl["mode"] == "train"
    model = build_aff(wl["preset"]).to(dev)
    model.train(train)
    amp = wl["dtype"] == "bf16"
    opt = None
    net = model
    sync_grads = None
    if train and hasattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch"):
        # (the graphed backward runs on the capture's side stream; AccumulateGrad nodes live on the default stream -- by design here)
        torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
    if train:
        if world > 1 and args.no_graph:
            net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local_rank], gradient_as_bucket_view=True)
        elif world > 1:
            # CUDA-graph training under data parallelism: same initial weights on every rank (torch.manual_seed(0) above), the backward
            # graph replays as on one GPU, then ONE flat NCCL all-reduce of the gradients (aff.FlatGradAllReduce)
            from autofocusformermod_b200.aff import FlatGradAllReduce
            sync_grads = FlatGradAllReduce(model.parameters())
        opt = torch.optim.AdamW(model.parameters(), lr=1e-4, fused=True)

    B = wl["batch"]
    x_host = make_images(B, wl["H"], wl["W"], seed=rank).pin_memory()
    x_dev = x_host.to(dev)

    # inference workloads go through the public CUDA-graph API (AFF.graphed): one replay per batch, no launch overhead
    graphed = None
    if not train and not args.no_graph:
        graphed = model.graphed(x_dev, autocast_dtype=torch.bfloat16 if amp else None)

    graphed_train = None
    if train and not args.no_graph:
        from autofocusformermod_b200.aff import graphed_training_forward
        graphed_train = graphed_training_forward(model, x_dev, autocast_dtype=torch.bfloat16 if amp else None)

    def eager_step(x):
        if train:
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
                out = net(x)
                loss = sum(out[f"res{i}"].float().mean() for i in range(2, 6))
            opt.zero_grad(set_to_none=True)
            loss.backward()
            if sync_grads is not None:
                sync_grads()
            opt.step()
            return {"loss": loss.detach()}
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
            return net(x)

    def step(x):
        if graphed is not None:
            return graphed(x)
        if graphed_train is not None:                  # forward and backward are one graph replay each; AdamW stays eager
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp, cache_enabled=False):
                feats = graphed_train(x)
                loss = sum(f.float().mean() for f in feats)
            opt.zero_grad(set_to_none=True)
            loss.backward()
            if sync_grads is not None:
                sync_grads()
            opt.step()
            return {"loss": loss.detach()}
        return eager_step(x)

    use_graph = graphed is not None or graphed_train is not None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        out = step(x_dev)
    barrier()
    # untimed pre-pass: which of our entry points dominates a step?
    ops.start_kernel_timer("*")
    eager_step(x_dev)
    per = {}
    for name, ms, nb in ops.stop_kernel_timer():
        e = per.setdefault(name, [0.0, 0, 0])
        e[0] += ms
        e[1] += nb
        e[2] += 1
    dominant = max((n for n in per if per[n][1] > 0), key=lambda n: per[n][0])
    torch.cuda.reset_peak_memory_stats()

    # ---- timed region 1: inputs resident in HBM --------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    k0 = ops.kernel_launches()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dom_flops = (0.0, 0.0)
    f0 = ops.flops_issued(dominant)
    if not use_graph:
        ops.start_kernel_timer(dominant)
    ev0.record()
    for _ in range(args.steps):
        step(x_dev)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    if not use_graph:
        dom = ops.stop_kernel_timer()
        dom_flops = tuple(a_ - b_ for a_, b_ in zip(ops.flops_issued(dominant), f0))
        launches = ops.kernel_launches() - k0
        roof_timing = "CUDA events around every launch of the entry point inside the timed region"
    else:
        # a graph replay cannot carry per-launch events: the dominant entry point is timed over the same number of EAGER
        # steps on the same inputs right after the timed region (same kernels, same arguments)
        k1 = ops.kernel_launches()
        f0 = ops.flops_issued(dominant)
        ops.start_kernel_timer(dominant)
        for _ in range(args.steps):
            eager_step(x_dev)
        dom = ops.stop_kernel_timer()
        dom_flops = tuple(a_ - b_ for a_, b_ in zip(ops.flops_issued(dominant), f0))
        # kernels of ours inside the replays of the timed region = what the same steps launch eagerly
        launches = graphed.launches_per_replay * args.steps if graphed is not None else ops.kernel_launches() - k1
        roof_timing = "CUDA events around every launch of the entry point over the same number of eager steps after the timed region (graph replays carry no per-launch events)"
    clocks = sampler.stop()
    peak_mem = torch.cuda.max_memory_allocated()
    ms_max = max_over_ranks(torch.tensor([ms], device=dev, dtype=torch.float64), world)

    # ---- timed region 2: end to end with host buffers --------------------------------------------------------------
    out = step(x_dev)
    host_out = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in out.items() if torch.is_tensor(v)}
    d2h = sum(v.numel() * v.element_size() for v in host_out.values())
    h2d = x_host.numel() * x_host.element_size()
    x_stage = torch.empty_like(x_dev)

    def e2e_step():
        if graphed is not None:
            o = graphed(x_host)                # pinned host batch -> the graph's input buffer (H2D) -> replay
        else:
            x_stage.copy_(x_host, non_blocking=True)
            o = step(x_stage)
        for k, buf in host_out.items():
            buf.copy_(o[k], non_blocking=True)
        torch.cuda.synchronize()          # the step's result is on the host

    if graphed is not None:
        # serving loop of the public API: H2D of batch i+1 and D2H of result i-1 overlap the replay of batch i; every step's
        # input still comes from pinned host memory and every step's outputs still land in pinned host memory
        sinks = [host_out, {k: torch.empty_like(v).pin_memory() for k, v in host_out.items()}]
        for _ in graphed.stream((x_host for _ in range(3)), sinks):
            pass
        barrier()
        t0 = time.perf_counter()
        done = sum(1 for _ in graphed.stream((x_host for _ in range(args.steps)), sinks))
        assert done == args.steps
        barrier()
        e2e_s = time.perf_counter() - t0
    else:
        for _ in range(2):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        barrier()
        e2e_s = time.perf_counter() - t0
    e2e_max = max_over_ranks(torch.tensor([e2e_s], device=dev, dtype=torch.float64), world)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (sustained copy)"
    else:
        peak, peak_src = 6650.0, "fallback of B200_PROFILING.md"
    dom_ms = sum(m for _, m, _ in dom)
    dom_bytes = sum(nb for _, _, nb in dom)
    achieved = dom_bytes / dom_ms / 1e6 if dom_ms > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")          # filled from an `ncu --set full` capture (per launch)
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(args.workload, {}).get(dominant)
    roofline = {"bound": "hbm", "kernel": dominant, "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_src,
                "launches_timed": len(dom), "algorithmic_bytes_per_launch_avg": int(dom_bytes / max(len(dom), 1)),
                "avg_launch_ms": round(dom_ms / max(len(dom), 1), 4),
                "share_of_step": round(dom_ms / ms, 4), "timing": roof_timing,
                "per_entry_ms_per_step": {n: round(v[0], 4) for n, v in sorted(per.items())}}

    # every entry point of the untimed eager pre-pass against the HBM roofline (one step; the section 8(a) kernels stay visible when
    # a Linear dominates the step)
    roofline["by_entry"] = {n: {"ms": round(v[0], 4), "gbps": round(v[1] / v[0] / 1e6, 1), "frac": round(v[1] / v[0] / 1e6 / peak, 4)}
                            for n, v in sorted(per.items()) if v[1] > 0 and v[0] > 0}
    if dom_flops[0] and dom_ms > 0:
        # GEMM-shaped dominant kernel: also against the tensor-core ceiling.  An fp32 product is three MMAs (hi/lo split of both
        # operands): fp16 halves run at the bf16 / fp16 dense rate, TF32 halves at half of it; `executed` counts them in bf16-rate
        # equivalents, so its ceiling is the measured bf16 throughput.
        bf16 = float(json.load(open(peaks_path)).get("bf16_tflops_sustained", 1404.1)) if os.path.exists(peaks_path) else 1404.1
        ex, alg = dom_flops[0] / dom_ms / 1e9, dom_flops[1] / dom_ms / 1e9
        roofline["tensor"] = {"achieved": round(ex, 1), "peak": bf16, "unit": "TFLOP/s executed on the tensor cores, bf16-rate equivalents",
                              "frac": round(ex / bf16, 4), "algorithmic_fp32_tflops": round(alg, 1),
                              "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained"}
        if roofline["tensor"]["frac"] > roofline["frac"]:
            roofline["note"] = ("the launches of this entry point range from HBM-bound (K = 32) to tensor-bound (K >= 256); both "
                                "ceilings are reported, the HBM one in the contract keys")

    execution = ("CUDA graph replay (AFF.graphed)" if graphed is not None else
                 ("CUDA graphs for forward and backward (graphed_training_forward), eager AdamW"
                  + (f", one flat NCCL all-reduce of {sync_grads.nbytes / 2**20:.0f} MiB of fp32 gradients per step" if sync_grads is not None else ""))
                 if graphed_train is not None else "eager")
    extras = {}
    if world == 1 and not args.no_extras and args.workload == DEFAULT_WORKLOAD:
        graphed = graphed_train = out = None              # release the graphs' memory pools before the op benchmarks
        torch.cuda.empty_cache()
        extras = contract_extras(dev, None)

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import aff_oracle as ao
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        cfg = ao.PRESETS[wl["preset"]]
        Wc = ao.synthetic_state(cfg)
        xs = make_images(CPU_SAMPLE_IMAGES, wl["H"], wl["W"], seed=0)
        if train:
            Wc = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in Wc.items()}
        reps, t0 = 0, time.perf_counter()
        while reps < 2 or (time.perf_counter() - t0 < 10.0 and reps < 8):
            if train:
                o = ao.aff_forward(xs, Wc, cfg)
                sum(o[f"res{i}"].float().mean() for i in range(2, 6)).backward()
            else:
                with torch.no_grad():
                    ao.aff_forward(xs, Wc, cfg)
            reps += 1
        dt = time.perf_counter() - t0
        cpu = {"value": round(CPU_SAMPLE_IMAGES * reps / dt, 4), "unit": "images/s", "cores": cores, "kind": "port",
               "sample": f"{reps} passes of {CPU_SAMPLE_IMAGES} images {wl['H']}x{wl['W']}, fp32, {wl['mode']} (oracle/aff_oracle.py)"}

    total_images = whole_job_images(B, world, args.steps)
    line = {
        "metric": "aff_backbone_images_per_sec", "value": round(total_images / (ms_max / 1e3), 2), "unit": "images/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_max / args.steps, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": wl["dtype"], "data": "synthetic",
        "config": {"workload": args.workload, "note": wl["note"], "batch_per_gpu": B, "global_batch": B * world,
                   "image": [wl["H"], wl["W"]], "host_cores_bound": numa, "parallelism": f"dp{world} (batch-sharded, no data-path collective"
                   + (", NCCL gradient all-reduce)" if train and world > 1 else ")"),
                   "execution": execution,
                   "l2": f"no explicit flush: one step streams {peak_mem / 2**20:.0f} MiB of live activations (>> 126 MB L2)",
                   "memoised": ("the position-only structures of the on-grid stage 0 (clustering, kNN, neighbourhoods, tile pack) are constants "
                                "of the input shape and are built once, outside the timed region (the reference memoises its stage-0 "
                                "clustering the same way in training, aff.py:461-467); their cost is in integer_path_us"),
                   # environment switches of the Python layer that differ from their defaults (README.md), so a line says what it ran
                   "switches": {k: v for k, v in sorted(os.environ.items()) if k.startswith("CLUSTEN_")}},
        "e2e": {"value": round(total_images / e2e_max, 2), "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
        **extras,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
