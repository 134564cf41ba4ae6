"""Cycle counters of the tcgen05 Linear (issuer warp + first epilogue warp), from a library built with -DCLUSTEN_TC_PROFILE:
    (tools/r2_lin_prof.sh says how autofocusformermod_b200/libclusten_b200_prof.so is built; the shipped library has no counters)
usage: python tools/lin_profile.py R K N [epilogue] [ln]"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autofocusformermod_b200 import _lib  # noqa: E402
_lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), "libclusten_b200_prof.so")      # the -DCLUSTEN_TC_PROFILE build
from autofocusformermod_b200 import ops  # noqa: E402

R, K, N = (int(v) for v in sys.argv[1:4])
epi = sys.argv[4] if len(sys.argv) > 4 else "bias"
use_ln = len(sys.argv) > 5 and sys.argv[5] == "ln"
g = torch.Generator().manual_seed(0)
x = torch.randn(R, K, generator=g).cuda()
w = (torch.randn(N, K, generator=g) * K ** -0.5).cuda()
b = torch.randn(N, generator=g).cuda()
res = torch.randn(R, N, generator=g).cuda()
lw, lb = torch.ones(K).cuda(), torch.zeros(K).cuda()
ln = ops.layer_norm_stats(x, lw, lb, 1e-5) + (lw, lb) if use_ln else None
for _ in range(3):
    y = ops.linear_tc(x, w, b, epi, res=res, ln=ln, split="f16")
torch.cuda.synchronize()
a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
y = ops.linear_tc(x, w, b, epi, res=res, ln=ln, split="f16")
e.record()
torch.cuda.synchronize()
# per-launch time inside a CUDA graph of 20 back-to-back launches (what the model's graph replay pays per Linear)
gr = torch.cuda.CUDAGraph()
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    ops.linear_tc(x, w, b, epi, res=res, ln=ln, split="f16")
    with torch.cuda.graph(gr, stream=st):
        for _ in range(20):
            y = ops.linear_tc(x, w, b, epi, res=res, ln=ln, split="f16")
torch.cuda.synchronize()
gr.replay()
torch.cuda.synchronize()
a2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a2.record()
gr.replay()
e2.record()
torch.cuda.synchronize()
graph_us = a2.elapsed_time(e2) * 1e3 / 20
buf = (ctypes.c_longlong * (148 * 16))()
L = _lib.lib()
L.clusten_linear_tc_profile.argtypes = [ctypes.c_void_p]
assert L.clusten_linear_tc_profile(buf) == 0
t = torch.tensor(list(buf), dtype=torch.float64).view(148, 16)
act = t[:, 4] > 0
n = int(act.sum())
m = t[act].mean(0)
ch = m[4]
mx = t[act].max(0).values
print(f"R={R} K={K} N={N} {epi} ln={use_ln}: {graph_us:.1f} us per launch in a graph of 20; {n} CTAs, {ch:.1f} issue rounds per CTA; slowest CTA: issuer {mx[5] / 1965:.1f} us, epilogue {mx[13] / 1965:.1f} us (cycles / 1965 MHz)")
print(f"  issuer  per chunk: wait drained acc {m[0] / ch:7.0f}  wait W {m[1] / ch:7.0f}  wait A {m[2] / ch:7.0f}  issue {m[3] / ch:7.0f}   role total {m[5] / ch:7.0f} cycles / chunk ({m[5]:.0f} cycles)")
print(f"  epilogue per chunk: wait acc {m[8] / ch:7.0f}  drain {m[9] / ch:7.0f}  wait staging+bar {m[10] / ch:7.0f}  math+STS {m[11] / ch:7.0f}  bar {m[12] / ch:7.0f}   role total {m[13] / ch:7.0f} cycles / chunk")
