"""Debug aid: feature errors of the real-preset goldens under a few switches (run on the GPU box)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import aff_oracle as ao
from conftest import GOLDEN, rel_err
from autofocusformermod_b200.aff import build_aff

def run(name, preset, B, H, W, tf32):
    torch.backends.cudnn.allow_tf32 = tf32
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    m = build_aff(preset); m.load_state_dict(ao.synthetic_state(ao.PRESETS[preset]), strict=False); m = m.cuda().eval()
    x = ao.synthetic_images(B, H, W).cuda()
    with torch.no_grad():
        out = m(x)
    r = {}
    for i in range(2, 6):
        s = int(g[f"res{i}_stride"])
        ok = torch.equal(out[f"res{i}_pos"].cpu().to(torch.int16), torch.from_numpy(g[f"res{i}_pos"]))
        r[f"res{i}"] = (ok, f"{rel_err(out[f'res{i}'][:, ::s], torch.from_numpy(g[f'res{i}_sub'])):.2e}" if ok else "-")
    print(name, "cudnn_tf32", tf32, "TMA", os.environ.get("CLUSTEN_TMA_ATTN", "1"), r, flush=True)

for tf32 in (True, False):
    run("aff_mini_512", "mini", 2, 512, 512, tf32)
    run("aff_base_256x512", "base", 1, 256, 512, tf32)
    run("aff_tiny_1_5_512", "tiny_1_5", 2, 512, 512, tf32)
