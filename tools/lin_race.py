"""Characterise wrong output blocks of clusten_linear_tc_f32: which 128 x 32 pieces differ, and what they contain instead."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from autofocusformermod_b200 import ops

torch.backends.cuda.matmul.allow_tf32 = False
R, K, N = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (16384, 256, 768)))
epi = sys.argv[4] if len(sys.argv) > 4 else "bias"
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(R, K, device="cuda", generator=g)
w = torch.randn(N, K, device="cuda", generator=g) * K ** -0.5
b = torch.randn(N, device="cuda", generator=g)
res = torch.randn(R, N, device="cuda", generator=g)
ref = x @ w.t() + b
if epi == "gelu":
    ref = torch.nn.functional.gelu(ref)
elif epi == "residual":
    ref = res + ref
# partial products per 128-wide K chain, to recognise "a chain is missing / doubled"
chains = [x[:, k:k + 128] @ w[:, k:k + 128].t() for k in range(0, K, 128)]
for trial in range(int(os.environ.get("TRIALS", "6"))):
    y = ops.linear_tc(x, w, b, epi, res=res)
    torch.cuda.synchronize()
    bad = ((y - ref).abs() > 1e-3).view(R // 128, 128, N // 32, 32).any(3).any(1)          # [row tiles, 32-col pieces]
    nbad = int(bad.sum())
    print(f"trial {trial}: {nbad} wrong pieces of {bad.numel()}", flush=True)
    if nbad:
        idx = bad.nonzero()
        print("  first wrong (row tile, piece):", idx[:12].tolist(), " row tiles:", sorted(set(idx[:, 0].tolist()))[:20])
        mt, pc = idx[0].tolist()
        blk = (slice(mt * 128, mt * 128 + 128), slice(pc * 32, pc * 32 + 32))
        d = y[blk] - ref[blk]
        rows_bad = ((d.abs() > 1e-3).any(1)).nonzero().flatten().tolist()
        print("  wrong rows inside the piece:", rows_bad[:8], "..", len(rows_bad), " cols:", ((d.abs() > 1e-3).any(0)).sum().item())
        if epi == "bias":
            for i, c in enumerate(chains):
                print(f"  |d + chain{i}| = {float((d + c[blk]).abs().max()):.3e}   |d - chain{i}| = {float((d - c[blk]).abs().max()):.3e}")
        # does a wrong ROW equal the same row-in-tile of some other tile of the reference (another 128-column block)?
        r_in = rows_bad[0]
        yrow = y[mt * 128 + r_in]
        nt_blk = pc // 4
        hits = []
        for m2 in range(R // 128):
            for n2 in range(N // 128):
                if float((yrow[nt_blk * 128:nt_blk * 128 + 128] - ref[m2 * 128 + r_in, n2 * 128:n2 * 128 + 128]).abs().max()) < 1e-3:
                    hits.append((m2, n2))
        print(f"  wrong row {r_in} of tile (m {mt}, n {nt_blk}) [linear tile {mt * (N // 128) + nt_blk}] equals row {r_in} of tile(s):", hits,
              [m2 * (N // 128) + n2 for m2, n2 in hits])
        # is the error of the wrong row a combination of its own 32-wide K-chunk products (a chunk missing: -1, doubled: +1)?
        cols = slice(nt_blk * 128, nt_blk * 128 + 128)
        xr = x[mt * 128 + r_in]
        P = torch.stack([xr[k:k + 32] @ w[cols, k:k + 32].t() for k in range(0, K, 32)])          # [K/32, 128]
        drow = (y[mt * 128 + r_in, cols] - ref[mt * 128 + r_in, cols])
        coef = torch.linalg.lstsq(P.t().double(), drow.double().unsqueeze(1)).solution.flatten()
        resid = float((P.t().double() @ coef - drow.double()).abs().max())
        print("  chunk coefficients:", [round(float(c), 3) for c in coef], " residual %.2e" % resid, " |d| %.2e" % float(drow.abs().max()))
        # does the piece equal some other piece of the reference?
        yb = y[blk]
        same = [(m2, p2) for m2 in range(R // 128) for p2 in range(N // 32)
                if float((yb - ref[m2 * 128:m2 * 128 + 128, p2 * 32:p2 * 32 + 32]).abs().max()) < 1e-3][:4]
        print("  equals reference piece(s):", same)
