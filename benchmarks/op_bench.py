"""Per-kernel roofline micro-benchmark of the CLUSTEN ops at backbone-scale shapes.

For every op: CUDA-event time of our C-ABI call (L2 flushed between iterations), ALGORITHMIC bytes (each operand read
once, each result written once, idx counted as the int64 the interface delivers), achieved GB/s and the fraction of
MEASURED_PEAKS.json's hbm_gbs.  With --ref also times the reference's own CUDA kernels (oracle/_ref) on the same inputs.

    python benchmarks/op_bench.py [--shape small_s0|mini_s0|tiny_s0|base_s0|cfg1] [--dtype bf16|f32] [--ref] [--iters 10]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SHAPES = {  # B, H, N, C(per head), M, m ; merge: N' ; WF channel dim = H*C
    "cfg1":     dict(B=2, H=2, N=4096, C=32, M=48, m=8, Nq_wf=1024, grid=(128, 128)),
    "mini_s0":  dict(B=16, H=2, N=16384, C=16, M=48, m=8, Nq_wf=4096, grid=(128, 128)),
    "tiny_s0":  dict(B=32, H=2, N=16384, C=32, M=48, m=8, Nq_wf=3276, grid=(128, 128)),
    "small_s0": dict(B=32, H=3, N=16384, C=32, M=48, m=8, Nq_wf=4096, grid=(128, 128)),
    "small_s1": dict(B=32, H=6, N=4096, C=32, M=48, m=8, Nq_wf=1024, grid=(128, 128)),
    "base_s0":  dict(B=4, H=4, N=32768, C=32, M=144, m=24, Nq_wf=8192, grid=(128, 256)),
}


def peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def real_idx(B, N, M, m, grid):
    """The stage pipeline of the backbone (product kernels): positions -> space_filling_cluster -> kNN clusters -> member_idx
    (aff.py:469-478).  Stage-0 tokens sit on the stem grid, so every sample of the batch has the same neighbourhoods --
    exactly what the ops see in the backbone; later-stage shapes reuse one sample's structure."""
    from _inputs import stage_structure
    h, w = grid
    nb = stage_structure(1, N, h, w, m, M, seed=0)[1]
    return nb.expand(B, -1, -1).contiguous()


def structured_idx(B, N, M, m, gen):
    """Curve-ordered, run-structured neighbourhoods like aff.py:475-478 produces: M/m runs of m consecutive rows drawn
    from clusters near the token's own cluster."""
    nnc = M // m
    k = N // m
    own = (torch.arange(N, device="cuda") // m).view(1, N, 1)
    off = torch.randint(-8, 9, (B, N, nnc), device="cuda", generator=gen)
    off[..., 0] = 0
    cl = (own + off).clamp_(0, k - 1)
    return (cl.unsqueeze(-1) * m + torch.arange(m, device="cuda")).reshape(B, N, M).contiguous()


def time_fn(fn, iters, flush):
    if iters <= 0:                          # --once: a single cold launch
        flush.add_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b)
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.add_(1)                       # > L2 (126 MB): evicts the operands
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def bench_ops(args, log=print):
    """Times every op of the shape; returns the rows (dicts: op, ms, algo_MB, GBs, frac = fraction of peak on the INTERFACE bytes,
    frac_data = the same on the data operands and results alone, i.e. without the int64 index the tile-union kernels never read)."""
    from autofocusformermod_b200 import _lib, ops
    if os.path.dirname(os.path.abspath(__file__)) not in sys.path:
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    print = log
    S = SHAPES[args.shape]
    B, H, N, C, M, m = S["B"], S["H"], S["N"], S["C"], S["M"], S["m"]
    dt = {"bf16": torch.bfloat16, "f32": torch.float32, "f16": torch.float16}[args.dtype]
    s = torch.finfo(dt).bits // 8
    gen = torch.Generator(device="cuda").manual_seed(0)
    peak, psrc = peak_gbs()
    flush = torch.zeros(256 << 20, dtype=torch.uint8, device="cuda")
    idx = torch.randint(0, N, (B, N, M), device="cuda", generator=gen) if args.random_idx else real_idx(B, N, M, m, S["grid"])
    # token-major memory, head-major views: exactly what aff.py:111-113 hands to the ops
    q = torch.randn(B, N, H, C, device="cuda", generator=gen).to(dt).permute(0, 2, 1, 3)
    kv = torch.randn(B, N, H, 2, C, device="cuda", generator=gen).to(dt).permute(3, 0, 2, 1, 4)
    k, v = kv[0], kv[1]
    attn = torch.randn(B, H, N, M, device="cuda", generator=gen).softmax(-1).to(dt)
    d_attn = torch.randn(B, H, N, M, device="cuda", generator=gen).to(dt)
    d_feat = torch.randn(B, N, H, C, device="cuda", generator=gen).to(dt).permute(0, 2, 1, 3)
    L = _lib.lib()
    st = lambda: torch.cuda.current_stream().cuda_stream
    code = _lib.dtype_code(q)
    off, ent = ops.inverse_neighbour_list(idx, N)
    ops.USE_TILE_KERNELS = not args.generic
    pack = ops.neighbourhood_pack(idx, N, inverse=True)
    pk = 0 if pack is None else pack.data_ptr()
    print(json.dumps({"pack_flags(generic,maxU,impure,overlimit)": ops.pack_flags(idx, N)}), flush=True)
    out_attn = torch.empty(B, H, N, M, device="cuda", dtype=dt)
    feat = torch.empty(B, N, H, C, device="cuda", dtype=dt).permute(0, 2, 1, 3)
    d_q, d_k, d_v = torch.empty_like(q), torch.empty_like(q), torch.empty_like(q)
    s3 = lambda t: (t.stride(0), t.stride(1), t.stride(2))
    BHNC, BHNM, BNM8 = B * H * N * C * s, B * H * N * M * s, B * N * M * 8
    rows = []

    def run(name, fn, nbytes, idx_bytes=0):
        ms = time_fn(fn, 0 if args.once else args.iters, flush)
        gbs = nbytes / ms / 1e6
        rows.append(dict(op=name, ms=round(ms, 4), algo_MB=round(nbytes / 1e6, 1), GBs=round(gbs, 1), frac=round(gbs / peak, 3),
                         frac_data=round((nbytes - idx_bytes) / ms / 1e6 / peak, 3)))
        print(json.dumps(rows[-1]), flush=True)

    ck = lambda rc: _lib.check(rc, "bench")
    run("qk_fwd", lambda: ck(L.clusten_qk_fwd(q.data_ptr(), k.data_ptr(), idx.data_ptr(), pk, out_attn.data_ptr(), B, H, N, N, C, M,
                                              *s3(q), *s3(k), code, st())), 2 * BHNC + BNM8 + BHNM, BNM8)
    run("av_fwd", lambda: ck(L.clusten_av_fwd(attn.data_ptr(), v.data_ptr(), idx.data_ptr(), pk, feat.data_ptr(), B, H, N, N, C, M,
                                              *s3(attn), *s3(v), *s3(feat), code, st())), BHNM + 2 * BHNC + BNM8, BNM8)
    run("qk_bwd", lambda: ck(L.clusten_qk_bwd(d_attn.data_ptr(), q.data_ptr(), k.data_ptr(), idx.data_ptr(), off.data_ptr(),
                                              ent.data_ptr(), pk, d_q.data_ptr(), d_k.data_ptr(), B, H, N, N, C, M,
                                              *s3(q), *s3(k), *s3(d_q), *s3(d_k), code, st())), BHNM + 4 * BHNC + BNM8, BNM8)
    run("av_bwd", lambda: ck(L.clusten_av_bwd(d_feat.data_ptr(), attn.data_ptr(), v.data_ptr(), idx.data_ptr(), off.data_ptr(),
                                              ent.data_ptr(), pk, out_attn.data_ptr(), d_v.data_ptr(), B, H, N, N, C, M,
                                              *s3(d_feat), *s3(attn), *s3(v), *s3(d_v), code, st())), 3 * BHNC + 2 * BHNM + BNM8, BNM8)
    pb = L.clusten_pack_bytes(B, N, M, N)
    pbuf = torch.empty(pb, dtype=torch.uint8, device="cuda")
    run("pack_build", lambda: ck(L.clusten_pack_build(idx.data_ptr(), 0, B, N, M, N, pbuf.data_ptr(), pb, st())), BNM8)
    run("pack_inverse", lambda: ck(L.clusten_pack_inverse(pbuf.data_ptr(), pb, B, N, M, N, st())), BNM8)
    ws_b = L.clusten_csr_workspace_bytes(B, N, M, N)
    ws = torch.empty(ws_b, dtype=torch.uint8, device="cuda")
    run("csr_build", lambda: ck(L.clusten_csr_build(idx.data_ptr(), B, N, M, N, off.data_ptr(), ent.data_ptr(), ws.data_ptr(), ws_b, None, st())),
        BNM8 + B * N * M * 4 + B * (N + 1) * 4)
    # WF merge: N' tokens in top-k (not curve) order gather M rows of the [B,N,H*C] feature map
    Nq, Cw, IC = S["Nq_wf"], H * C, 4
    sel = torch.stack([torch.randperm(N, device="cuda", generator=gen)[:Nq] for _ in range(B)])
    idx_w = idx.gather(1, sel.unsqueeze(-1).expand(-1, -1, M)).contiguous()
    w = torch.randn(B, Nq, M, IC, device="cuda", generator=gen).to(dt)
    f = torch.randn(B, N, Cw, device="cuda", generator=gen).to(dt)
    out_w = torch.empty(B, Nq, IC, Cw, device="cuda", dtype=dt)
    d_out = torch.randn(B, Nq, IC, Cw, device="cuda", generator=gen).to(dt)
    d_w, d_f = torch.empty_like(w), torch.empty_like(f)
    plan = ops.wf_plan(idx_w, N) if not args.generic else None          # (fp32: only the backward's octet-form d_f uses it)
    pl = 0 if plan is None else plan.data_ptr()
    offw, entw = ops.inverse_neighbour_list(idx_w, N, wf_plan_buf=plan)
    if plan is not None:
        print(json.dumps({"wf_plan_flags(generic,impure,maxlist)": ops.wf_plan_flags(idx_w, N)}), flush=True)
        plb = L.clusten_wf_plan_bytes(B, Nq, M, N)
        pbuf2 = torch.empty(plb, dtype=torch.uint8, device="cuda")
        run("wf_plan_build", lambda: ck(L.clusten_wf_plan_build(idx_w.data_ptr(), B, Nq, M, N, pbuf2.data_ptr(), plb, st())), B * Nq * M * 8)
    wb, fb, ib, ob = B * Nq * M * IC * s, B * N * Cw * s, B * Nq * M * 8, B * Nq * IC * Cw * s
    run("wf_fwd", lambda: ck(L.clusten_wf_fwd(w.data_ptr(), f.data_ptr(), idx_w.data_ptr(), pl, out_w.data_ptr(), B, Nq, N, Cw, M, IC,
                                              f.stride(0), f.stride(1), code, st())), wb + fb + ib + ob, ib)
    if plan is not None:
        run("wf_fwd (no plan: top-k token order)", lambda: ck(L.clusten_wf_fwd(w.data_ptr(), f.data_ptr(), idx_w.data_ptr(), 0, out_w.data_ptr(), B, Nq, N, Cw, M, IC,
                                                                       f.stride(0), f.stride(1), code, st())), wb + fb + ib + ob, ib)
    run("wf_bwd", lambda: ck(L.clusten_wf_bwd(d_out.data_ptr(), w.data_ptr(), f.data_ptr(), idx_w.data_ptr(), offw.data_ptr(),
                                              entw.data_ptr(), pl, d_w.data_ptr(), d_f.data_ptr(), B, Nq, N, Cw, M, IC,
                                              f.stride(0), f.stride(1), d_f.stride(0), d_f.stride(1), code, st())),
        ob + wb + fb + ib + wb + fb, ib)
    # weighted gather at the FPN upsample of the pixel decoder (point_utils.py:103-114): every token of the stage pulls K = 4
    # nearest tokens of the next stage (Nq_wf of them), channel dim 256
    from autofocusformermod_b200 import point_utils as pu
    from _inputs import grid_positions, random_positions
    hh, ww = S["grid"]
    if N == hh * ww:
        Cg, K = 256, 4
        qpos = grid_positions(1, hh, ww).cuda()
        dpos = random_positions(1, Nq, hh, ww, seed=1).cuda()
        idx_g = pu.knn_keops(qpos, dpos, K).expand(B, -1, -1).contiguous()
        wg = torch.rand(B, N, K, device="cuda", generator=gen).to(dt)
        fg = torch.randn(B, Nq, Cg, device="cuda", generator=gen).to(dt)
        og = torch.empty(B, N, Cg, device="cuda", dtype=dt)
        run("wg_fwd", lambda: ck(L.clusten_wg_fwd(idx_g.data_ptr(), wg.data_ptr(), fg.data_ptr(), og.data_ptr(), B, N, Nq, Cg, K,
                                                  fg.stride(0), fg.stride(1), code, st())),
            B * N * K * s + B * Nq * Cg * s + B * N * K * 8 + B * N * Cg * s, B * N * K * 8)
    if args.ref and dt != torch.bfloat16:
        from oracle import ref_cuda
        qc, kc, vc = q.contiguous(), k.contiguous(), v.contiguous()
        run("REF qk_fwd (incl. its K transpose)", lambda: ref_cuda.qk_forward(qc, kc, idx), 2 * BHNC + BNM8 + BHNM, BNM8)
        run("REF av_fwd", lambda: ref_cuda.av_forward(attn, vc, idx), BHNM + 2 * BHNC + BNM8, BNM8)
        run("REF qk_bwd", lambda: ref_cuda.qk_backward(d_attn, qc, kc, idx), BHNM + 4 * BHNC + BNM8, BNM8)
        run("REF av_bwd", lambda: ref_cuda.av_backward(d_feat.contiguous(), attn, vc, idx), 3 * BHNC + 2 * BHNM + BNM8, BNM8)
        run("REF wf_fwd", lambda: ref_cuda.wf_forward(w, f, idx_w), wb + fb + ib + ob)
        run("REF wf_bwd", lambda: ref_cuda.wf_backward(d_out, w, f, idx_w), ob + wb + fb + ib + wb + fb)
    print(json.dumps(dict(shape=args.shape, dtype=args.dtype, peak_gbs=peak, peak_source=psrc, idx="random" if args.random_idx else "reference pipeline (clustered)")))
    return rows


def default_args(**kw):
    a = argparse.Namespace(shape="small_s0", dtype="bf16", iters=10, ref=False, random_idx=False, once=False, generic=False)
    a.__dict__.update(kw)
    return a


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="small_s0")
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--ref", action="store_true")
    ap.add_argument("--random-idx", action="store_true")
    ap.add_argument("--once", action="store_true", help="run every op exactly once (for ncu captures)")
    ap.add_argument("--generic", action="store_true", help="row-gather kernels only (no tile pack)")
    bench_ops(ap.parse_args())


if __name__ == "__main__":
    main()
