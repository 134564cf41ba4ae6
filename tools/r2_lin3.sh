mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    try: r=json.loads(l)
    except Exception: print(l.rstrip()[:300]); continue
    if "layer" in r: print(r["layer"], r["R"], r["K"], r["N"], "err %.1e/%.1e"%(r["err_tc"], r["err_cublas"]), r["tail_equal"], r.get("us_tc"), r.get("us_cublas"), r.get("gbps_tc"), r.get("tflops_tc"))
    else: print(r)
PY
}
timeout 300 python benchmarks/linear_bench.py --model mini $LB_ARGS > gpurun_out/lin_mini.log 2>&1; echo "mini rc=$?"; show gpurun_out/lin_mini.log
timeout 300 python benchmarks/linear_bench.py --model small $LB_ARGS > gpurun_out/lin_small.log 2>&1; echo "small rc=$?"; show gpurun_out/lin_small.log
