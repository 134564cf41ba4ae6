mkdir -p gpurun_out
timeout 300 python benchmarks/linear_bench.py --model mini > gpurun_out/lin_mini.log 2>&1; echo "mini rc=$?"; python - <<'PY'
import json
for f in ("gpurun_out/lin_mini.log",):
    for l in open(f):
        try: r=json.loads(l)
        except Exception: print(l.strip()); continue
        if "layer" in r: print(r["layer"], r["R"], r["K"], r["N"], "err %.1e/%.1e"%(r["err_tc"], r["err_cublas"]), r["tail_equal"], r.get("us_tc"), r.get("us_cublas"), r.get("gbps_tc"), r.get("tflops_tc"))
        else: print(r)
PY
timeout 300 python benchmarks/linear_bench.py --model small > gpurun_out/lin_small.log 2>&1; echo "small rc=$?"; python - <<'PY'
import json
for f in ("gpurun_out/lin_small.log",):
    for l in open(f):
        try: r=json.loads(l)
        except Exception: print(l.strip()); continue
        if "layer" in r: print(r["layer"], r["R"], r["K"], r["N"], "err %.1e/%.1e"%(r["err_tc"], r["err_cublas"]), r["tail_equal"], r.get("us_tc"), r.get("us_cublas"), r.get("gbps_tc"), r.get("tflops_tc"))
        else: print(r)
PY
CMD="python benchmarks/linear_bench.py --model mini --only 2:q+kv --iters 2"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:linear_tc_kernel -s 3 -c 1 -f -o gpurun_out/r2_linear_tc_v2_mid $CMD > gpurun_out/ncu_lin.log 2>&1; echo "ncu $?"
