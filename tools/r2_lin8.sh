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
for sp in f16 tf32; do for m in mini small; do
timeout 300 python benchmarks/linear_bench.py --model $m --split $sp > gpurun_out/lin_${m}_$sp.log 2>&1; echo "$m $sp rc=$?"; show gpurun_out/lin_${m}_$sp.log | cut -c1-110
done; done
for s in "16384 256 768" "4096 768 2304"; do TRIALS=6 timeout 300 python tools/lin_race.py $s bias 2>&1 | grep -E "^trial" | tr "\n" ";"; echo; done
