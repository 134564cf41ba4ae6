mkdir -p gpurun_out
for wl in aff_tiny15_train_b32_512_bf16 aff_base_train_b2_512x1024_bf16; do
timeout 900 python bench.py --workload $wl --no-extras > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "bench $wl rc=$?"; python - $wl <<'PY'
import json, sys
try:
    r=json.loads(open(f"gpurun_out/bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print(sys.argv[1], r["value"], r["ms_per_step"], "e2e", r["e2e"]["value"], r["roofline"]["kernel"], r["roofline"]["frac"])
except Exception as e:
    print("parse failed", e); print(open(f"gpurun_out/bench_{sys.argv[1]}.err").read()[-800:])
PY
done
