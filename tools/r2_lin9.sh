mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_aff.py tests/test_gpu_dropin.py tests/test_gpu_scale.py -x -q -m gpu > gpurun_out/pytest_aff.log 2>&1; echo "pytest aff rc=$?"; tail -4 gpurun_out/pytest_aff.log
for wl in aff_mini_fwd_b16_512 aff_small_fwd_b16_512; do
timeout 600 python bench.py --workload $wl --no-extras > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "bench rc=$?"; python - $wl <<'PY'
import json, sys
r=json.loads(open(f"gpurun_out/bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
print(sys.argv[1], r["value"], r["ms_per_step"], "e2e", r["e2e"]["value"], r["roofline"]["kernel"], r["roofline"]["frac"], r["roofline"].get("tensor"))
print(r["roofline"]["per_entry_ms_per_step"])
PY
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
