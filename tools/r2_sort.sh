mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_integer.py tests/test_gpu_aff.py tests/test_gpu_dropin.py -x -q -m gpu > gpurun_out/pytest_int.log 2>&1; echo "pytest int rc=$?"; tail -4 gpurun_out/pytest_int.log
timeout 600 python bench.py --no-extras > gpurun_out/bench_sort.json 2> gpurun_out/bench_sort.err; echo "bench rc=$?"; python - <<'PY'
import json
r=json.loads(open("gpurun_out/bench_sort.json").read().strip().splitlines()[-1])
print(r["value"], r["ms_per_step"], "e2e", r["e2e"]["value"], r["gpu_launches"])
print(r["roofline"]["per_entry_ms_per_step"])
PY
