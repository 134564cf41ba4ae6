mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_integer.py tests/test_gpu_aff.py -x -q -m gpu > gpurun_out/pytest_int.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_int.log
timeout 600 python -c "
import sys; sys.path.insert(0,'benchmarks')
import int_bench, json
print(json.dumps(int_bench.integer_path_us()))" 2>&1 | tail -2
timeout 600 python bench.py --no-extras > gpurun_out/bench_prep.json 2> gpurun_out/bench_prep.err; python -c "
import json
r=json.loads(open('gpurun_out/bench_prep.json').read().strip().splitlines()[-1]); print(r['value'], r['ms_per_step'], r['e2e']['value'])"
