mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_integer.py tests/test_gpu_aff.py -m gpu -q -x > gpurun_out/pytest_int.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_int.log; tail -4 gpurun_out/pytest_int.log
timeout 600 python bench.py --no-cpu-baseline --steps 10 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('mini', d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['per_entry_ms_per_step'])"
timeout 600 python bench.py --no-cpu-baseline --workload aff_small_fwd_b1_1024x2048 --steps 5 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('small-big', d['ms_per_step'], d['value'], d['roofline']['per_entry_ms_per_step'])"
