mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_ops.py tests/test_gpu_aff.py -m gpu -q -x -k "fused or golden" > gpurun_out/pytest_fused.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_fused.log; tail -15 gpurun_out/pytest_fused.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_mini_v3.json 2> gpurun_out/bench_mini_v3.err; cat gpurun_out/bench_mini_v3.json; tail -3 gpurun_out/bench_mini_v3.err
