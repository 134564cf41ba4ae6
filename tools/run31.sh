mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_mini_v9.json 2> gpurun_out/bench_mini_v9.err; cat gpurun_out/bench_mini_v9.json; tail -3 gpurun_out/bench_mini_v9.err
timeout 600 python bench.py --no-cpu-baseline --workload aff_tiny15_train_b32_512_bf16 --steps 5 --warmup 3 > gpurun_out/bench_tiny_v9.json 2> gpurun_out/bench_tiny_v9.err; cat gpurun_out/bench_tiny_v9.json; tail -3 gpurun_out/bench_tiny_v9.err
