mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -25 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_mini_v5.json 2> gpurun_out/bench_mini_v5.err; cat gpurun_out/bench_mini_v5.json; tail -3 gpurun_out/bench_mini_v5.err
timeout 600 python bench.py --no-cpu-baseline --workload aff_tiny15_train_b32_512_bf16 --steps 3 --warmup 3 > gpurun_out/bench_tiny_v5.json 2> gpurun_out/bench_tiny_v5.err; cat gpurun_out/bench_tiny_v5.json; tail -3 gpurun_out/bench_tiny_v5.err
