mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_mini_v8.json 2> gpurun_out/bench_mini_v8.err; cat gpurun_out/bench_mini_v8.json; tail -3 gpurun_out/bench_mini_v8.err
timeout 600 python bench.py --no-cpu-baseline --no-graph > gpurun_out/bench_mini_v8_eager.json 2> gpurun_out/bench_mini_v8_eager.err; cat gpurun_out/bench_mini_v8_eager.json | cut -c1-400; tail -3 gpurun_out/bench_mini_v8_eager.err
timeout 600 python bench.py --no-cpu-baseline --workload aff_tiny15_train_b32_512_bf16 --steps 5 --warmup 3 > gpurun_out/bench_tiny_v8.json 2> gpurun_out/bench_tiny_v8.err; cat gpurun_out/bench_tiny_v8.json | cut -c1-300; tail -3 gpurun_out/bench_tiny_v8.err
