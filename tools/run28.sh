mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log
for cfg in "655 8" "3276 4" "131 16"; do set -- $cfg; echo "== n=$1 heads=$2"; timeout 300 python benchmarks/attn_bench.py --n $1 --heads $2 2>&1 | tail -6; done > gpurun_out/attn_bench_v7.log 2>&1
cat gpurun_out/attn_bench_v7.log
timeout 600 python bench.py --no-cpu-baseline --workload aff_tiny15_train_b32_512_bf16 --steps 5 --warmup 3 > gpurun_out/bench_tiny_v7.json 2> gpurun_out/bench_tiny_v7.err; cat gpurun_out/bench_tiny_v7.json; tail -3 gpurun_out/bench_tiny_v7.err
