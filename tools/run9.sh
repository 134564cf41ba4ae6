mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -8 gpurun_out/pytest_gpu.log
for cfg in "small_s0 bf16" "small_s0 f32"; do set -- $cfg
timeout 300 python benchmarks/op_bench.py --shape $1 --dtype $2 > gpurun_out/op_$1_$2.log 2>&1; grep -E "qk_fwd|av_fwd|qk_bwd|av_bwd|pack|csr|shape" gpurun_out/op_$1_$2.log; done
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_mini_v2.json 2> gpurun_out/bench_mini_v2.err; cat gpurun_out/bench_mini_v2.json
timeout 600 python bench.py --no-cpu-baseline --workload aff_tiny15_train_b32_512_bf16 --steps 5 > gpurun_out/bench_tiny_v2.json 2> gpurun_out/bench_tiny_v2.err; cat gpurun_out/bench_tiny_v2.json; tail -3 gpurun_out/bench_tiny_v2.err
