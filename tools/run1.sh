mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 300 python benchmarks/op_bench.py --shape small_s0 --dtype bf16 > gpurun_out/op_small_bf16.log 2>&1
timeout 300 python benchmarks/op_bench.py --shape small_s0 --dtype f32 --ref > gpurun_out/op_small_f32.log 2>&1
timeout 300 python benchmarks/op_bench.py --shape cfg1 --dtype f32 --ref > gpurun_out/op_cfg1_f32.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -3; cat gpurun_out/op_small_bf16.log
