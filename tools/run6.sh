mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "tile or config1 or views" > gpurun_out/pytest_ops.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_ops.log; tail -12 gpurun_out/pytest_ops.log
timeout 300 python benchmarks/op_bench.py --shape small_s0 --dtype bf16 > gpurun_out/op_small_bf16.log 2>&1; grep -E "qk_fwd|av_fwd|qk_bwd|av_bwd" gpurun_out/op_small_bf16.log
timeout 300 python benchmarks/op_bench.py --shape small_s0 --dtype f32 > gpurun_out/op_small_f32.log 2>&1; grep -E "qk_fwd|av_fwd|qk_bwd|av_bwd" gpurun_out/op_small_f32.log
timeout 300 python benchmarks/op_bench.py --shape mini_s0 --dtype f32 > gpurun_out/op_mini_f32.log 2>&1; grep -E "qk_fwd|av_fwd|qk_bwd|av_bwd" gpurun_out/op_mini_f32.log
