mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -8 gpurun_out/pytest_gpu.log
timeout 300 python benchmarks/op_bench.py --shape base_s0 --dtype bf16 > gpurun_out/op_base_bf16.log 2>&1; grep -E "pack_flags|qk_fwd|av_fwd|qk_bwd|av_bwd|shape" gpurun_out/op_base_bf16.log
