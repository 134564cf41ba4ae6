mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_vs_reference_kernels.py tests/test_gpu_aff.py -m gpu -q -x > gpurun_out/pytest_ops.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_ops.log; tail -8 gpurun_out/pytest_ops.log
for cfg in "small_s0 bf16" "base_s0 bf16"; do set -- $cfg
timeout 300 python benchmarks/op_bench.py --shape $1 --dtype $2 > gpurun_out/op_$1_$2.log 2>&1; grep -E "qk_fwd|av_fwd|qk_bwd|av_bwd|wf_|shape" gpurun_out/op_$1_$2.log; done
