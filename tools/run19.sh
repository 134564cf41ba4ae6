mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "wf or weighted or layer_norm or table" > gpurun_out/pytest_wf.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_wf.log; tail -15 gpurun_out/pytest_wf.log
for sh in small_s0 mini_s0 base_s0 small_s1; do timeout 300 python benchmarks/op_bench.py --shape $sh --dtype bf16 > gpurun_out/op_${sh}_bf16_v7.log 2>&1; tail -8 gpurun_out/op_${sh}_bf16_v7.log; done
