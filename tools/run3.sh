mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_aff.py -m gpu -q > gpurun_out/pytest_aff.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_aff.log; tail -4 gpurun_out/pytest_aff.log
timeout 300 python benchmarks/op_bench.py --shape small_s0 --dtype bf16 --once > gpurun_out/once_bf16.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"dot_rows|axpy_rows|csr_rows|wf_fwd|wf_dw|wf_df" -c 8 -o gpurun_out/r1_ops_small_s0_bf16 -f python benchmarks/op_bench.py --shape small_s0 --dtype bf16 --once > gpurun_out/ncu_bf16.log 2>&1
echo "ncu exit $?"
timeout 300 python benchmarks/op_bench.py --shape mini_s0 --dtype f32 > gpurun_out/op_mini_f32.log 2>&1; cat gpurun_out/op_mini_f32.log
timeout 300 python benchmarks/op_bench.py --shape base_s0 --dtype bf16 > gpurun_out/op_base_bf16.log 2>&1; cat gpurun_out/op_base_bf16.log
