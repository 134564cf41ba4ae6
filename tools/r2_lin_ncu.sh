mkdir -p gpurun_out
CMD="python benchmarks/linear_bench.py --model small --only 2:q+kv --check-only --split f16"
$CMD > gpurun_out/lin_plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:linear_tc_kernel -c 1 -f -o gpurun_out/r2_linear_tc_v9_f16 $CMD > gpurun_out/ncu_lin.log 2>&1; echo "ncu $?"
