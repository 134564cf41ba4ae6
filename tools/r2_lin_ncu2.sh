mkdir -p gpurun_out
CMD="python benchmarks/linear_bench.py --model mini --only 2:q+kv --check-only --split f16"
$CMD > gpurun_out/lin2_plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:linear_tc_kernel -c 1 -f -o gpurun_out/r2_linear_tc_resident $CMD > gpurun_out/ncu_lin2.log 2>&1; echo "ncu $?"
CLUSTEN_TC_RESIDENT=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:linear_tc_kernel -c 1 -f -o gpurun_out/r2_linear_tc_streamed $CMD > gpurun_out/ncu_lin3.log 2>&1; echo "ncu $?"
