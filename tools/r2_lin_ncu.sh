mkdir -p gpurun_out
CMD="python benchmarks/linear_bench.py --model mini --only 2:q+kv,1:fc1 --iters 2"
$CMD > gpurun_out/lin_plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:linear_tc_kernel -s 3 -c 1 -f -o gpurun_out/r2_linear_tc_v3_fc1 $CMD > gpurun_out/ncu_lin.log 2>&1; echo "ncu $?"
CMD="python benchmarks/linear_bench.py --model mini --only 2:q+kv --iters 2"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:linear_tc_kernel -s 3 -c 1 -f -o gpurun_out/r2_linear_tc_v3_mid $CMD > gpurun_out/ncu_lin2.log 2>&1; echo "ncu $?"
