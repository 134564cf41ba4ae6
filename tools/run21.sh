mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^fwd_kernel|^dw_kernel|^df_oct_kernel" -o gpurun_out/r1_wf2_small_s0_bf16_v8 -f python benchmarks/op_bench.py --shape small_s0 --dtype bf16 --once > gpurun_out/ncu_wf2.log 2>&1
echo "ncu wf2 exit $?"; tail -3 gpurun_out/ncu_wf2.log
