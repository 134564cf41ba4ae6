# launch list + DRAM traffic of the default bench command (B200_PROFILING.md recipe: plain run first), N=1 final lines
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
$CMD > gpurun_out/plain_bench.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 2600 --csv --log-file gpurun_out/r2_launches_default.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches exit $?"
$CMD > gpurun_out/plain_bench2.log 2>&1 &&
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:attn_fused -s 60 -c 24 --csv --log-file gpurun_out/r2_traffic_attn_fwd.csv $CMD > gpurun_out/ncu_traffic.log 2>&1
echo "ncu traffic exit $?"
timeout 300 python benchmarks/op_bench.py --shape small_s0 --dtype f32 --iters 5 2>&1 | grep -E 'wf_bwd|wf_fwd"'
for wl in aff_tiny15_train_b32_512_bf16 aff_base_train_b2_512x1024_bf16 aff_small_fwd_b16_512 aff_small_fwd_b1_1024x2048; do
  timeout 900 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_${wl}_n1.json 2> gpurun_out/r2_bench_${wl}_n1.err; echo "$wl exit $?"; cut -c1-200 gpurun_out/r2_bench_${wl}_n1.json
done
