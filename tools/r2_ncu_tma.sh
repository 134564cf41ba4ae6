# ncu --set full of the TMA-staged fused forward at Small stage 0 (bf16 and fp32), per the profiling recipe (plain run first)
mkdir -p gpurun_out
for dt in bf16 f32; do
  CMD="python benchmarks/attn_bench.py --n 16384 --heads 3 --c 32 --batch 16 --dtype $dt --fwd-only --iters 2"
  $CMD > gpurun_out/plain_$dt.log 2>&1 &&
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fused_tma -s 3 -c 1 -f -o gpurun_out/r2_tma_v1_small_s0_$dt $CMD > gpurun_out/ncu_$dt.log 2>&1
  echo "ncu $dt exit $?"; tail -3 gpurun_out/ncu_$dt.log
done
