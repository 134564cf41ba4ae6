# round 2, third GPU call: TMA kernel v3 (prefetched bias, parallel prologue) -- parity, timing (cap sweep), ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_scale.py tests/test_gpu_aff.py -m gpu -q -x -k "fused or golden or scale or graph or mask" 2>&1 | tail -3
for cfg in "16384 3 32 16 small_s0" "16384 2 16 16 mini_s0" "4096 6 32 16 small_s1" "1024 12 32 16 small_s2" "3276 4 32 32 tiny15_s1" "32768 4 32 2 base_s0"; do set -- $cfg
 M=48; MM=8; G=128; if [ "$5" = "base_s0" ]; then M=144; MM=24; G=256; fi
 for dt in bf16 f32; do for tma in 1 0; do
  echo "== $5 $dt tma=$tma"; CLUSTEN_TMA_ATTN=$tma timeout 300 python benchmarks/attn_bench.py --n $1 --heads $2 --c $3 --batch $4 --dtype $dt --fwd-only --m $MM --nbhd $M 2>&1 | grep -E "clusten_attn_fwd|Error|error|pack_flags" | cut -c1-260 | tee -a gpurun_out/attn_bench_r2_third.log
 done; done; done
for cap in 24 28 32 40; do echo "== small_s0 bf16 cap=$cap"; CLUSTEN_TMA_CAP=$cap timeout 300 python benchmarks/attn_bench.py --n 16384 --heads 3 --c 32 --batch 16 --dtype bf16 --fwd-only 2>&1 | grep -E "clusten_attn_fwd" | tee -a gpurun_out/attn_bench_r2_third.log; done
for dt in bf16 f32; do
  CMD="python benchmarks/attn_bench.py --n 16384 --heads 3 --c 32 --batch 16 --dtype $dt --fwd-only --iters 2"
  $CMD > gpurun_out/plain_$dt.log 2>&1 &&
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fused_tma -s 3 -c 1 -f -o gpurun_out/r2_tma_v3_small_s0_$dt $CMD > gpurun_out/ncu_$dt.log 2>&1
  echo "ncu $dt exit $?"
done
