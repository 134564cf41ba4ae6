# round 2, second GPU call: rewritten TMA kernel (unrolled fast path, K/V staging reuse) -- parity at small and bench scale, timing, ncu
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "fused_attention_core" > gpurun_out/t1.log 2>&1; echo "t1 exit $?"; tail -4 gpurun_out/t1.log
timeout 900 python -m pytest tests/test_gpu_scale.py tests/test_gpu_aff.py -m gpu -q -s > gpurun_out/t_scale.log 2>&1; echo "scale exit $?"; grep -E "rel err|res2|d_w|passed|failed|Error|error" gpurun_out/t_scale.log | tail -40
for cfg in "16384 3 32 16 small_s0" "16384 2 16 16 mini_s0" "4096 6 32 16 small_s1" "1024 12 32 16 small_s2"; do set -- $cfg
 for dt in bf16 f32; do for tma in 1 0; do
  echo "== $5 $dt tma=$tma"; CLUSTEN_TMA_ATTN=$tma timeout 300 python benchmarks/attn_bench.py --n $1 --heads $2 --c $3 --batch $4 --dtype $dt --fwd-only 2>&1 | grep -E "clusten_attn_fwd|Error|error|pack_flags" | tee -a gpurun_out/attn_bench_r2_second.log
 done; done; done
for dt in bf16 f32; do
  CMD="python benchmarks/attn_bench.py --n 16384 --heads 3 --c 32 --batch 16 --dtype $dt --fwd-only --iters 2"
  $CMD > gpurun_out/plain_$dt.log 2>&1 &&
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fused_tma -s 3 -c 1 -f -o gpurun_out/r2_tma_v2_small_s0_$dt $CMD > gpurun_out/ncu_$dt.log 2>&1
  echo "ncu $dt exit $?"
done
