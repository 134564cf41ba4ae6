mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_ops.py -m gpu -q -s -k "dropin or reference_aff or msdetrpc or fused or wg or weighted or golden" 2>&1 | grep -E "passed|failed|FAILED|Error|\[\(" | tail -12
for cfg in "16384 3 32 16 small_s0" "16384 2 16 16 mini_s0" "4096 6 32 16 small_s1" "1024 12 32 16 small_s2"; do set -- $cfg
 for dt in bf16 f32; do for mode in "1 1 0" "1 0 0" "1 0 32" "0 1 0"; do set -- $cfg; m=($mode)
  echo "== $5 $dt tma=${m[0]} packed=${m[1]} cap=${m[2]}"; CLUSTEN_TMA_ATTN=${m[0]} CLUSTEN_TMA_PACKED=${m[1]} CLUSTEN_TMA_CAP=${m[2]} timeout 300 python benchmarks/attn_bench.py --n $1 --heads $2 --c $3 --batch $4 --dtype $dt --fwd-only 2>&1 | grep -E "clusten_attn_fwd|Error|error" | cut -c1-200 | tee -a gpurun_out/attn_bench_r2_sixth.log
 done; done; done
