mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_scale.py tests/test_gpu_aff.py -m gpu -q -x -s -k "fused or golden or scale or mask" 2>&1 | grep -E "rel err|aff_|passed|failed" | tail -16
for cfg in "16384 3 32 16 small_s0" "16384 2 16 16 mini_s0" "4096 6 32 16 small_s1" "1024 12 32 16 small_s2"; do set -- $cfg
 for dt in bf16 f32; do for tma in 1 0; do
  echo "== $5 $dt tma=$tma"; CLUSTEN_TMA_ATTN=$tma timeout 300 python benchmarks/attn_bench.py --n $1 --heads $2 --c $3 --batch $4 --dtype $dt --fwd-only 2>&1 | grep -E "clusten_attn_fwd|Error|error" | cut -c1-200 | tee -a gpurun_out/attn_bench_r2_fourth.log
 done; done; done
timeout 900 python bench.py > gpurun_out/bench_default_r2a.json 2> gpurun_out/bench_default_r2a.err; echo "bench exit $?"; cat gpurun_out/bench_default_r2a.json; tail -5 gpurun_out/bench_default_r2a.err
