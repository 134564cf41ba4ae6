# in-kernel position bias vs table variant: per-entry times at AFF-Small stage 0 / stage 2, and the end-to-end lines
mkdir -p gpurun_out
for cfg in "16384 3 32 16 small_s0" "1024 12 32 16 small_s2" "655 8 32 32 tiny15_s2"; do set -- $cfg
 for dt in bf16 f32; do for pb in "" "--inkernel-bias"; do
  echo "== $5 $dt $pb"; timeout 300 python benchmarks/attn_bench.py --n $1 --heads $2 --c $3 --batch $4 --dtype $dt $pb 2>&1 | grep -E '"entry"' | cut -c1-160 | tee -a gpurun_out/attn_bench_r2_pb_ab.log
 done; done; done
for pb in 0 1; do
 echo "== bench mini fp32 INKERNEL_BIAS=$pb"; CLUSTEN_INKERNEL_BIAS=$pb timeout 600 python bench.py --no-cpu-baseline --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['per_entry_ms_per_step'])"
 echo "== bench tiny train bf16 INKERNEL_BIAS=$pb"; CLUSTEN_INKERNEL_BIAS=$pb timeout 900 python bench.py --no-cpu-baseline --no-extras --workload aff_tiny15_train_b32_512_bf16 --steps 8 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['per_entry_ms_per_step'])"
done
echo "== goldens with the 3xTF32 Linear"; CLUSTEN_TC_LINEAR=1 timeout 600 python -m pytest tests/test_gpu_aff.py tests/test_gpu_dropin.py -m gpu -q -s -k "golden or reference_aff_class_runs" 2>&1 | grep -E "aff_|passed|failed|\(" | tail -14
for wl in aff_mini_fwd_b16_512 aff_small_fwd_b16_512; do for tc in 0 1; do
 echo "== bench $wl TC_LINEAR=$tc"; CLUSTEN_TC_LINEAR=$tc timeout 600 python bench.py --no-cpu-baseline --no-extras --workload $wl --steps 10 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
done; done
