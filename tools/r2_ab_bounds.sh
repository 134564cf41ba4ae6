mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_scale.py tests/test_gpu_aff.py tests/test_gpu_dropin.py -m gpu -q -x -k "fused or golden or scale or mask or dropin or reference_aff" 2>&1 | tail -2
for cfg in "16384 3 32 16 small_s0 8 48" "16384 2 16 16 mini_s0 8 48" "1024 12 32 16 small_s2 8 48" "32768 4 32 2 base_s0 24 144" "32760 4 32 2 base_s0_nomask 24 144"; do set -- $cfg
 for dt in bf16 f32; do
  echo "== $5 $dt"; timeout 300 python benchmarks/attn_bench.py --n $1 --heads $2 --c $3 --batch $4 --dtype $dt --fwd-only --m $6 --nbhd $7 --grid 256 2>&1 | grep -E "clusten_attn_fwd|Error|error" | cut -c1-200
 done; done
timeout 900 python bench.py --workload aff_base_train_b2_512x1024_bf16 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('base train', d['value'], d['ms_per_step'], d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['per_entry_ms_per_step']['clusten_attn_fwd'])"
timeout 900 python bench.py --no-cpu-baseline --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('mini', d['value'], d['ms_per_step'], d['roofline']['kernel'], d['roofline']['frac'])"
