mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dropin.py -m gpu -q -s 2>&1 | grep -E "passed|failed|FAILED|\[\(" | tail -5
for tma in 0 1; do
 echo "== bench mini fp32 TMA=$tma"; CLUSTEN_TMA_ATTN=$tma timeout 600 python bench.py --no-cpu-baseline --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['roofline']['per_entry_ms_per_step']['clusten_attn_fwd'])"
 echo "== bench small fp32 TMA=$tma"; CLUSTEN_TMA_ATTN=$tma timeout 600 python bench.py --no-cpu-baseline --no-extras --workload aff_small_fwd_b16_512 --steps 10 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['per_entry_ms_per_step'].get('clusten_attn_fwd'))"
 echo "== bench tiny train bf16 TMA=$tma"; CLUSTEN_TMA_ATTN=$tma timeout 900 python bench.py --no-cpu-baseline --no-extras --workload aff_tiny15_train_b32_512_bf16 --steps 8 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['per_entry_ms_per_step'].get('clusten_attn_fwd'))"
done
timeout 600 python benchmarks/head_bench.py 2>&1 | tail -2
timeout 600 python benchmarks/op_bench.py --shape small_s0 --dtype f32 --iters 5 2>&1 | grep -E "wg_fwd|qk_fwd|av_fwd|wf_fwd"
