mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/scores_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/scores_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
line() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['gpu_launches'], d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['traffic'])"; }
for g in 1 0; do
CLUSTEN_MERGE_SCORES=$g timeout 600 python bench.py --no-cpu-baseline --no-extras 2>/dev/null | tee gpurun_out/scores_bench_mini_$g.json | line mini_scores_$g
done
timeout 600 python bench.py --no-cpu-baseline --no-extras --workload aff_small_fwd_b16_512 --steps 10 2>/dev/null | tee gpurun_out/scores_bench_small_1.json | line small_scores_1
timeout 600 python bench.py --no-cpu-baseline --no-extras --workload aff_tiny15_train_b32_512_bf16 --steps 5 --warmup 3 2>/dev/null | tee gpurun_out/scores_bench_tiny_train.json | line tiny_train
