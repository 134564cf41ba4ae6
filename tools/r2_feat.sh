mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/feat_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/feat_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
line() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['gpu_launches'], d['roofline']['kernel'], d['roofline']['frac'])"; }
timeout 600 python bench.py --no-cpu-baseline --no-extras 2>/dev/null | tee gpurun_out/feat_bench_mini_1.json | line mini_feat_1
CLUSTEN_REL_POS_FEATURES=0 timeout 600 python bench.py --no-cpu-baseline --no-extras 2>/dev/null | tee gpurun_out/feat_bench_mini_0.json | line mini_feat_0
