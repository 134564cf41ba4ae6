mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/stem_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/stem_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
line() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['roofline']['kernel'], d['roofline']['frac'])"; }
for g in 1 0; do
CLUSTEN_STEM_GEMM=$g timeout 600 python bench.py --no-cpu-baseline --no-extras 2>/dev/null | tee gpurun_out/stem_bench_mini_$g.json | line mini_stemgemm_$g
CLUSTEN_STEM_GEMM=$g timeout 600 python bench.py --no-cpu-baseline --no-extras --workload aff_small_fwd_b16_512 --steps 10 2>/dev/null | tee gpurun_out/stem_bench_small_$g.json | line small_stemgemm_$g
done
