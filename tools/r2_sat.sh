mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "linear_tc" > gpurun_out/sat_tests.log 2>&1; echo "linear tests exit $?"; tail -3 gpurun_out/sat_tests.log
line() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['roofline']['kernel'], d['roofline']['frac'])"; }
timeout 600 python bench.py --no-cpu-baseline --no-extras 2>/dev/null | tee gpurun_out/sat_bench_mini.json | line mini
timeout 600 python bench.py --no-cpu-baseline --no-extras --workload aff_small_fwd_b16_512 --steps 10 2>/dev/null | tee gpurun_out/sat_bench_small.json | line small
for shape in "16384 256 768 bias ln" "16384 512 256 residual" "65536 128 128 residual" "4096 384 1152 bias ln" "16384 384 1152 bias ln" "16384 768 384 residual"; do
timeout 120 python tools/lin_profile.py $shape 2>&1 | tail -3
done > gpurun_out/sat_lin_prof.txt 2>&1
cat gpurun_out/sat_lin_prof.txt
