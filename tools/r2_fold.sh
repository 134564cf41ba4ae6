mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "linear_tc" > gpurun_out/fold_tests.log 2>&1; echo "linear tests exit $?"; tail -3 gpurun_out/fold_tests.log
timeout 900 python -m pytest tests/test_gpu_aff.py tests/test_gpu_dropin.py -m gpu -q -x > gpurun_out/fold_model_tests.log 2>&1; echo "model tests exit $?"; tail -3 gpurun_out/fold_model_tests.log
line() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['roofline']['kernel'], d['roofline']['frac'])"; }
for f in 1 0; do
CLUSTEN_LN_FOLD=$f timeout 600 python bench.py --no-cpu-baseline --no-extras 2>/dev/null | tee gpurun_out/fold_bench_mini_$f.json | line mini_fold_$f
CLUSTEN_LN_FOLD=$f timeout 600 python bench.py --no-cpu-baseline --no-extras --workload aff_small_fwd_b16_512 --steps 10 2>/dev/null | tee gpurun_out/fold_bench_small_$f.json | line small_fold_$f
done
for shape in "16384 256 768 bias ln" "4096 384 1152 bias ln" "16384 384 1152 bias ln" "262144 32 96 bias ln"; do
timeout 120 python tools/lin_profile.py $shape 2>&1 | tail -3
done > gpurun_out/fold_lin_prof.txt 2>&1
cat gpurun_out/fold_lin_prof.txt
