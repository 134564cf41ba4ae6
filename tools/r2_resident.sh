# resident-row tile walk of linear_tc: parity (bit-equal to the streamed walk), A/B on the default line and the north-star model
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "linear_tc" > gpurun_out/res_tests.log 2>&1; echo "linear tests exit $?"; tail -3 gpurun_out/res_tests.log
line() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['roofline']['kernel'], d['roofline']['frac'])"; }
for r in 1 0; do
CLUSTEN_TC_RESIDENT=$r timeout 600 python bench.py --no-cpu-baseline --no-extras 2>/dev/null | tee gpurun_out/res_bench_mini_$r.json | line mini_resident_$r
CLUSTEN_TC_RESIDENT=$r timeout 600 python bench.py --no-cpu-baseline --no-extras --workload aff_small_fwd_b16_512 --steps 10 2>/dev/null | tee gpurun_out/res_bench_small_$r.json | line small_resident_$r
CLUSTEN_TC_RESIDENT=$r timeout 600 python benchmarks/linear_bench.py --model mini > gpurun_out/res_linear_bench_mini_$r.jsonl 2>&1; tail -1 gpurun_out/res_linear_bench_mini_$r.jsonl
done
