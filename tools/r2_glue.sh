# Round-2 late session: row gather / on-chip presence marks / fused stem / LayerNorm rows-in-flight -- parity, full GPU suite, A/B.
# Run through: gpurun --timeout 1500 -- 'bash tools/r2_glue.sh'
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "gather_rows or stem_conv or stage_prepare or layer_norm or ln_" > gpurun_out/glue_new_tests.log 2>&1; echo "new tests exit $?"; tail -3 gpurun_out/glue_new_tests.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/glue_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/glue_pytest_gpu.log
line() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['roofline']['kernel'], d['roofline']['frac'])"; }
timeout 600 python bench.py --no-cpu-baseline --no-extras 2>gpurun_out/glue_bench_all.err | tee gpurun_out/glue_bench_all.json | line all_on
CLUSTEN_GATHER_ROWS=0 timeout 600 python bench.py --no-cpu-baseline --no-extras 2>/dev/null | tee gpurun_out/glue_bench_nogather.json | line gather_off
CLUSTEN_PREPARE_BITMAP=0 timeout 600 python bench.py --no-cpu-baseline --no-extras 2>/dev/null | tee gpurun_out/glue_bench_nobitmap.json | line bitmap_off
CLUSTEN_FUSED_STEM=0 timeout 600 python bench.py --no-cpu-baseline --no-extras 2>/dev/null | tee gpurun_out/glue_bench_nostem.json | line stem_off
timeout 600 python benchmarks/profile_step.py --workload aff_mini_fwd_b16_512 --rows 70 > gpurun_out/glue_profile_mini.txt 2>&1; head -40 gpurun_out/glue_profile_mini.txt
timeout 300 python benchmarks/int_bench.py > gpurun_out/glue_int_bench.txt 2>&1; tail -12 gpurun_out/glue_int_bench.txt
