mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -k "msdetrpc or point_conv or graphed or fused" > gpurun_out/pytest_new.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_new.log; tail -8 gpurun_out/pytest_new.log
for cfg in "16384 2" "655 8"; do set -- $cfg; echo "== n=$1 heads=$2"; timeout 300 python benchmarks/attn_bench.py --n $1 --heads $2 2>&1 | tail -5; done > gpurun_out/attn_bench_v8.log 2>&1
cat gpurun_out/attn_bench_v8.log
timeout 900 python bench.py --no-cpu-baseline --workload aff_tiny15_train_b32_512_bf16 --steps 5 --warmup 3 > gpurun_out/bench_tiny_v10.json 2> gpurun_out/bench_tiny_v10.err; echo "tiny exit $?"; cat gpurun_out/bench_tiny_v10.json; tail -5 gpurun_out/bench_tiny_v10.err
