mkdir -p gpurun_out
for cfg in "655 8" "656 8" "3276 4" "3280 4" "131 16" "16384 2"; do set -- $cfg; echo "== n=$1 heads=$2"; timeout 300 python benchmarks/attn_bench.py --n $1 --heads $2 2>&1 | tail -8; done > gpurun_out/attn_bench_v6.log 2>&1
cat gpurun_out/attn_bench_v6.log
