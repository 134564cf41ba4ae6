# ablations of linear_tc (CLUSTEN_TC_DBG bits: 1 no X loads, 2 no W loads, 4 no split, 8 no epilogue); results are wrong, times are what counts
mkdir -p gpurun_out
for res in 1 0; do for dbg in 0 4 8 12 15; do
echo "== resident=$res dbg=$dbg"; CLUSTEN_TC_RESIDENT=$res CLUSTEN_TC_DBG=$dbg timeout 300 python benchmarks/linear_bench.py --model mini --only 0:q+kv,0:fc1,1:q+kv,1:fc1,2:q+kv,2:fc1,3:q+kv,3:fc1 2>&1 | grep '"layer"' | python -c "
import sys,json
print('   ', ' | '.join(f\"{d['layer']} {d['K']}x{d['N']} {d['us_tc']}\" for d in map(json.loads, sys.stdin)))"
done; done
