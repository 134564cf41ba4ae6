mkdir -p gpurun_out
timeout 300 python benchmarks/linear_bench.py --model mini --chain 0 > gpurun_out/lin_mini.log 2>&1; echo "mini rc=$?"; tail -20 gpurun_out/lin_mini.log | cut -c1-420
timeout 300 python benchmarks/linear_bench.py --model small --chain 0 > gpurun_out/lin_small.log 2>&1; echo "small rc=$?"; tail -20 gpurun_out/lin_small.log | cut -c1-420
