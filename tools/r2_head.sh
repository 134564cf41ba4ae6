mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "point_conv or pointconv or PointConv or msdeform or decoder or head or shepard" > gpurun_out/pytest_head.log 2>&1; echo "pytest head rc=$?"; tail -3 gpurun_out/pytest_head.log
timeout 600 python benchmarks/head_bench.py > gpurun_out/head_bench.json 2>&1; echo "head rc=$?"; tail -2 gpurun_out/head_bench.json
CLUSTEN_TCGEN05_LINEAR=0 timeout 600 python benchmarks/head_bench.py > gpurun_out/head_bench_cublas.json 2>&1; tail -1 gpurun_out/head_bench_cublas.json
