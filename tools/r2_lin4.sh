mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_aff.py tests/test_gpu_dropin.py -x -q -m gpu > gpurun_out/pytest_aff.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_aff.log
timeout 600 python bench.py --no-extras > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/bench_default.json
CLUSTEN_TCGEN05_LINEAR=0 timeout 600 python bench.py --no-extras > gpurun_out/bench_default_cublas.json 2> gpurun_out/bench_default_cublas.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_default_cublas.json
