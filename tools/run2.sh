mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_mini.json 2> gpurun_out/bench_mini.err; echo "bench exit $?"; cat gpurun_out/bench_mini.json; tail -5 gpurun_out/bench_mini.err
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_mini.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches exit $?"
