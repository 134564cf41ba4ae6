mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_aff.py -m gpu -q -x -k "table or aff or golden" > gpurun_out/pytest_tab.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_tab.log; tail -8 gpurun_out/pytest_tab.log
timeout 600 python bench.py --no-cpu-baseline --workload aff_tiny15_train_b32_512_bf16 --steps 3 --warmup 3 > gpurun_out/bench_tiny_v4.json 2> gpurun_out/bench_tiny_v4.err; cat gpurun_out/bench_tiny_v4.json; tail -3 gpurun_out/bench_tiny_v4.err
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -s 12000 -c 14000 --csv --log-file gpurun_out/launches_tiny_train_v4.csv python bench.py --no-cpu-baseline --workload aff_tiny15_train_b32_512_bf16 --steps 1 --warmup 3 > gpurun_out/ncu_launch_tiny.log 2>&1
echo "ncu launches exit $?"
