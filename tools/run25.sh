mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "wf or weighted" > gpurun_out/pytest_wf.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_wf.log; tail -5 gpurun_out/pytest_wf.log
for sh in small_s0 mini_s0 base_s0 small_s1; do timeout 300 python benchmarks/op_bench.py --shape $sh --dtype bf16 > gpurun_out/op_${sh}_bf16_v12.log 2>&1; tail -7 gpurun_out/op_${sh}_bf16_v12.log | grep wf; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_opbench_small_s0_v12.csv python benchmarks/op_bench.py --shape small_s0 --dtype bf16 --once > gpurun_out/ncu_op_once.log 2>&1
