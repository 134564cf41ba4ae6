mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_aff.py -m gpu -q -x -k "table or aff or fused" > gpurun_out/pytest_tab.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_tab.log; tail -4 gpurun_out/pytest_tab.log
timeout 600 python bench.py --no-cpu-baseline --workload aff_tiny15_train_b32_512_bf16 --steps 5 --warmup 3 > gpurun_out/bench_tiny_v6.json 2> gpurun_out/bench_tiny_v6.err; cat gpurun_out/bench_tiny_v6.json; tail -3 gpurun_out/bench_tiny_v6.err
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -s 10000 -c 16000 --csv --log-file gpurun_out/launches_tiny_train_v6.csv python bench.py --no-cpu-baseline --workload aff_tiny15_train_b32_512_bf16 --steps 1 --warmup 3 > gpurun_out/ncu_launch_tiny.log 2>&1
echo "ncu launches exit $?"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"attn_bwd_tile_kernel|scat2_kernel16|attn_fused_tile_kernel" -s 111 -c 9 -o gpurun_out/r1_fusedbwd_tiny_v6 -f python bench.py --no-cpu-baseline --workload aff_tiny15_train_b32_512_bf16 --steps 1 --warmup 3 > gpurun_out/ncu_fusedbwd.log 2>&1
echo "ncu fusedbwd exit $?"
