mkdir -p gpurun_out
timeout 300 python benchmarks/op_bench.py --shape small_s0 --dtype bf16 --once > gpurun_out/once_bf16.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tile16_kernel" -c 6 -o gpurun_out/r1_tile_small_s0_bf16 -f python benchmarks/op_bench.py --shape small_s0 --dtype bf16 --once > gpurun_out/ncu_bf16.log 2>&1
echo "ncu exit $?"
timeout 300 python benchmarks/op_bench.py --shape small_s0 --dtype f32 --once > gpurun_out/once_f32.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tile32_kernel" -c 6 -o gpurun_out/r1_tile_small_s0_f32 -f python benchmarks/op_bench.py --shape small_s0 --dtype f32 --once > gpurun_out/ncu_f32.log 2>&1
echo "ncu exit $?"
timeout 600 python bench.py > gpurun_out/bench_mini_v1.json 2> gpurun_out/bench_mini_v1.err; cat gpurun_out/bench_mini_v1.json
