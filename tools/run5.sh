mkdir -p gpurun_out
timeout 300 python benchmarks/op_bench.py --shape small_s0 --dtype bf16 > gpurun_out/op_small_bf16.log 2>&1; cat gpurun_out/op_small_bf16.log
timeout 300 python benchmarks/op_bench.py --shape small_s0 --dtype f32 > gpurun_out/op_small_f32.log 2>&1; cat gpurun_out/op_small_f32.log
timeout 300 python benchmarks/op_bench.py --shape mini_s0 --dtype f32 > gpurun_out/op_mini_f32.log 2>&1; cat gpurun_out/op_mini_f32.log
timeout 300 python benchmarks/op_bench.py --shape base_s0 --dtype bf16 > gpurun_out/op_base_bf16.log 2>&1; cat gpurun_out/op_base_bf16.log
