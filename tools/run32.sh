mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_mini_v10.json 2> gpurun_out/bench_mini_v10.err; echo "bench exit $?"; cat gpurun_out/bench_mini_v10.json; tail -3 gpurun_out/bench_mini_v10.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_v10.json 2> gpurun_out/bench_ref_v10.err; echo "ref exit $?"; cut -c1-300 gpurun_out/bench_ref_v10.json
# launch list of the bench command (graph replays + the eager roofline pass), a window of ~2 passes
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 2600 --csv --log-file gpurun_out/launches_mini_v10.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches exit $?"
# DRAM traffic of the dominant kernel: every launch of one forward (12 blocks)
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"attn_fused_tile_kernel" -c 12 --csv --log-file gpurun_out/traffic_mini_attn_fwd_v10.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_traffic.log 2>&1
echo "ncu traffic exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"attn_fused_tile_kernel" -c 4 -o gpurun_out/r1_fusedfwd_mini_v10 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_fused.log 2>&1
echo "ncu fused exit $?"
for sh in small_s0 mini_s0 base_s0; do timeout 300 python benchmarks/op_bench.py --shape $sh --dtype bf16 > gpurun_out/op_${sh}_bf16_v13.log 2>&1; done
timeout 300 python benchmarks/op_bench.py --shape small_s0 --dtype f32 --ref > gpurun_out/op_small_s0_f32_v13.log 2>&1
timeout 300 python benchmarks/op_bench.py --shape cfg1 --dtype f32 --ref > gpurun_out/op_cfg1_f32_v13.log 2>&1
tail -14 gpurun_out/op_small_s0_bf16_v13.log
