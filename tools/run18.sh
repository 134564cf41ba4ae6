mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_mini_v6.json 2> gpurun_out/bench_mini_v6.err; echo "bench exit $?"; cat gpurun_out/bench_mini_v6.json; tail -3 gpurun_out/bench_mini_v6.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_v6.json 2> gpurun_out/bench_ref_v6.err; echo "ref exit $?"; cat gpurun_out/bench_ref_v6.json
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2400 -c 2400 --csv --log-file gpurun_out/launches_mini_v6.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_fused_tile -s 30 -c 6 -o gpurun_out/r1_fusedfwd_mini_v6 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_fused.log 2>&1
echo "ncu fused exit $?"
for sh in small_s0 mini_s0 base_s0; do timeout 300 python benchmarks/op_bench.py --shape $sh --dtype bf16 > gpurun_out/op_${sh}_bf16_v6.log 2>&1; tail -12 gpurun_out/op_${sh}_bf16_v6.log; done
timeout 300 python benchmarks/op_bench.py --shape small_s0 --dtype f32 > gpurun_out/op_small_s0_f32_v6.log 2>&1; tail -12 gpurun_out/op_small_s0_f32_v6.log
