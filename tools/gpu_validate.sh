# Round-end validation on a B200 box (run through: gpurun --timeout 2400 -- 'bash tools/gpu_validate.sh').
# Everything lands in gpurun_out/; copy what should be judged into profiles/.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"; cat gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref exit $?"; cut -c1-300 gpurun_out/bench_reference.json
timeout 900 python bench.py --no-cpu-baseline --workload aff_tiny15_train_b32_512_bf16 --steps 5 --warmup 3 > gpurun_out/bench_tiny_train.json 2> gpurun_out/bench_tiny_train.err; echo "tiny exit $?"; cut -c1-300 gpurun_out/bench_tiny_train.json
# launch list of the default bench command (a window of ~2 passes: graph replays + the eager roofline pass)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 5000 -c 2400 --csv --log-file gpurun_out/launches_default.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches exit $?"
for sh in small_s0 mini_s0; do timeout 300 python benchmarks/op_bench.py --shape $sh --dtype bf16 > gpurun_out/op_${sh}_bf16_final.log 2>&1; done
tail -14 gpurun_out/op_small_s0_bf16_final.log
