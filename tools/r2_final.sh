# Round-2 final validation on one B200 (gpurun --timeout 2400 -- 'bash tools/r2_final.sh'); everything lands in gpurun_out/final_*
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/final_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/final_pytest_gpu.log; tail -3 gpurun_out/final_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/final_smoke.log; tail -2 gpurun_out/final_smoke.log
timeout 900 python bench.py > gpurun_out/final_bench_default.json 2> gpurun_out/final_bench_default.err; echo "bench exit $?"; cut -c1-400 gpurun_out/final_bench_default.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; echo "ref exit $?"; cut -c1-300 gpurun_out/final_bench_reference.json
timeout 600 python bench.py --no-cpu-baseline --no-extras --workload aff_small_fwd_b16_512 --steps 10 > gpurun_out/final_bench_small.json 2>/dev/null; echo "small exit $?"; cut -c1-200 gpurun_out/final_bench_small.json
timeout 900 python bench.py --no-cpu-baseline --no-extras --workload aff_tiny15_train_b32_512_bf16 --steps 5 --warmup 3 > gpurun_out/final_bench_tiny_train.json 2> gpurun_out/final_bench_tiny_train.err; echo "tiny exit $?"; cut -c1-200 gpurun_out/final_bench_tiny_train.json
timeout 900 python bench.py --no-cpu-baseline --no-extras --workload aff_base_train_b2_512x1024_bf16 --steps 5 --warmup 3 > gpurun_out/final_bench_base_train.json 2> gpurun_out/final_bench_base_train.err; echo "base exit $?"; cut -c1-200 gpurun_out/final_bench_base_train.json
timeout 600 python benchmarks/profile_step.py --workload aff_mini_fwd_b16_512 --rows 60 > gpurun_out/final_profile_mini.txt 2>&1
timeout 600 python benchmarks/profile_step.py --workload aff_small_fwd_b16_512 --rows 40 > gpurun_out/final_profile_small.txt 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/final_launches_default.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/final_ncu_launch.log 2>&1; echo "ncu launches exit $?"
