mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "^test|passed|failed|FAILED|Error|rel err|\{'|res2" | tail -40
timeout 900 python bench.py > gpurun_out/bench_default_r2a.json 2> gpurun_out/bench_default_r2a.err; echo "bench exit $?"; cat gpurun_out/bench_default_r2a.json; tail -5 gpurun_out/bench_default_r2a.err
