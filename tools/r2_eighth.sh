mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6
for sh in small_s0 base_s0; do for dt in f32 bf16; do echo "== $sh $dt"; timeout 300 python benchmarks/op_bench.py --shape $sh --dtype $dt --iters 5 2>&1 | grep -E '"op"' | cut -c1-150 | tee -a gpurun_out/r2_opbench_${sh}_${dt}_v20.jsonl; done; done
