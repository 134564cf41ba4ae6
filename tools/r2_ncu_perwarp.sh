mkdir -p gpurun_out
CMD="python benchmarks/attn_bench.py --n 16384 --heads 3 --c 32 --batch 16 --dtype bf16 --fwd-only --iters 2"
$CMD > gpurun_out/plain_a.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fused_tile -s 3 -c 1 -f -o gpurun_out/r2_perwarp_small_s0_bf16 $CMD > gpurun_out/ncu_a.log 2>&1; echo "ncu a $?"
CMD="python benchmarks/attn_bench.py --n 16384 --heads 2 --c 16 --batch 16 --dtype f32 --fwd-only --iters 2"
$CMD > gpurun_out/plain_b.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fused_tile -s 3 -c 1 -f -o gpurun_out/r2_perwarp_mini_s0_f32 $CMD > gpurun_out/ncu_b.log 2>&1; echo "ncu b $?"
