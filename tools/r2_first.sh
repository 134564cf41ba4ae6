# round 2, first GPU call: the TMA-staged fused forward -- parity, then timing against the per-warp kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "fused_attention_core" > gpurun_out/t1.log 2>&1; echo "t1 exit $?"; tail -12 gpurun_out/t1.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1; echo "all exit $?"; tail -8 gpurun_out/t_all.log
for cfg in "16384 3 32 16 small_s0" "16384 2 16 16 mini_s0" "4096 6 32 16 small_s1"; do set -- $cfg
 for dt in bf16 f32; do for tma in 1 0; do
  echo "== $5 $dt tma=$tma"; CLUSTEN_TMA_ATTN=$tma timeout 300 python benchmarks/attn_bench.py --n $1 --heads $2 --c $3 --batch $4 --dtype $dt --fwd-only 2>&1 | grep -E "clusten_attn_fwd|Error|error" | tee -a gpurun_out/attn_bench_r2_first.log
 done; done; done
CLUSTEN_INKERNEL_BIAS=1 timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "inkernel" > gpurun_out/t_pb.log 2>&1; echo "pb exit $?"; tail -5 gpurun_out/t_pb.log
CLUSTEN_TC_LINEAR=1 timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "linear_f32" > gpurun_out/t_lin.log 2>&1; echo "lin exit $?"; tail -5 gpurun_out/t_lin.log
