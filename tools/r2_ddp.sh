# data-parallel training with CUDA graphs + one flat NCCL all-reduce: N GPUs of one box
N=$1; mkdir -p gpurun_out
for wl in aff_tiny15_train_b32_512_bf16 aff_base_train_b2_512x1024_bf16; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 10 --warmup 3 --workload $wl --no-cpu-baseline > gpurun_out/r2_bench_${wl}_n$N.json 2> gpurun_out/r2_bench_${wl}_n$N.err
  echo "$wl N=$N exit $?"; cut -c1-700 gpurun_out/r2_bench_${wl}_n$N.json; tail -3 gpurun_out/r2_bench_${wl}_n$N.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2_bench_default_n$N.json 2> gpurun_out/r2_bench_default_n$N.err; echo "default N=$N exit $?"; cut -c1-600 gpurun_out/r2_bench_default_n$N.json
