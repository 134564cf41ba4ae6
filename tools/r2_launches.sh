mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-extras"
$CMD > gpurun_out/b_plain.json 2> gpurun_out/b_plain.err && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_b.log 2>&1; echo "ncu rc=$?"
