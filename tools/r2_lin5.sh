mkdir -p gpurun_out
timeout 300 python tools/lin_race.py 16384 256 768 bias > gpurun_out/race1.log 2>&1; tail -40 gpurun_out/race1.log
timeout 300 python tools/lin_race.py 4096 768 2304 bias > gpurun_out/race2.log 2>&1; tail -30 gpurun_out/race2.log
