# DRAM traffic per launch of the dominant entry point (clusten_linear_tc_f32): the 49 linear_tc_kernel launches of one eager AFF-Mini forward
mkdir -p gpurun_out
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:linear_tc_kernel -s 49 -c 49 --csv --log-file gpurun_out/traffic_linear_tc.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/traffic_linear_tc.log 2>&1; echo "ncu exit $?"; wc -l gpurun_out/traffic_linear_tc.csv
