# Cycle counters of the tcgen05 Linear on the block's shapes.  The profiling copy of the library is built in the authoring container:
#   cd autofocusformermod_b200/csrc && nvcc <the flags of _build.py> -DCLUSTEN_TC_PROFILE -c linear_tc.cu -o /tmp/lt_prof.o
#   cd .. && nvcc -shared -o libclusten_b200_prof.so $(ls build/*.o | grep -v linear_tc.o) /tmp/lt_prof.o -gencode arch=compute_100a,code=sm_100a
mkdir -p gpurun_out
for res in 1 0; do for dbg in 0 15; do
echo "=== resident=$res dbg=$dbg"
for shape in "16384 256 768 bias ln" "16384 512 256 residual" "65536 128 384 bias ln" "262144 32 96 bias ln" "4096 384 1152 bias ln"; do
CLUSTEN_TC_RESIDENT=$res CLUSTEN_TC_DBG=$dbg timeout 120 python tools/lin_profile.py $shape 2>&1 | tail -3
done; done; done > gpurun_out/lin_prof.txt 2>&1
cat gpurun_out/lin_prof.txt
