# A/B of the opt-in kernels on a B200 box:  gpurun --timeout 1200 -- 'bash tools/ab_optins.sh'
# 1. parity of the opt-in kernels, 2. model-level parity with them switched on, 3. bench lines with / without each switch.
mkdir -p gpurun_out
CLUSTEN_INKERNEL_BIAS=1 CLUSTEN_TC_LINEAR=1 timeout 300 python -m pytest tests/test_gpu_ops.py -q -k "inkernel or linear_f32" > gpurun_out/optin_ops.log 2>&1; tail -3 gpurun_out/optin_ops.log
CLUSTEN_INKERNEL_BIAS=1 CLUSTEN_TC_LINEAR=1 timeout 300 python -m pytest tests/test_gpu_aff.py -q > gpurun_out/optin_aff.log 2>&1; tail -3 gpurun_out/optin_aff.log
run() {   # name, env assignments..., then "--", then bench arguments
    name=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
    env "${envs[@]}" timeout 200 python bench.py --no-cpu-baseline "$@" > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
    echo "$name: $(cut -c1-140 gpurun_out/ab_$name.json)"
}
run mini_base X=0 --
run mini_posbias CLUSTEN_INKERNEL_BIAS=1 --
run mini_tclinear CLUSTEN_TC_LINEAR=1 --
run mini_both CLUSTEN_INKERNEL_BIAS=1 CLUSTEN_TC_LINEAR=1 --
S="--workload aff_small_fwd_b16_512 --steps 10 --warmup 3"     # the north-star scaling model: 80 % of its step is fp32 sgemm
run small_base X=0 -- $S
run small_tclinear CLUSTEN_TC_LINEAR=1 -- $S
T="--workload aff_tiny15_train_b32_512_bf16 --steps 5 --warmup 3"
run tiny_base X=0 -- $T
run tiny_posbias CLUSTEN_INKERNEL_BIAS=1 -- $T
run tiny_mergewf CLUSTEN_MERGE_WF_AUTOCAST=1 -- $T
run tiny_both CLUSTEN_INKERNEL_BIAS=1 CLUSTEN_MERGE_WF_AUTOCAST=1 -- $T
for v in "" "--inkernel-bias"; do
    timeout 100 python benchmarks/attn_bench.py --n 16384 --heads 2 --c 32 --batch 32 --dtype bf16 $v > gpurun_out/ab_attn_bf16_s0$v.log 2>&1
    timeout 100 python benchmarks/attn_bench.py --n 16384 --heads 2 --c 16 --batch 16 --dtype f32 $v > gpurun_out/ab_attn_f32_s0$v.log 2>&1
    grep -h "entry" gpurun_out/ab_attn_bf16_s0$v.log gpurun_out/ab_attn_f32_s0$v.log | cut -c1-160
done
