#!/bin/bash
# usage (under gpurun, ONE GPU): tools/gpu_profile_r2b.sh  -- second profiling pass of round 2 (kernels added late in the round)
#   gpurun_out/r2b_launches.csv            every launch of a short bench.py run with its device time
#   gpurun_out/r2b_prefill_fused.ncu-rep   prefill_fused_kernel at S = 8192 (BASELINE config 3's prefill, one 7B layer)
#   gpurun_out/r2b_quant_c4.ncu-rep        decode_attn_q4_kernel (mma.sync tiles + ticket merge) at config 4
mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on -f"
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2b_launches.csv \
    python bench.py --steps 2 --warmup 1 --min-warm-s 0 --no-cpu-baseline > gpurun_out/r2b_launches_bench.log 2>&1
$N -k regex:"prefill_fused|prefill_pack|mask_classify" -s 3 -c 3 -o gpurun_out/r2b_prefill_fused python tools/microbench.py --prefill --iters 20 --prefill-s 8192 > /dev/null 2>&1
$N -k regex:"q4" -s 8 -c 1 -o gpurun_out/r2b_quant_c4 python tools/ncu_quant.py > /dev/null 2>&1
ls -la gpurun_out/r2b_* | awk '{print $5, $9}'
