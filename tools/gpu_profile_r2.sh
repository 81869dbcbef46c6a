#!/bin/bash
# usage (under gpurun, ONE GPU): tools/gpu_profile_r2.sh
#   gpurun_out/r2_launches.csv          every launch of a short bench.py run with its device time
#   gpurun_out/r2_<kernel>.ncu-rep      --set full captures of every kernel family at its bench shape
mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on -f"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 1 --min-warm-s 0 --no-cpu-baseline > gpurun_out/r2_launches_bench.log 2>&1
$N -k regex:decode_step -s 1 -c 1 -o gpurun_out/r2_step_c2_independent python tools/ncu_step.py --shape 1,32,128,2048,32 --no-chain > /dev/null 2>&1
$N -k regex:decode_step -s 1 -c 1 -o gpurun_out/r2_step_c2_chained python tools/ncu_step.py --shape 1,32,128,2048,32 > /dev/null 2>&1
$N -k regex:decode_step -s 1 -c 1 -o gpurun_out/r2_step_c5_shard python tools/ncu_step.py --shape 8,5,128,16384,4 --no-chain > /dev/null 2>&1
$N -k regex:decode_attn_kernel -s 40 -c 1 -o gpurun_out/r2_decode_attn_c2 python tools/microbench.py --shape 1,32,128,2048,32 --iters 1 --no-graph > /dev/null 2>&1
$N -k regex:"q4|quant_merge" -s 8 -c 2 -o gpurun_out/r2_quant_c4 python tools/ncu_quant.py > /dev/null 2>&1
$N -k regex:local_v -s 2 -c 1 -o gpurun_out/r2_local_v python tools/microbench.py --extra --iters 2 > /dev/null 2>&1
$N -k regex:"topk|gather" -s 2 -c 2 -o gpurun_out/r2_prune python tools/prune_bench.py --reps 1 > /dev/null 2>&1
$N -k regex:"umma|softmax_rows|colsum" -s 3 -c 3 -o gpurun_out/r2_prefill python tools/microbench.py --prefill --iters 10 --prefill-s 4096 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep | awk '{print $5, $9}'
