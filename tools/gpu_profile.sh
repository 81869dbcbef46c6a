#!/bin/bash
# usage (under gpurun, ONE GPU): tools/gpu_profile.sh <tag>
#   gpurun_out/<tag>_launches.csv   every launch of a short bench.py run with its device time
#   gpurun_out/<tag>.ncu-rep        --set full capture of the decode kernel at the bench shapes
TAG=${1:-prof}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --min-warm-s 0 --no-cpu-baseline > gpurun_out/${TAG}_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:decode_attn -s 40 -c 2 \
    -o gpurun_out/$TAG -f python tools/microbench.py --shape 1,32,128,2048,32 --iters 1 --no-graph > gpurun_out/$TAG.ncu.log 2>&1
tail -2 gpurun_out/$TAG.ncu.log
