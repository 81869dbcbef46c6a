#!/bin/bash
# usage (under gpurun): tools/gpu_profile.sh <tag>   -> gpurun_out/<tag>.ncu-rep + launches csv
TAG=${1:-prof}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:decode_attn -s 40 -c 2 \
    -o gpurun_out/$TAG -f python tools/microbench.py --shape 1,32,128,2048,32 --iters 1 --no-graph > gpurun_out/$TAG.ncu.log 2>&1
tail -3 gpurun_out/$TAG.ncu.log
