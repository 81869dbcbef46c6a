#!/bin/bash
# usage (under gpurun --gpus N): tools/run_scaling.sh N [steps]   -> gpurun_out/bench_nN.json (+ .log)
N=${1:-1}; STEPS=${2:-300}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 300 python bench.py --steps $STEPS --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.log
else
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
      bench.py --gpus $N --steps $STEPS --warmup 5 > gpurun_out/bench_n$N.log 2>&1
  grep '^{"metric"' gpurun_out/bench_n$N.log | tail -1 > gpurun_out/bench_n$N.json
fi
python - <<PY
import json
d=json.load(open("gpurun_out/bench_n$N.json"))
print({k:(round(v,2) if isinstance(v,float) else v) for k,v in d.items() if k in ("value","n_gpus","ms_per_step","cuda_graph","collective")}, "roofline.frac", round(d["roofline"]["frac"],3), "share", round(d["roofline"]["share_of_step"],3), "e2e", round(d["e2e"]["value"],1))
PY
