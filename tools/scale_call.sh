#!/bin/bash
# One multi-GPU gpurun call: headline bench (graph replay), the two ablations the scaling analysis
# needs (no all-reduce; kernel-by-kernel launch), the 512x512 / batch 32 configuration, dp_check.
# Usage: gpurun --gpus N --timeout 900 -- tools/scale_call.sh N [quick]     -> gpurun_out/scale_nN_*.json
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
[ "$N" = 1 ] && T="python"
COMMON="--gpus $N --steps 20 --warmup 5 --no-extras --no-cpu-baseline"
run() {  # name, extra flags
  timeout 150 $T bench.py $COMMON $2 > gpurun_out/scale_n${N}_$1.json 2> gpurun_out/scale_n${N}_$1.err
  echo "$1 rc=$? $(python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/scale_n${N}_$1.json')); print(round(d['value'],1), 'img/s', round(d['ms_per_step'],3), 'ms', d['per_rank'])
except Exception as e: print('no line', e)")"
}
run graph ""
if [ "$2" = ctas ]; then  # how many SMs may NCCL take away from the persistent convolution kernels?
  NCCL_MAX_CTAS=2 run graph_ctas2 ""
  NCCL_MAX_CTAS=8 run graph_ctas8 ""
elif [ "$2" != quick ]; then
  run noallreduce "--no-allreduce"
  run eager "--no-graph"
fi
run 512 "--size 512 --batch 32"
if [ "$N" != 1 ]; then
  timeout 120 $T tools/dp_check.py 2>&1 | grep -E "dp_check world" | tee gpurun_out/dp_check_n${N}.txt
fi
