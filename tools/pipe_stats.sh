#!/bin/bash
# Builds a DEBUG copy of the library with -DWU_PIPE_STATS (cycle counters around every barrier wait of
# the MMA-issuing and TMA-producing threads of the 3x3 kernels) into tools/scratch/ and prints, per
# layer, where those two threads spend their time.  Usage (GPU box): tools/pipe_stats.sh
cd "$(dirname "$0")/.."
set -e
OUT=tools/scratch/libwu_b200_stats.so
if [ ! -f $OUT ] || [ weather-unet_b200/csrc/wu_conv3x3.cu -nt $OUT ]; then
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -DWU_PIPE_STATS -shared -cudart shared \
      -o $OUT weather-unet_b200/csrc/*.cu 2>&1 | grep -v deprecated || true
fi
python tools/pipe_stats.py "$@"
