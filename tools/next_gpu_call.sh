#!/bin/bash
# Everything that was written after round 1's GPU budget ran out and still has to meet the hardware,
# in one gpurun call (about three minutes):
#   1. the UMMA operand-start probe, K-major (already green) and MN-major (weight gradients) halves;
#   2. the opt-in one-box convolution kernels (WU_CONV_IMPL=3 / 4): parity, per-layer timings, bench step;
#   3. the AdaIN kernel timings of the current build.
# Usage: gpurun --timeout 400 -- tools/next_gpu_call.sh      (outputs in gpurun_out/)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
if [ ! -x tools/scratch/umma_unaligned_probe ]; then
  nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -o tools/scratch/umma_unaligned_probe \
      tools/scratch/umma_unaligned_probe.cu weather-unet_b200/csrc/wu_host.cu -lcuda
fi
timeout 30 tools/scratch/umma_unaligned_probe > gpurun_out/umma_probe.txt 2>&1
grep RESULT gpurun_out/umma_probe.txt
tools/check_conv_impl.sh
timeout 120 python tools/time_adain.py | tail -n 1
