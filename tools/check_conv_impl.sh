#!/bin/bash
# Validate and time the experimental one-box A loading of the 3x3 convolution kernels
# (WU_CONV_IMPL=3 / 4, csrc/wu_conv3x3.cu ConvCfg2<.., ABOX = 1>) against the default kernels on the
# GPU box: parity tests of fprop / dgrad under each setting, then the per-layer CUDA-event timings.
# Usage (from the repo root, on a B200): tools/check_conv_impl.sh        -> gpurun_out/conv_impl_*.txt
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for impl in 0 3 4; do
  echo "== WU_CONV_IMPL=$impl"
  WU_CONV_IMPL=$impl timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x --tb=short \
      -p no:cacheprovider -k "conv3x3_fprop or conv3x3_dgrad or pool_fused or last_fused" > gpurun_out/conv_impl_${impl}_tests.txt 2>&1
  echo "   tests: exit $? ($(tail -n 1 gpurun_out/conv_impl_${impl}_tests.txt))"
  WU_CONV_IMPL=$impl timeout 300 python tools/layer_bench.py 64 256 10 > gpurun_out/conv_impl_${impl}_layers.txt 2>&1
  grep -E "TOTAL (fprop|dgrad)" gpurun_out/conv_impl_${impl}_layers.txt
done
# whole iteration under the T = 4 variant (fused last / pool layers included) next to the default
for impl in 0 3; do
  WU_CONV_IMPL=$impl timeout 300 python bench.py --steps 10 --warmup 5 --no-extras --no-cpu-baseline \
      2> /dev/null | python -c "import sys, json; d = json.loads(sys.stdin.read()); print('WU_CONV_IMPL=$impl', d['value'], 'images/s', d['ms_per_step'], 'ms')"
done
