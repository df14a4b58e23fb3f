#!/bin/bash
# Runs each GPU test group in its own process (a trapped kernel kills only its group) and keeps
# full logs under gpurun_out/.  Usage: tools/gpu_check.sh [group ...]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
groups=("$@")
if [ ${#groups[@]} -eq 0 ]; then
  groups=(test_kernels_gpu.py::test_conv3x3_fprop test_kernels_gpu.py::test_conv3x3_dgrad
          test_kernels_gpu.py::test_conv3x3_wgrad test_kernels_gpu.py::test_conv_first
          test_kernels_gpu.py::test_conv_last_tanh test_kernels_gpu.py::test_maxpool
          test_kernels_gpu.py::test_adain_up_drop test_kernels_gpu.py::test_dropout_philox_rate
          test_kernels_gpu.py::test_layout_roundtrip test_kernels_gpu.py::test_errors
          test_generator_gpu.py)
fi
rc=0
for g in "${groups[@]}"; do
  name=$(echo "$g" | tr ':/.' '___')
  timeout 600 python -m pytest "tests/$g" -q -m gpu -x --tb=short -p no:cacheprovider \
      > "gpurun_out/check_$name.log" 2>&1
  r=$?
  echo "== $g -> exit $r: $(tail -n 1 gpurun_out/check_$name.log)"
  [ $r -ne 0 ] && { rc=1; grep -E "^(FAILED|ERROR|E  )" "gpurun_out/check_$name.log" | head -n 12; }
done
exit $rc
