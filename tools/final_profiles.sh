#!/bin/bash
# Round-end evidence run on the GPU box: bench lines, launch list of one bench step, ncu --set full
# metrics (as CSV, the .ncu-rep files stay on the box), per-layer timings.  Outputs in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/r01_bench_n1.json 2> gpurun_out/r01_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01_bench_reference_arm.json 2> gpurun_out/r01_bench_reference_arm.err
python bench.py --size 512 --batch 32 --steps 5 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r01_bench_512.json 2> gpurun_out/r01_bench_512.err
python tools/layer_bench.py 64 256 10 > gpurun_out/r01_layer_bench.txt 2>&1
python bench.py --steps 1 --warmup 5 --no-extras --no-cpu-baseline --profiler-range > /dev/null 2>&1 && \
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/r01_launches_bench_step.csv \
      python bench.py --steps 1 --warmup 5 --no-extras --no-cpu-baseline --profiler-range > gpurun_out/ncu_step.log 2>&1
python tools/profile_layers.py 64 > /dev/null 2>&1 && \
  ncu --profile-from-start off --set full --clock-control none -k regex:"igemm|wgrad_v2" -o /tmp/layers \
      python tools/profile_layers.py 64 > gpurun_out/ncu_layers.log 2>&1 && \
  ncu -i /tmp/layers.ncu-rep --page raw --csv > gpurun_out/r01_ncu_conv_layers_raw.csv
python tools/profile_hbm.py 64 > /dev/null 2>&1 && \
  ncu --profile-from-start off --set full --clock-control none -o /tmp/hbm \
      python tools/profile_hbm.py 64 > gpurun_out/ncu_hbm.log 2>&1 && \
  ncu -i /tmp/hbm.ncu-rep --page raw --csv > gpurun_out/r01_ncu_hbm_raw.csv
ls -la gpurun_out/r01_*
