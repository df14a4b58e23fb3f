#!/bin/bash
# Round-end evidence run on ONE B200 (gpurun --timeout 1500 -- tools/final_profiles.sh [tag]): bench lines
# (our arm with all extras, the reference arm, the 512x512 configuration, the kernel-by-kernel launch
# mode), per-layer timings, the ncu launch list of one bench step (graph replay), ncu --set full of the
# convolution layers and of the HBM-bound kernels.  Outputs in gpurun_out/<tag>_*; copy the ones to
# keep into profiles/.
TAG=${1:-r02}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/$TAG
python bench.py --steps 20 --warmup 5 > ${O}_bench_n1.json 2> ${O}_bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > ${O}_bench_reference_arm.json 2> ${O}_bench_reference_arm.err; echo "reference arm rc=$?"
python bench.py --size 512 --batch 32 --steps 10 --warmup 5 --no-extras --no-cpu-baseline > ${O}_bench_512_n1.json 2> /dev/null
python bench.py --steps 20 --warmup 5 --no-graph --no-extras --no-cpu-baseline > ${O}_bench_n1_eager.json 2> /dev/null
python tools/layer_bench.py 64 256 10 > ${O}_layer_bench.txt 2>&1
# launch list of one timed bench step (the program has just exited 0 without ncu, above)
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file ${O}_launches_bench_step.csv \
    python bench.py --steps 1 --warmup 5 --no-extras --no-cpu-baseline --profiler-range > ${O}_ncu_step.log 2>&1
echo "launch list rc=$?"
python tools/launch_summary.py ${O}_launches_bench_step.csv 60 > ${O}_launches_bench_step_summary.txt 2>&1
python tools/profile_layers.py 64 > /dev/null 2>&1 && \
  ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"igemm|wgrad_v2" \
      -f -o ${O}_conv_layers python tools/profile_layers.py 64 > ${O}_ncu_layers.log 2>&1
python tools/profile_hbm.py 64 > /dev/null 2>&1 && \
  ncu --profile-from-start off --set full --clock-control none -k regex:"adain|conv_k27|maxpool|conv_last" \
      -f -o ${O}_hbm python tools/profile_hbm.py 64 > ${O}_ncu_hbm.log 2>&1
gzip -f ${O}_launches_bench_step.csv
ls -la ${O}_*
