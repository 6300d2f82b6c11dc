#!/bin/bash
# Round-2 ncu evidence for the headline d = 4 kernels (after the bench command itself has exited 0 without ncu):
#   1. launch list of the default bench (kernel share of the step)
#   2. one --set full capture of the smoother and the filter at the bench sub-batch (32,768 series, short T)
mkdir -p gpurun_out
B="python bench.py --no-sweep --no-e2e --no-cpu-baseline"
$B --steps 2 --warmup 3 > gpurun_out/r2_prof_plain.json 2> gpurun_out/r2_prof.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_c5d4.csv \
  $B --steps 2 --warmup 3 > gpurun_out/r2_prof_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:seq_ -c 2 -f -o gpurun_out/r2_ncu_c5d4 \
  $B --series 32768 --T 600 --steps 1 --warmup 1 > gpurun_out/r2_prof_ncu2.log 2>&1
tail -c 300 gpurun_out/r2_prof.err
# 3. launch list of the CVI step (metric 2)
python bench.py --workload cvi --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_prof_cvi_plain.json 2>> gpurun_out/r2_prof.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_cvi.csv \
  python bench.py --workload cvi --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r2_prof_ncu3.log 2>&1
