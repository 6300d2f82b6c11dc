#!/bin/bash
# ncu --set full capture of the two kernels of the posterior-only call (second, warm invocation)
mkdir -p gpurun_out
python tools/prof_packed.py > gpurun_out/r2_prof_packed_plain.log 2>&1 || { tail -5 gpurun_out/r2_prof_packed_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:seq_ --launch-skip 2 -c 2 -f -o gpurun_out/r2_ncu_packed \
  python tools/prof_packed.py > gpurun_out/r2_prof_packed_ncu.log 2>&1
tail -3 gpurun_out/r2_prof_packed_ncu.log
ls -la gpurun_out/r2_ncu_packed.ncu-rep
