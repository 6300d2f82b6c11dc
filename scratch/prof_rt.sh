#!/bin/bash
# ncu --set full of the register-tile kernels (one filter + one smoother launch, short T); CSV pages exported on the box
set -x
for d in "$@"; do
  B=$((16384/d))
  CMD="python bench.py --workload c5 --state-dim $d --series $B --sub-batch $B --T 300 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
  $CMD > gpurun_out/prt_d$d.json 2> gpurun_out/prt_d$d.err || exit 1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:rt_ -c 2 -f -o /tmp/rt_d$d $CMD > gpurun_out/prt_ncu_d$d.log 2>&1
  ncu -i /tmp/rt_d$d.ncu-rep --page raw --csv > gpurun_out/rt_d${d}_raw.csv
  ncu -i /tmp/rt_d$d.ncu-rep --page source --csv --print-source sass,cuda > gpurun_out/rt_d${d}_source.csv 2>/dev/null || \
  ncu -i /tmp/rt_d$d.ncu-rep --page source --csv > gpurun_out/rt_d${d}_source.csv
done
ls -la gpurun_out/rt_d*
