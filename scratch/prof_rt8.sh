#!/bin/bash
set -x

CMD="python bench.py --workload c5 --state-dim 8 --series 7104 --sub-batch 7104 --T 300 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/prt_d8.json 2> gpurun_out/prt_d8.err || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rt_ -c 2 -f -o /tmp/rt_d8 $CMD > gpurun_out/prt_ncu_d8.log 2>&1
ncu -i /tmp/rt_d8.ncu-rep --page raw --csv > gpurun_out/rt_d8_raw.csv
ncu -i /tmp/rt_d8.ncu-rep --page source --csv --print-source sass,cuda > gpurun_out/rt_d8_source.csv 2>/dev/null || ncu -i /tmp/rt_d8.ncu-rep --page source --csv > gpurun_out/rt_d8_source.csv
ls -la gpurun_out/rt_d8*
