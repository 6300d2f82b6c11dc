#!/bin/bash
CMD="python bench.py --workload grad --series 32768 --T 400 --steps 1 --warmup 1"
$CMD > /dev/null 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kf_vjp -c 1 -f -o /tmp/vjp $CMD > gpurun_out/prof_vjp.log 2>&1
ncu -i /tmp/vjp.ncu-rep --page raw --csv > gpurun_out/vjp_raw.csv
ls -la gpurun_out/vjp_raw.csv
