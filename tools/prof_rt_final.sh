#!/bin/bash
# ncu --set full of the final register-tile kernels at d = 8 / 16 / 32 (one filter + one smoother launch each, one full wave)
for cfg in "8 7104" "16 2664" "32 740"; do
  set -- $cfg; d=$1; B=$2
  CMD="python bench.py --workload c5 --state-dim $d --series $B --sub-batch $B --T 300 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
  $CMD > /dev/null 2>&1 || exit 1
  timeout 600 ncu --set full --clock-control none -k regex:rt_ -c 2 -f -o /tmp/rtf_d$d $CMD > gpurun_out/prtf_d$d.log 2>&1
  ncu -i /tmp/rtf_d$d.ncu-rep --page raw --csv > gpurun_out/rtf_d${d}_raw.csv
done
ls -la gpurun_out/rtf_d*_raw.csv
