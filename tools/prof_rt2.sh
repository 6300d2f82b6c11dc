#!/bin/bash
# two-kernel smoother: parity tests, then A-B timing against the one-kernel smoother at d = 8 / 16 / 32
mkdir -p gpurun_out
python -m pytest tests/test_gpu_grp.py -m gpu -x -q > gpurun_out/r2_t_rt2.log 2>&1; tail -15 gpurun_out/r2_t_rt2.log
for cfg in "8 14208 7104" "16 5328 2664" "32 1480 740"; do
  set -- $cfg
  for v in 1 0; do
    PHYSS_TWO_KERNEL=$v python bench.py --workload c5 --state-dim $1 --series $2 --sub-batch $3 --no-sweep --no-e2e --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/r2_rt2_d$1_v$v.json 2>> gpurun_out/r2_rt2.err
  done
done
tail -c 600 gpurun_out/r2_rt2.err
