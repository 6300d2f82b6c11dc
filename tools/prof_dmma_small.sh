#!/bin/bash
# A-B timing of the DMMA products at d = 8 / 16 (PHYSS_RT_DMMA bit mask of padded dims; 56 = 8|16|32, 32 = d=32 only)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_grp.py -m gpu -x -q > gpurun_out/r2_t_dmma2.log 2>&1
PHYSS_RT_DMMA=56 python -m pytest tests/test_gpu_grp.py tests/test_gpu_pscan.py -m gpu -x -q >> gpurun_out/r2_t_dmma2.log 2>&1; tail -3 gpurun_out/r2_t_dmma2.log
for cfg in "8 14208 7104" "16 5328 2664"; do
  set -- $cfg
  for v in 56 32; do
    PHYSS_RT_DMMA=$v python bench.py --workload c5 --state-dim $1 --series $2 --sub-batch $3 --no-sweep --no-e2e --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/r2_d$1_dmma$v.json 2>> gpurun_out/r2_dsmall.err
  done
done
(time python bench.py --steps 10 --warmup 3) > gpurun_out/r2_b7.json 2> gpurun_out/r2_b7.err; tail -c 400 gpurun_out/r2_b7.err
