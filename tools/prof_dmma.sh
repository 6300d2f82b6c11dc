#!/bin/bash
# DMMA experiment at d = 32: A-B timing of the smoother with the two dense products on the
# FP64 tensor cores (PHYSS_RT_DMMA=1, default) vs DFMA (PHYSS_RT_DMMA=0), then one ncu --set full capture of each.
mkdir -p gpurun_out
python -m pytest tests/test_gpu_grp.py -m gpu -x -q > gpurun_out/r2_t_dmma.log 2>&1; tail -3 gpurun_out/r2_t_dmma.log
B="python bench.py --workload c5 --state-dim 32 --series 1480 --sub-batch 740 --no-sweep --no-e2e --no-cpu-baseline"
PHYSS_RT_DMMA=1 $B --steps 2 --warmup 1 > gpurun_out/r2_d32_dmma1.json 2> gpurun_out/r2_d32.err
PHYSS_RT_DMMA=0 $B --steps 2 --warmup 1 > gpurun_out/r2_d32_dmma0.json 2>> gpurun_out/r2_d32.err
for v in 1 0; do
  PHYSS_RT_DMMA=$v ncu --set full --clock-control none --import-source on -k regex:rt_smooth -c 1 -f \
    -o gpurun_out/r2_ncu_d32_dmma$v $B --T 300 --steps 1 --warmup 1 > gpurun_out/r2_ncu_d32_dmma$v.log 2>&1
done
tail -c 300 gpurun_out/r2_d32.err
