B="--state-dim 8 --series 16384 --sub-batch 8192 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
PHYSS_NO_SEQ8=1 PHYSS_RT_WIDE=1 python bench.py $B > gpurun_out/b_d8_rtw.json 2>gpurun_out/e2.err; tail -c 300 gpurun_out/e2.err
B="--state-dim 16 --series 4096 --sub-batch 2048 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
PHYSS_RT_WIDE=1 python bench.py $B > gpurun_out/b_d16_rtw.json 2>gpurun_out/e3.err; tail -c 300 gpurun_out/e3.err
B="--state-dim 32 --series 1024 --sub-batch 512 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
PHYSS_RT_WIDE=1 python bench.py $B > gpurun_out/b_d32_rtw.json 2>gpurun_out/e5.err; tail -c 300 gpurun_out/e5.err
PHYSS_RT_WIDE=1 python bench.py --workload c3 --steps 2 --warmup 1 --no-e2e > gpurun_out/b_c3_rtw.json 2>gpurun_out/e7.err; tail -c 300 gpurun_out/e7.err
PHYSS_RT_WIDE=1 python -m pytest tests/test_gpu_grp.py tests/test_gpu_pscan.py -x -q 2>&1 | tail -3
