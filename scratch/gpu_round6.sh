set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -12
B="--state-dim 8 --series 16384 --sub-batch 8192 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
python bench.py $B > gpurun_out/b_d8_seq.json 2>gpurun_out/e1.err; tail -c 300 gpurun_out/e1.err
PHYSS_NO_SEQ8=1 python bench.py $B > gpurun_out/b_d8_rt.json 2>gpurun_out/e2.err; tail -c 300 gpurun_out/e2.err
B="--state-dim 16 --series 4096 --sub-batch 2048 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
python bench.py $B > gpurun_out/b_d16_rt.json 2>gpurun_out/e3.err; tail -c 300 gpurun_out/e3.err
PHYSS_FORCE_GRP=1 python bench.py $B > gpurun_out/b_d16_grp.json 2>gpurun_out/e4.err; tail -c 300 gpurun_out/e4.err
B="--state-dim 32 --series 1024 --sub-batch 512 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
python bench.py $B > gpurun_out/b_d32_rt.json 2>gpurun_out/e5.err; tail -c 300 gpurun_out/e5.err
PHYSS_FORCE_GRP=1 python bench.py $B > gpurun_out/b_d32_grp.json 2>gpurun_out/e6.err; tail -c 300 gpurun_out/e6.err
python bench.py --workload c3 --steps 2 --warmup 1 --no-e2e > gpurun_out/b_c3_rt.json 2>gpurun_out/e7.err; tail -c 300 gpurun_out/e7.err
