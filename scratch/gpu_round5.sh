set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --state-dim 8 --series 16384 --sub-batch 8192 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_c5_d8_seq.json 2>gpurun_out/e1.err; tail -c 300 gpurun_out/e1.err
PHYSS_FORCE_GRP=1 python bench.py --state-dim 8 --series 16384 --sub-batch 8192 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_c5_d8_grp.json 2>gpurun_out/e2.err; tail -c 300 gpurun_out/e2.err
python bench.py --workload c3 --steps 2 --warmup 1 --no-e2e > gpurun_out/bench_c3_d8_seq.json 2>gpurun_out/e3.err; tail -c 300 gpurun_out/e3.err
python bench.py --workload c3 --steps 2 --warmup 1 --no-e2e --obs-dim 1 --chunk-len 64 > gpurun_out/bench_c3_d8m1_seq.json 2>gpurun_out/e4.err; tail -c 300 gpurun_out/e4.err
python bench.py --workload c3 --steps 2 --warmup 1 --no-e2e --state-dim 4 --obs-dim 1 --chunk-len 64 > gpurun_out/bench_c3_d4.json 2>gpurun_out/e5.err; tail -c 300 gpurun_out/e5.err
