set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --workload cvi --steps 3 --warmup 2 > gpurun_out/bench_cvi_auto.json 2> gpurun_out/bench_cvi_auto.err; tail -c 400 gpurun_out/bench_cvi_auto.err
python bench.py --workload c3 --steps 2 --warmup 1 --state-dim 4 --obs-dim 1 --chunk-len 64 > gpurun_out/bench_c3_d4.json 2> gpurun_out/bench_c3_d4.err; tail -c 400 gpurun_out/bench_c3_d4.err
