set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --workload cvi --steps 3 --warmup 2 > gpurun_out/bench_cvi_auto.json 2> gpurun_out/bench_cvi_auto.err; tail -c 400 gpurun_out/bench_cvi_auto.err
python scratch/time_pscan_small.py
