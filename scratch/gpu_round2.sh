set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --workload cvi --steps 3 --warmup 2 --filter-type b200 > gpurun_out/bench_cvi_seq.json 2> gpurun_out/bench_cvi_seq.err; tail -c 400 gpurun_out/bench_cvi_seq.err
python bench.py --workload cvi --steps 3 --warmup 2 > gpurun_out/bench_cvi_auto.json 2> gpurun_out/bench_cvi_auto.err; tail -c 400 gpurun_out/bench_cvi_auto.err
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5_e2e.json 2> gpurun_out/bench_c5_e2e.err; tail -c 600 gpurun_out/bench_c5_e2e.err
