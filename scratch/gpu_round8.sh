set -x
timeout 900 python -m pytest tests/test_gpu_big.py tests/test_gpu_cvi.py -x -q 2>&1 | tail -12
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --workload c2 --steps 1 --warmup 1 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 600 gpurun_out/bench_c2.err; cut -c1-900 gpurun_out/bench_c2.json
