#!/bin/bash
# what the driver runs at round end, plus the CPU arms: GPU parity suite, smoke(), default bench, reference arms
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/r2_final2_tests.log 2>&1; tail -4 gpurun_out/r2_final2_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final2_smoke.log 2>&1; tail -2 gpurun_out/r2_final2_smoke.log
(time python bench.py --gpus 1 --steps 20 --warmup 5) > gpurun_out/r2_final2_bench.json 2> gpurun_out/r2_final2_bench.err; tail -c 300 gpurun_out/r2_final2_bench.err
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_final2_ref_c5.json 2>/dev/null
python bench.py --impl reference --workload cvi --steps 3 --warmup 1 > gpurun_out/r2_final2_ref_cvi.json 2>/dev/null
python bench.py --impl reference --workload c3 --steps 3 --warmup 1 > gpurun_out/r2_final2_ref_c3.json 2>/dev/null
