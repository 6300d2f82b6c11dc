#!/bin/bash
# what the driver runs at round end, plus the CPU arms: GPU parity suite, smoke(), default bench, reference arms,
# and the ncu launch list of the (short) default bench command
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/r2_final3_tests.log 2>&1; tail -4 gpurun_out/r2_final3_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final3_smoke.log 2>&1; tail -2 gpurun_out/r2_final3_smoke.log
(time python bench.py --gpus 1 --steps 20 --warmup 5) > gpurun_out/r2_final3_bench.json 2> gpurun_out/r2_final3_bench.err; tail -c 300 gpurun_out/r2_final3_bench.err
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_final3_ref_c5.json 2>/dev/null
B="python bench.py --no-sweep --no-e2e --no-cpu-baseline --steps 2 --warmup 3"
$B > gpurun_out/r2_final3_short.json 2>/dev/null && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_final3_launches.csv \
  $B > gpurun_out/r2_final3_ncu.log 2>&1
tail -2 gpurun_out/r2_final3_ncu.log
