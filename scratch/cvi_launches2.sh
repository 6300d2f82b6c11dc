#!/bin/bash
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_cvi.csv python bench.py --workload cvi --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_cvi.log 2>&1
