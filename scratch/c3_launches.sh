#!/bin/bash
python bench.py --state-dim 8 --series 16384 --sub-batch 8192 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/g_c5_d8.json 2>gpurun_out/g_c5_d8.err
python -c "
import json; j=json.load(open('gpurun_out/g_c5_d8.json')); print('d8', j['value'], {k:round(v['avg_ms'],2) for k,v in j['roofline']['kernels'].items()})"
python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c3_d8_v2.csv python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_c3.log 2>&1
tail -2 gpurun_out/ncu_c3.log
