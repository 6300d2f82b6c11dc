#!/bin/bash
run() { name=$1; shift; "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err || tail -c 300 gpurun_out/$name.err
python - <<PY
import json
try:
    j=json.load(open('gpurun_out/$name.json'))
    r=j.get('roofline',{})
    print('$name', '%.4g'%j['value'], 'ms/step %.2f'%j['ms_per_step'], {k:round(v['avg_ms'],2) for k,v in r.get('kernels',{}).items()}, 'whole_step_frac', r.get('whole_step_frac'), 'fp64', r.get('fp64',{}).get('frac'))
except Exception as e: print('$name failed', e)
PY
}
C="--steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
run g_c5_d32 python bench.py --state-dim 32 --series 1024 --sub-batch 512 $C
run g_c5_d16 python bench.py --state-dim 16 --series 4096 --sub-batch 2048 $C
run g_c5_d8 python bench.py --state-dim 8 --series 16384 --sub-batch 8192 $C
PHYSS_NO_SEQ8=1 run g_c5_d8_rt python bench.py --state-dim 8 --series 16384 --sub-batch 8192 $C
run g_c3 python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline
