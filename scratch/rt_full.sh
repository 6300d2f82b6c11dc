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
timeout 900 python -m pytest tests/test_gpu_grp.py tests/test_gpu_pscan.py tests/test_gpu_golden.py -x -q 2>&1 | tail -2
C="--steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
run g_c5_d32 python bench.py --state-dim 32 --series 1184 --sub-batch 592 $C
run g_c5_d16 python bench.py --state-dim 16 --series 4736 --sub-batch 2368 $C
run g_c5_d8 python bench.py --state-dim 8 --series 14208 --sub-batch 7104 $C
PHYSS_NO_SEQ8=1 run g_c5_d8_rt python bench.py --state-dim 8 --series 14208 --sub-batch 7104 $C
run g_c3 python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline
