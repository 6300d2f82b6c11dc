#!/bin/bash
# parity of the general-d kernels + short timing runs at d = 8 (rt forced), 16, 32
timeout 900 python -m pytest tests/test_gpu_grp.py tests/test_gpu_pscan.py tests/test_gpu_golden.py tests/test_gpu_cvi.py -x -q 2>&1 | tail -5
for d in 32 16 8; do
  B=$((16384/d))
  PHYSS_NO_SEQ8=1 python bench.py --workload c5 --state-dim $d --series $B --sub-batch $B --T 300 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/prt2_d$d.json 2> gpurun_out/prt2_d$d.err
  python - <<PY
import json
j=json.load(open('gpurun_out/prt2_d$d.json'))
print($d, j['value'], {k:round(v['avg_ms'],3) for k,v in j['roofline']['kernels'].items()})
PY
done
