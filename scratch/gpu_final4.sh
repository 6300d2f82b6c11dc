timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
C="--steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
python bench.py --state-dim 8 --series 14208 --sub-batch 7104 $C > gpurun_out/f_c5_d8.json 2>/dev/null
python bench.py --state-dim 16 --series 5328 --sub-batch 2664 $C > gpurun_out/f_c5_d16.json 2>/dev/null
python bench.py --state-dim 32 --series 1480 --sub-batch 740 $C > gpurun_out/f_c5_d32.json 2>/dev/null
python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/f_c3.json 2>/dev/null
for f in f_c5_d8 f_c5_d16 f_c5_d32 f_c3; do python - <<PY
import json
j=json.load(open('gpurun_out/$f.json')); r=j.get('roofline') or {}
print('$f', '%.4g'%j['value'], 'ms/step %.2f'%j['ms_per_step'], 'whole', r.get('whole_step_frac'), 'fp64', (r.get('fp64') or {}).get('frac'))
PY
done
