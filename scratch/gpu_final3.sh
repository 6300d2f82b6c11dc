set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/f_c5.json 2> gpurun_out/f_c5.err; tail -c 300 gpurun_out/f_c5.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/f_ref.json 2> gpurun_out/f_ref.err
python bench.py --workload c3 --steps 3 --warmup 3 > gpurun_out/f_c3.json 2> gpurun_out/f_c3.err; tail -c 300 gpurun_out/f_c3.err
python bench.py --workload c3 --state-dim 4 --obs-dim 1 --chunk-len 64 --steps 3 --warmup 3 > gpurun_out/f_c3_d4.json 2> gpurun_out/f_c3d4.err
python bench.py --workload c3cvi --steps 3 --warmup 3 > gpurun_out/f_c3cvi.json 2> gpurun_out/f_c3cvi.err
python bench.py --workload cvi --steps 3 --warmup 3 > gpurun_out/f_cvi.json 2> gpurun_out/f_cvi.err
python bench.py --workload grad --steps 3 --warmup 3 > gpurun_out/f_grad.json 2> gpurun_out/f_grad.err
python bench.py --workload c2 --steps 1 --warmup 1 > gpurun_out/f_c2.json 2> gpurun_out/f_c2.err; tail -c 300 gpurun_out/f_c2.err
C="--steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
python bench.py --state-dim 8 --series 14208 --sub-batch 7104 $C > gpurun_out/f_c5_d8.json 2>gpurun_out/f_c5_d8.err
python bench.py --state-dim 16 --series 5328 --sub-batch 2664 $C > gpurun_out/f_c5_d16.json 2>gpurun_out/f_c5_d16.err
python bench.py --state-dim 32 --series 1480 --sub-batch 740 $C > gpurun_out/f_c5_d32.json 2>gpurun_out/f_c5_d32.err
for f in f_c5 f_ref f_c3 f_c3_d4 f_c3cvi f_cvi f_grad f_c2 f_c5_d8 f_c5_d16 f_c5_d32; do python - <<PY
import json
try:
    j=json.load(open('gpurun_out/$f.json')); r=j.get('roofline') or {}
    print('$f', '%.4g'%j['value'], j['unit'], 'ms/step %.2f'%j['ms_per_step'], 'frac', r.get('frac'), 'whole', r.get('whole_step_frac'), 'fp64', (r.get('fp64') or {}).get('frac'), 'e2e', (j.get('e2e') or {}).get('value'))
except Exception as e: print('$f failed', e)
PY
done
