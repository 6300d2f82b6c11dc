for L in 256 512 1024 2048; do python bench.py --workload c3cvi --steps 2 --warmup 1 --chunk-len $L 2>&1 | tail -2 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('L=$L', j['value'], 'ms')
    else: print('L=$L', l[:120].strip())
"; done
