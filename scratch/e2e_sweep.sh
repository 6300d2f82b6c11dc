for sb in 8192 16384; do
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-sub-batch $sb > gpurun_out/e2e_$sb.json 2> gpurun_out/e2e_$sb.err || tail -c 400 gpurun_out/e2e_$sb.err
python -c "
import json; j=json.load(open('gpurun_out/e2e_$sb.json')); print($sb, 'value', j['value'], 'e2e', j['e2e']['value'], 'loss_only', j['e2e']['loss_only']['value'])"
done
