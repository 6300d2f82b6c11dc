#!/bin/bash
# second multi-GPU pass: time-sharded config 3 at T = 8M on 1 / 2 / 4 / 8 GPUs, and the strong-scaling point at 8
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python bench.py --workload c3 --T 8000000 --steps 3 --warmup 2 > gpurun_out/r2_mg_c3_n1.json 2> gpurun_out/r2_mg2.err
for n in 2 4 8; do
  $TR --nproc-per-node $n --master-port $((29820+n)) bench.py --gpus $n --workload c3 --T 8000000 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2_mg_c3_n$n.json 2>> gpurun_out/r2_mg2.err
done
$TR --nproc-per-node 8 --master-port 29840 bench.py --gpus 8 --scaling strong --steps 5 --warmup 3 --no-sweep --no-cpu-baseline > gpurun_out/r2_mg_c5strong_n8_bound.json 2>> gpurun_out/r2_mg2.err
$TR --nproc-per-node 8 --master-port 29841 bench.py --gpus 8 --steps 5 --warmup 3 --no-sweep --no-cpu-baseline > gpurun_out/r2_mg_c5weak_n8_bound.json 2>> gpurun_out/r2_mg2.err
grep -i -E "error|Traceback" gpurun_out/r2_mg2.err | head -5
