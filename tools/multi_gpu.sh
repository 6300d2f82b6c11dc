#!/bin/bash
# Multi-GPU measurements on ONE box with N GPUs (gpurun --gpus 8): strong scaling of BASELINE config 5 (65,536 series
# in total split over the ranks), and the time-sharded scan of config 3 at T = 8M (every GPU holds >= 2 waves of
# chunks) at 1 / 2 / 4 / 8 GPUs; plus the NCCL parity check of the time-sharded path.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
NG=${1:-8}
$TR --nproc-per-node 2 --master-port 29711 tests/dist/run_timeshard_nccl.py 400000 8 > gpurun_out/r2_mg_timeshard_check.json 2> gpurun_out/r2_mg.err
tail -c 600 gpurun_out/r2_mg_timeshard_check.json
python bench.py --workload c3 --T 8000000 --steps 3 --warmup 2 --no-e2e > gpurun_out/r2_mg_c3_n1.json 2>> gpurun_out/r2_mg.err
python bench.py --scaling strong --steps 5 --warmup 3 --no-sweep --no-cpu-baseline > gpurun_out/r2_mg_c5strong_n1.json 2>> gpurun_out/r2_mg.err
for n in 2 4 8; do
  if [ $n -le $NG ]; then
    $TR --nproc-per-node $n --master-port $((29720+n)) bench.py --gpus $n --workload c3 --T 8000000 --steps 3 --warmup 2 --no-e2e > gpurun_out/r2_mg_c3_n$n.json 2>> gpurun_out/r2_mg.err
    $TR --nproc-per-node $n --master-port $((29730+n)) bench.py --gpus $n --scaling strong --steps 5 --warmup 3 --no-sweep --no-cpu-baseline > gpurun_out/r2_mg_c5strong_n$n.json 2>> gpurun_out/r2_mg.err
  fi
done
tail -c 800 gpurun_out/r2_mg.err
