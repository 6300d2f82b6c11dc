set -x
python bench.py > gpurun_out/f_c5.json 2> gpurun_out/f_c5.err; tail -c 300 gpurun_out/f_c5.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/f_ref.json 2> gpurun_out/f_ref.err
python bench.py --workload c3 --steps 3 --warmup 3 > gpurun_out/f_c3.json 2> gpurun_out/f_c3.err; tail -c 300 gpurun_out/f_c3.err
python bench.py --workload c3 --state-dim 4 --obs-dim 1 --chunk-len 64 --steps 3 --warmup 3 > gpurun_out/f_c3_d4.json 2> gpurun_out/f_c3d4.err; tail -c 300 gpurun_out/f_c3d4.err
python bench.py --workload cvi --steps 3 --warmup 3 > gpurun_out/f_cvi.json 2> gpurun_out/f_cvi.err; tail -c 300 gpurun_out/f_cvi.err
python bench.py --workload c2 --steps 1 --warmup 1 > gpurun_out/f_c2.json 2> gpurun_out/f_c2.err; tail -c 300 gpurun_out/f_c2.err
python bench.py --state-dim 8 --series 16384 --sub-batch 8192 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/f_c5_d8.json 2>/dev/null
python bench.py --state-dim 16 --series 4096 --sub-batch 2048 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/f_c5_d16.json 2>/dev/null
python bench.py --state-dim 32 --series 1024 --sub-batch 512 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/f_c5_d32.json 2>/dev/null
