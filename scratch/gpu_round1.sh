set -x
python bench.py > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; tail -c 600 gpurun_out/bench_c5.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py --workload c3 --steps 2 --warmup 1 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -c 600 gpurun_out/bench_c3.err
python bench.py --workload cvi --steps 3 --warmup 2 > gpurun_out/bench_cvi.json 2> gpurun_out/bench_cvi.err; tail -c 600 gpurun_out/bench_cvi.err
python bench.py --steps 1 --warmup 1 --series 32768 --no-e2e --no-cpu-baseline > gpurun_out/plain_small.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01b.csv python bench.py --steps 1 --warmup 1 --series 32768 --no-e2e --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:seq_ -c 2 -o gpurun_out/prof_seq_r01b -f python bench.py --steps 1 --warmup 1 --series 32768 --no-e2e --no-cpu-baseline > gpurun_out/ncu_f.log 2>&1
ls -la gpurun_out
