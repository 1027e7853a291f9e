set -x
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r02a.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --profile-step > gpurun_out/ncu_l.log 2>&1
tail -2 gpurun_out/ncu_l.log
python bench.py --no-cpu-baseline > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; cut -c1-400 gpurun_out/bench_a.json
