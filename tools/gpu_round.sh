set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python bench.py --kernel-breakdown > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r01b.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --profile-step > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm2p -c 3 -f -o gpurun_out/prof_gemm2p_r01b python bench.py --steps 2 --warmup 1 --no-cpu-baseline --profile-step > gpurun_out/ncu_g2.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attention -c 3 -f -o gpurun_out/prof_attn_r01b python bench.py --steps 2 --warmup 1 --no-cpu-baseline --profile-step > gpurun_out/ncu_at.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_a.json; cat gpurun_out/smoke.log | tail -2
