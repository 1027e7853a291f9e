set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_stdout.txt 2> gpurun_out/bench_stderr.txt
echo "stdout lines: $(wc -l < gpurun_out/bench_stdout.txt)"; cut -c1-200 gpurun_out/bench_stdout.txt
