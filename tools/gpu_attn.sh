timeout 900 python -m pytest tests/test_ops_gpu.py -q -x -k "attention" > gpurun_out/attn_tests.log 2>&1; echo "rc=$?" >> gpurun_out/attn_tests.log
tail -3 gpurun_out/attn_tests.log
timeout 600 python tools/bench_attn.py > gpurun_out/bench_attn.log 2>&1
tail -20 gpurun_out/bench_attn.log | cut -c1-110
