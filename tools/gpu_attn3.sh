mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py tests/test_train_ops_gpu.py -q -k "attention" 2>&1 | tail -5
BENCH_SHAPES=5 timeout 100 python tools/bench_attn_fwd.py 2>&1 | tee gpurun_out/bench_attn_fwd.log | cut -c1-150
