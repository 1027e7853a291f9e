set -x
timeout 600 python -m pytest tests/test_train_ops_gpu.py -q -k "not attention" > gpurun_out/train_ops_a.log 2>&1
timeout 600 python -m pytest tests/test_train_ops_gpu.py -q -k "attention" > gpurun_out/train_ops_b.log 2>&1
tail -40 gpurun_out/train_ops_a.log; tail -40 gpurun_out/train_ops_b.log
