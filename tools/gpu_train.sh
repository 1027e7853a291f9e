set -x
timeout 600 python -m pytest tests/test_train_gpu.py -q -x > gpurun_out/train_gpu.log 2>&1
tail -5 gpurun_out/train_gpu.log
