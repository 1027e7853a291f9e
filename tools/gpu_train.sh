set -x
timeout 900 python -m pytest tests/test_train_gpu.py -q -x -s > gpurun_out/train_gpu.log 2>&1
tail -60 gpurun_out/train_gpu.log
