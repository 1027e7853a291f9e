set -x
timeout 900 python -m pytest tests/test_train_ops_gpu.py -q -k "artificial or generated" > gpurun_out/art.log 2>&1
tail -25 gpurun_out/art.log
