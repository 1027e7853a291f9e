set -x
timeout 900 python -m pytest tests/test_model_gpu.py -q -x -k "general_patch" > gpurun_out/model_gpu.log 2>&1
tail -30 gpurun_out/model_gpu.log
