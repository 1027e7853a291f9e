set -x
timeout 600 python -m pytest tests/test_ops_gpu.py -q -k "label_prop or topk" > gpurun_out/lp.log 2>&1
tail -15 gpurun_out/lp.log
timeout 600 python -m pytest tests/test_model_gpu.py -q -x > gpurun_out/model_gpu.log 2>&1
tail -8 gpurun_out/model_gpu.log
