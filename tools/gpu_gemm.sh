timeout 900 python -m pytest tests/test_gemm_ts_gpu.py -q -x -k "conv" > gpurun_out/conv_tests.log 2>&1; echo "rc=$?" >> gpurun_out/conv_tests.log
tail -30 gpurun_out/conv_tests.log
