for g in 1 2 4; do echo "groups $g"; SGF_STEM_GROUPS=$g python bench.py --no-cpu-baseline --no-train-record 2>/dev/null | cut -c100-160; done
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_parity_gpu.py -q -x 2>&1 | tail -3
