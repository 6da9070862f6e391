timeout 900 python -m pytest tests/test_gpu_fused_ln.py tests/test_gpu_models.py -q -m gpu -x 2>&1 | tail -25
