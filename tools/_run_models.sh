timeout 900 python -m pytest tests/test_gpu_models.py -q -m gpu 2>&1 | tail -25
