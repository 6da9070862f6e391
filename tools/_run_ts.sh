timeout 900 python -m pytest tests/test_gpu_fused_ln.py -q -m gpu -x -k "tall_skinny or tsmm" 2>&1 | tail -3
timeout 300 python tools/prof_kernels.py --cfg 2,3,4 --only aux 2>&1 | grep "tsmm" | cut -c1-170
