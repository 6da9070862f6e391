timeout 900 python -m pytest tests/test_gpu_scan.py -q -m gpu -x 2>&1 | tail -4
timeout 300 python tools/prof_kernels.py --cfg 2,3,4,5_4k --only scan_bwd 2>&1 | grep scan_bwd | cut -c1-125
