# round 2, call L: warpgroup backward kernel iterations - quick parity subset + timing
set -x
timeout 900 python -m pytest tests/test_gpu_scan.py -m gpu -x -q -k "backward or bwd or grad" 2>&1 | tail -4
timeout 300 python tools/prof_kernels.py --cfg 2,3,4 --only scan_bwd 2>&1 | cut -c1-130
