# round 2, call Q: warpgroup forward kernel (scan_fwd_wg.cu) - parity + timing against the state-parallel kernel
set -x
timeout 900 python -m pytest tests/test_gpu_scan.py -m gpu -x -q 2>&1 | tail -8
timeout 300 python tools/prof_kernels.py --cfg 1,2,3,4 --only scan_fwd,scan_fwd_infer 2>&1 | cut -c1-160
CM_SCAN_NO_WG=1 timeout 300 python tools/prof_kernels.py --cfg 2,3 --only scan_fwd 2>&1 | cut -c1-160
