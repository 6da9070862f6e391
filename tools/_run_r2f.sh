# round 2, call F: first run of the lane-per-channel TMA backward kernel
set -x
timeout 900 python -m pytest tests/test_gpu_scan.py -m gpu -x -q -k "tma_backward or benchmark_widths or bf16_backward" 2>&1 | tail -12
timeout 300 python tools/prof_kernels.py --cfg 3,4 --only scan_bwd 2>&1 | cut -c1-200
CM_SCAN_NO_LC=1 timeout 300 python tools/prof_kernels.py --cfg 3 --only scan_bwd 2>&1 | cut -c1-200
