timeout 900 python -m pytest tests/test_gpu_scan.py -q -m gpu -x 2>&1 | tail -6
for nw in 1 0; do if [ $nw = 1 ]; then export CM_SCAN_NO_WINDOWS=1; else unset CM_SCAN_NO_WINDOWS; fi; echo "NO_WINDOWS=$nw"; timeout 300 python tools/prof_kernels.py --cfg 5_1k,5_4k,5,5_30k --only scan_fwd_infer 2>&1 | grep scan_fwd; done
