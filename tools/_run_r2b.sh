# round 2, call B: first run of the lane-per-channel TMA forward kernel
set -x
timeout 900 python -m pytest tests/test_gpu_scan.py -m gpu -x -q 2>&1 | tail -15
timeout 300 python tools/prof_kernels.py --cfg 2,3,4 --only scan_fwd,scan_fwd_infer > gpurun_out/r2b_prof_lc.log 2>&1
CM_SCAN_NO_POLY=1 timeout 300 python tools/prof_kernels.py --cfg 2,3 --only scan_fwd > gpurun_out/r2b_prof_lc_nopoly.log 2>&1
CM_SCAN_NO_LC=1 timeout 300 python tools/prof_kernels.py --cfg 2,3 --only scan_fwd > gpurun_out/r2b_prof_sp.log 2>&1
cut -c1-170 gpurun_out/r2b_prof_lc.log gpurun_out/r2b_prof_lc_nopoly.log gpurun_out/r2b_prof_sp.log
