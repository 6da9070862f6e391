# round 2, call K: ncu --set full of the warpgroup backward kernel at the ConMamba-large shape
set -x
timeout 300 python tools/prof_kernels.py --cfg 3 --only scan_bwd --iters 3 > gpurun_out/r2k_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_bwd_wg -s 2 -c 1 -o gpurun_out/r2k_bwd_wg_cfg3 python tools/prof_kernels.py --cfg 3 --only scan_bwd --iters 3 > gpurun_out/r2k_ncu.log 2>&1
tail -2 gpurun_out/r2k_ncu.log
