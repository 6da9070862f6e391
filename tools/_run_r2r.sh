# round 2, call R: ncu --set full of the warpgroup forward kernel at the ConMamba-large shape
set -x
timeout 300 python tools/prof_kernels.py --cfg 3 --only scan_fwd --iters 3 > gpurun_out/r2r_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_fwd_wg -s 2 -c 1 -o gpurun_out/r2r_fwd_wg_cfg3 python tools/prof_kernels.py --cfg 3 --only scan_fwd --iters 3 > gpurun_out/r2r_ncu.log 2>&1
tail -2 gpurun_out/r2r_ncu.log
