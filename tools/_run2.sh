ncu --set full --import-source on --clock-control none -k regex:scan_fwd_sp --launch-skip 2 --launch-count 1 -o gpurun_out/scanfwd_sp_a -f python tools/prof_kernels.py --cfg 3 --only scan_fwd --iters 2 > gpurun_out/ncu_sp.log 2>&1
tail -3 gpurun_out/ncu_sp.log
