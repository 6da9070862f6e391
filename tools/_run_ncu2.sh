timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_bwd_sp --launch-skip 1 --launch-count 1 -o gpurun_out/s3b_scanbwd_cfg3 -f python tools/prof_kernels.py --cfg 3 --only scan_bwd --iters 2 > gpurun_out/s3b_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_fwd_sp --launch-skip 2 --launch-count 1 -o gpurun_out/s3b_scanfwd_cfg3 -f python tools/prof_kernels.py --cfg 3 --only scan_fwd --iters 2 > gpurun_out/s3b_ncu2.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:scan_bwd_sp --launch-skip 1 --launch-count 1 -o gpurun_out/s3b_scanbwd_cfg2 -f python tools/prof_kernels.py --cfg 2 --only scan_bwd --iters 2 > gpurun_out/s3b_ncu3.log 2>&1
ls -la gpurun_out/s3b_*.ncu-rep
