# round 2, call T: ncu --set full of the shipped conv kernels at the ConMamba-large shape
set -x
timeout 300 python tools/prof_kernels.py --cfg 3 --only conv_fwd,conv_bwd --iters 3 > gpurun_out/r2t_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_.*_cl_kernel -s 2 -c 2 -o gpurun_out/r2t_conv_cfg3 python tools/prof_kernels.py --cfg 3 --only conv_fwd,conv_bwd --iters 3 > gpurun_out/r2t_ncu.log 2>&1
tail -2 gpurun_out/r2t_ncu.log; cat gpurun_out/r2t_plain.log | cut -c1-150
