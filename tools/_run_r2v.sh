# round 2, call V: ncu --set full of the shipped backward kernels at the ConMamba-large shape (scan_bwd_wg final, conv_bwd_cl)
set -x
timeout 300 python tools/prof_kernels.py --cfg 3 --only scan_bwd,conv_bwd --iters 3 > gpurun_out/r2v_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_bwd_wg -s 2 -c 1 -o gpurun_out/r2v_bwd_wg_cfg3 python tools/prof_kernels.py --cfg 3 --only scan_bwd --iters 3 > gpurun_out/r2v_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_bwd_cl -s 2 -c 1 -o gpurun_out/r2v_conv_bwd_cfg3 python tools/prof_kernels.py --cfg 3 --only conv_bwd --iters 3 > gpurun_out/r2v_ncu2.log 2>&1
tail -1 gpurun_out/r2v_ncu1.log gpurun_out/r2v_ncu2.log; cut -c1-140 gpurun_out/r2v_plain.log
