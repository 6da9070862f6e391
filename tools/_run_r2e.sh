set -x
./tools/ubench/smsp_map
CM_LIB_PATH=$PWD/mamba_asr_b200/lib/libconmamba_b200_np0.so timeout 300 python tools/prof_kernels.py --cfg 3 --only scan_fwd --iters 3 > gpurun_out/r2e_plain.log 2>&1 && \
CM_LIB_PATH=$PWD/mamba_asr_b200/lib/libconmamba_b200_np0.so timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_fwd_lc -s 2 -c 1 -o gpurun_out/r2e_fwd_lc_np0_cfg3 python tools/prof_kernels.py --cfg 3 --only scan_fwd --iters 3 > gpurun_out/r2e_ncu.log 2>&1
tail -2 gpurun_out/r2e_ncu.log
