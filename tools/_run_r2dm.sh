set -x
mkdir -p gpurun_out
timeout 600 python tools/step_profile.py --graphed --top 30 > gpurun_out/r2dm_step_large.txt 2>&1; grep -i "dwconv\|total CUDA" gpurun_out/r2dm_step_large.txt | cut -c1-150
timeout 600 ncu --set full --clock-control none --import-source on --graph-profiling node -k regex:dwconv_fwd_mma_kernel --launch-skip 10 -c 2 -f -o gpurun_out/r2dm_dwconv_mma python tools/step_profile.py --graphed --top 1 > gpurun_out/r2dm_ncu.log 2>&1
ls -la gpurun_out/r2dm_dwconv_mma.ncu-rep
