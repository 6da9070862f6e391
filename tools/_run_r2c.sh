# round 2, call C: lc forward through the model-level tests + ncu --set full of the kernel at cfg3
set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 300 python tools/prof_kernels.py --cfg 3 --only scan_fwd --iters 3 > gpurun_out/r2c_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_fwd_lc -s 2 -c 1 -o gpurun_out/r2c_fwd_lc_cfg3 python tools/prof_kernels.py --cfg 3 --only scan_fwd --iters 3 > gpurun_out/r2c_ncu.log 2>&1
tail -3 gpurun_out/r2c_ncu.log
ls -la gpurun_out/*.ncu-rep | tail -3
