timeout 600 python -m pytest tests/test_gpu_scan.py -q -m gpu 2>&1 | tail -5
ncu --set full --import-source on --clock-control none -k regex:scan_bwd_sp --launch-skip 1 --launch-count 1 -o gpurun_out/scanbwd_sp_a -f python tools/prof_kernels.py --cfg 3 --only scan_bwd --iters 2 > gpurun_out/ncu_sp.log 2>&1
tail -3 gpurun_out/ncu_sp.log
