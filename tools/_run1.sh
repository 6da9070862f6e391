timeout 600 python -m pytest tests/test_gpu_scan.py -q -m gpu -x 2>&1 | tail -15 > gpurun_out/pytest_sp.log
cat gpurun_out/pytest_sp.log
timeout 300 python tools/prof_kernels.py --cfg 2,3,5_4k --only scan_fwd,scan_bwd 2>&1 | tee gpurun_out/prof_sp2.log
