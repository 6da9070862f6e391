# round 2, call G: lc forward v3 (2 lanes per channel, scalar tile in smem)
set -x
timeout 900 python -m pytest tests/test_gpu_scan.py -m gpu -x -q -k "tma_forward or benchmark_widths or bf16_backward" 2>&1 | tail -5
timeout 300 python tools/prof_kernels.py --cfg 2,3,4 --only scan_fwd,scan_fwd_infer 2>&1 | cut -c1-150
echo "== no poly"; CM_SCAN_NO_POLY=1 timeout 300 python tools/prof_kernels.py --cfg 2,3 --only scan_fwd 2>&1 | cut -c1-150
echo "== lanes 1"; CM_FWDLC_LANES=1 timeout 300 python tools/prof_kernels.py --cfg 2,3 --only scan_fwd 2>&1 | cut -c1-150
timeout 300 python tools/prof_kernels.py --cfg 3 --only scan_fwd --iters 3 > gpurun_out/r2g_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_fwd_lc -s 2 -c 1 -o gpurun_out/r2g_fwd_lc2_cfg3 python tools/prof_kernels.py --cfg 3 --only scan_fwd --iters 3 > gpurun_out/r2g_ncu.log 2>&1
tail -2 gpurun_out/r2g_ncu.log
