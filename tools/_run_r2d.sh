# round 2, call D: lc forward v2 (plain smem loads, pipelined scalars): TMA-path tests, NPOLY variants
set -x
timeout 900 python -m pytest tests/test_gpu_scan.py -m gpu -x -q -k "tma or benchmark_widths or bf16_backward" 2>&1 | tail -5
for v in "" _np0 _np1 _np3; do
  echo "== variant '$v'"
  CM_LIB_PATH=$PWD/mamba_asr_b200/lib/libconmamba_b200$v.so timeout 300 python tools/prof_kernels.py --cfg 2,3 --only scan_fwd 2>&1 | cut -c1-150
done
CM_LIB_PATH=$PWD/mamba_asr_b200/lib/libconmamba_b200.so timeout 300 python tools/prof_kernels.py --cfg 4,5 --only scan_fwd,scan_fwd_infer 2>&1 | cut -c1-150
