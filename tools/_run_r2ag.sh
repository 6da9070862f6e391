set -x
timeout 600 python -m pytest tests/test_gpu_scan.py -m gpu -x -q -k "warpgroup_backward" 2>&1 | tail -3
for v in "" _p0 _p1f0 _p1hr3; do
  echo "== variant '$v'"
  CM_LIB_PATH=$PWD/mamba_asr_b200/lib/libconmamba_b200$v.so timeout 300 python tools/prof_kernels.py --cfg 3,4 --only scan_bwd 2>&1 | cut -c1-110
done
