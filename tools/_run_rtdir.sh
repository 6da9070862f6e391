for v in "1 1" "0 0" "1 0" "0 1"; do set -- $v
  export CM_NVCC_EXTRA="-DCM_BWDSP_RTDIR=$1 -DCM_FWDSP_RTDIR=$2"
  python mamba_asr_b200/build.py >/dev/null 2>&1 || echo BUILD FAIL
  echo "== BWD_RTDIR=$1 FWD_RTDIR=$2"
  timeout 300 python tools/prof_kernels.py --cfg 2,3 --only scan_fwd,scan_bwd 2>&1 | grep "scan_" | cut -c1-130
done
unset CM_NVCC_EXTRA
python mamba_asr_b200/build.py >/dev/null 2>&1
timeout 600 python -m pytest tests/test_gpu_scan.py -q -m gpu -x 2>&1 | tail -2
