for v in "4 4 2" "3 8 4" "5 0 2" "4 4 4" "3 4 2" "6 0 2"; do set -- $v
  export CM_NVCC_EXTRA="-DCM_BWDSP_MINB=$1 -DCM_BWDSP_HREG=$2 -DCM_BWDSP_UNROLL=$3"
  python mamba_asr_b200/build.py >/dev/null 2>&1 || echo BUILD FAIL
  echo "== MINB=$1 HREG=$2 UNR=$3"
  timeout 300 python tools/prof_kernels.py --cfg 2,3,4 --only scan_bwd 2>&1 | grep scan_bwd | cut -c1-120
done
for v in 2 3 4; do
  export CM_NVCC_EXTRA="-DCM_FWDSP_MINB=$v"
  python mamba_asr_b200/build.py >/dev/null 2>&1 || echo BUILD FAIL
  echo "== FWD MINB=$v"
  timeout 300 python tools/prof_kernels.py --cfg 2,3,4 --only scan_fwd 2>&1 | grep scan_fwd | cut -c1-120
done
unset CM_NVCC_EXTRA
python mamba_asr_b200/build.py >/dev/null 2>&1
