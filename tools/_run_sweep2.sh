for v in "5 4 2" "5 2 2" "5 3 2" "4 5 2" "4 6 2"; do set -- $v
  export CM_NVCC_EXTRA="-DCM_BWDSP_MINB=$1 -DCM_BWDSP_HREG=$2 -DCM_BWDSP_UNROLL=$3"
  python mamba_asr_b200/build.py >/dev/null 2>&1 || echo BUILD FAIL
  echo "== MINB=$1 HREG=$2 UNR=$3"
  cuobjdump -res-usage mamba_asr_b200/build/scan_bwd_sp.o | grep -A1 "bfloat" | grep REG
  timeout 300 python tools/prof_kernels.py --cfg 2,3 --only scan_bwd 2>&1 | grep scan_bwd | cut -c1-120
done
unset CM_NVCC_EXTRA
python mamba_asr_b200/build.py >/dev/null 2>&1
