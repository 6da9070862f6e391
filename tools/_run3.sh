for v in "4 4 2" "4 4 4" "3 8 4" "3 8 8" "2 8 8" "5 0 2"; do set -- $v
  export CM_NVCC_EXTRA="-DCM_BWDSP_MINB=$1 -DCM_BWDSP_HREG=$2 -DCM_BWDSP_UNROLL=$3"
  python mamba_asr_b200/build.py >/dev/null 2>&1 || echo BUILD FAIL
  echo "== MINB=$1 HREG=$2 UNR=$3"
  timeout 300 python tools/prof_kernels.py --cfg 2,3 --only scan_bwd 2>&1 | grep scan_bwd | cut -c1-120
done 2>&1 | tee gpurun_out/prof_spb4.log
