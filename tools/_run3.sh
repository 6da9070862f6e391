for v in "3 8 2" "3 8 1" "4 4 1" "4 4 2" "3 4 2" "3 0 2" "4 8 2"; do set -- $v
  export CM_NVCC_EXTRA="-DCM_BWDSP_MINB=$1 -DCM_BWDSP_HREG=$2 -DCM_BWDSP_UNROLL=$3"
  python mamba_asr_b200/build.py >/dev/null 2>&1 || echo BUILD FAIL
  echo "== MINB=$1 HREG=$2 UNR=$3"
  if [ "$v" = "3 8 2" ]; then timeout 300 python -m pytest tests/test_gpu_scan.py -q -m gpu -x -k "backward or deterministic" 2>&1 | tail -3; fi
  timeout 300 python tools/prof_kernels.py --cfg 2,3,5_4k --only scan_bwd 2>&1 | grep scan_bwd | cut -c1-130
done 2>&1 | tee gpurun_out/prof_spb3.log
