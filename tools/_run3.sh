for v in "2 2 64" "3 2 64" "3 1 64" "4 1 64" "2 1 64"; do set -- $v
  export CM_NVCC_EXTRA="-DCM_FWDSP_MINB=$1 -DCM_FWDSP_SUB=$2 -DCM_FWDSP_IO=$3"
  python mamba_asr_b200/build.py >/dev/null 2>&1 || echo BUILD FAIL
  echo "== MINB=$1 SUB=$2 IO=$3"
  if [ "$v" = "2 2 64" ]; then timeout 300 python -m pytest tests/test_gpu_scan.py -q -m gpu -x 2>&1 | tail -3; fi
  timeout 300 python tools/prof_kernels.py --cfg 2,3,5_4k --only scan_fwd 2>&1 | grep scan_fwd
done 2>&1 | tee gpurun_out/prof_sp3.log
