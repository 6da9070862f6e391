for v in "16 4" "8 8" "8 4" "4 16" "4 8"; do set -- $v
  export CM_NVCC_EXTRA="-DCM_CONV_TL=$1 -DCM_CONV_TY=$2"
  python mamba_asr_b200/build.py >/dev/null 2>&1 || echo BUILD FAIL
  echo "== CONV TL=$1 TY=$2"; cuobjdump -res-usage mamba_asr_b200/build/conv.o | grep -A1 "conv_bwd_cl_kernelI13__nv_bfloat16Li2\|conv_fwd_cl_kernelI13__nv_bfloat16Li2" | grep REG | cut -c1-40
  timeout 300 python tools/prof_kernels.py --cfg 2,3,4 --only conv_fwd,conv_bwd 2>&1 | grep "conv_" | cut -c1-140
done
unset CM_NVCC_EXTRA
python mamba_asr_b200/build.py >/dev/null 2>&1
