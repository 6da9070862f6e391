for mb in 3 2 4; do
  export CM_NVCC_EXTRA="-DCM_FL_BWD_MINB=$mb"
  python mamba_asr_b200/build.py >/dev/null 2>&1 || echo BUILD FAIL
  echo "== MINB=$mb"
  timeout 300 python tools/prof_kernels.py --cfg 2,3 --only aux 2>&1 | grep "add_ln\|ln_bwd" | cut -c1-140
done
unset CM_NVCC_EXTRA
python mamba_asr_b200/build.py >/dev/null 2>&1
