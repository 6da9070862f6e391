for v in "" _f0 _f1 _f4 _f0r128; do
  echo "== variant '$v'"
  CM_LIB_PATH=$PWD/mamba_asr_b200/lib/libconmamba_b200$v.so timeout 300 python tools/prof_kernels.py --cfg 3 --only scan_bwd 2>&1 | cut -c1-110
done
