for v in "" _mb4 _mb5 _mb6; do
  echo "== variant '$v'"
  CM_LIB_PATH=$PWD/mamba_asr_b200/lib/libconmamba_b200$v.so timeout 300 python tools/prof_kernels.py --cfg 2,3 --only conv_bwd 2>&1 | cut -c1-110
done
