# round 2, call O: packed FP32x2 against scalar FP32 in the scan kernels
for v in "" _scalar; do
  echo "== variant '$v'"
  CM_LIB_PATH=$PWD/mamba_asr_b200/lib/libconmamba_b200$v.so timeout 300 python tools/prof_kernels.py --cfg 3 --only scan_bwd,scan_fwd 2>&1 | cut -c1-110
done
