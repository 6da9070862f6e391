# round 2, call N: timing ablations of the warpgroup backward kernel (results are wrong by construction)
for v in "" _noex2 _nomma _nopb _nofwd _nofwdpbmma; do
  echo "== variant '$v'"
  CM_LIB_PATH=$PWD/mamba_asr_b200/lib/libconmamba_b200$v.so timeout 300 python tools/prof_kernels.py --cfg 3 --only scan_bwd 2>&1 | cut -c1-110
done
