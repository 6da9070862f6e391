# round 2, call M: warpgroup backward kernel variants (IO warps, register history depth, register split)
for v in "" _ia4 _hr3 _hr5 _hr6 _hr8 _r128 _r128hr6; do
  echo "== variant '$v'"
  CM_LIB_PATH=$PWD/mamba_asr_b200/lib/libconmamba_b200$v.so timeout 300 python tools/prof_kernels.py --cfg 3 --only scan_bwd 2>&1 | cut -c1-110
done
