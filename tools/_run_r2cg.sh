# round 2, call CG: ncu --set full of the kernels added in this session, inside the graphed step
set -x
mkdir -p gpurun_out
for k in reduce_batch_kernel dwconv_fwd_tile_kernel glu_bwd_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on --graph-profiling node -k regex:$k --launch-skip 30 -c 8 -f -o gpurun_out/r2cg_$k python tools/step_profile.py --graphed --top 1 > gpurun_out/r2cg_ncu_$k.log 2>&1
done
ls -la gpurun_out/r2cg_*.ncu-rep
