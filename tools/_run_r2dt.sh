# round 2, call DT: GELU + dropout forward - standalone timing and ncu --set full inside the step
set -x
mkdir -p gpurun_out
timeout 300 python tools/prof_elementwise.py 2>&1 | grep -i "gelu\|glu\|layernorm\|ln" > gpurun_out/r2dt_prof.log; cat gpurun_out/r2dt_prof.log
timeout 400 ncu --set full --clock-control none --import-source on --graph-profiling node -k regex:gelu_dropout_fwd_kernel --launch-skip 40 -c 2 -f -o gpurun_out/r2dt_gelu_fwd python tools/step_profile.py --graphed --top 1 > gpurun_out/r2dt_ncu.log 2>&1
ls -la gpurun_out/r2dt_gelu_fwd.ncu-rep
