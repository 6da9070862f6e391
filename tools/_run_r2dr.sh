# round 2, call DR: evidence on the final tree - default bench line, ncu launch list of the eager bench command, `ncu --set
# full` of the two scan kernels as shipped
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2dr_large.log 2> gpurun_out/r2dr_large.err; tail -c 200 gpurun_out/r2dr_large.err
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/r2dr_nograph.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r2dr_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/r2dr_ncu_list.log 2>&1
for k in scan_fwd_sp_kernel scan_bwd_wg_kernel; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 3 -c 1 -f -o gpurun_out/r2dr_$k python tools/step_profile.py --top 1 > gpurun_out/r2dr_ncu_$k.log 2>&1
done
ls -la gpurun_out/r2dr_*.ncu-rep gpurun_out/r2dr_launches.csv
tail -n 1 gpurun_out/r2dr_large.log | cut -c1-400
