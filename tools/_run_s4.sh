# session 4: ncu launch list of the bench command (eager launches) + one full capture of the ln_act kernels
timeout 165 ncu --metrics gpu__time_duration.sum --clock-control none -c 4500 --csv --log-file gpurun_out/s4_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/s4_ncu_launch.log 2>&1; echo "launch list rc=$?"
ls -la gpurun_out/s4_launches.csv
timeout 40 ncu --set full --clock-control none --import-source on -k regex:ln_act -c 9 -f -o gpurun_out/s4_lnact python tools/prof_kernels.py --cfg 3 --only ln_act --iters 1 > gpurun_out/s4_ncu_lnact.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/s4_lnact.ncu-rep
