# ncu launch list of the bench command (eager launches so that every kernel is visible to the profiler)
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/final_b_nograph.log 2>&1
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 4500 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/final_ncu3.log 2>&1
ls -la gpurun_out/final_launches.csv
