python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -3
ncu --set full --import-source on --clock-control none -k regex:scan_fwd_sp --launch-skip 2 --launch-count 1 -o gpurun_out/scanfwd_sp_final -f python tools/prof_kernels.py --cfg 3 --only scan_fwd --iters 2 > gpurun_out/ncu_sp.log 2>&1
tail -2 gpurun_out/ncu_sp.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2500 -c 2600 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-300
wc -l gpurun_out/launches_r1b.csv
