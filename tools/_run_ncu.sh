set -x
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_bwd_sp --launch-skip 1 --launch-count 1 -o gpurun_out/s3_scanbwd_cfg3 -f python tools/prof_kernels.py --cfg 3 --only scan_bwd --iters 2 > gpurun_out/s3_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"add_ln|gelu_dropout" --launch-skip 8 --launch-count 4 -o gpurun_out/s3_fused_cfg3 -f python tools/prof_aux2.py > gpurun_out/s3_ncu2.log 2>&1
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/s3_b_nograph.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/s3_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/s3_ncu3.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/s3_launches.csv
