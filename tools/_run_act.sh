#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused_ln.py tests/test_gpu_models.py -q > gpurun_out/act_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/act_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/act_bench.log 2>&1
CM_DROPOUT_STORE_MASK=1 CM_NO_FUSE_BIAS_GRAD=1 timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/act_bench_old.log 2>&1
tail -25 gpurun_out/act_tests.log | cut -c1-300; tail -1 gpurun_out/act_bench.log | cut -c1-200;  tail -1 gpurun_out/act_bench_old.log | cut -c1-200
