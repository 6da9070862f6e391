#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ctc.py -q > gpurun_out/ctc_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/ctc_tests.log
timeout 300 python tools/prof_ctc.py > gpurun_out/ctc_prof.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/ctc_bench.log 2>&1
tail -15 gpurun_out/ctc_tests.log; cat gpurun_out/ctc_prof.log; tail -1 gpurun_out/ctc_bench.log | cut -c1-400
