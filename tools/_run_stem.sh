#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused_ln.py -q -k "stem or front_end" > gpurun_out/stem_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/stem_tests.log
timeout 300 python tools/prof_stem.py > gpurun_out/stem_prof.log 2>&1
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_fused_ln.py -q -x > gpurun_out/stem_tests2.log 2>&1; tail -3 gpurun_out/stem_tests2.log
tail -25 gpurun_out/stem_tests.log; cat gpurun_out/stem_prof.log
