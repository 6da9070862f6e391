#!/bin/bash
mkdir -p gpurun_out
echo "== default (quad, 4 CTAs/SM)" > gpurun_out/addln_prof.log
timeout 300 python tools/prof_elementwise.py >> gpurun_out/addln_prof.log 2>&1
echo "== m3 (quad, 3 CTAs/SM, 80 regs)" >> gpurun_out/addln_prof.log
CM_LIB_PATH=$PWD/mamba_asr_b200/lib/libconmamba_b200_m3.so timeout 300 python tools/prof_elementwise.py >> gpurun_out/addln_prof.log 2>&1
echo "== pair kernels" >> gpurun_out/addln_prof.log
CM_ADD_LN_NO_QUAD=1 timeout 300 python tools/prof_elementwise.py >> gpurun_out/addln_prof.log 2>&1
timeout 900 python -m pytest tests/test_gpu_fused_ln.py tests/test_gpu_models.py -q -x > gpurun_out/addln_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/addln_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/addln_bench.log 2>&1
cat gpurun_out/addln_prof.log; tail -15 gpurun_out/addln_tests.log | cut -c1-250; tail -1 gpurun_out/addln_bench.log | cut -c1-200
