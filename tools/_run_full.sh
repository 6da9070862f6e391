#!/bin/bash
# full GPU validation: the whole -m gpu suite, smoke(), the default bench line
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/full_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/full_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/full_smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/full_bench.log 2>&1
tail -4 gpurun_out/full_tests.log; tail -2 gpurun_out/full_smoke.log; tail -1 gpurun_out/full_bench.log | cut -c1-250
