timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -8 > gpurun_out/pytest.log
cat gpurun_out/pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench5.log 2> gpurun_out/bench5.err; tail -c 3000 gpurun_out/bench5.log; tail -5 gpurun_out/bench5.err
