# round 2, call CB: batched / deferred reductions: parity, then the bench step with and without
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_deferred_reduce.py -x -q -m gpu > gpurun_out/r2cb_tests.log 2>&1; tail -15 gpurun_out/r2cb_tests.log
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_fused_ln.py tests/test_gpu_conv_mamba_fbank.py -x -q -m gpu > gpurun_out/r2cb_tests2.log 2>&1; tail -5 gpurun_out/r2cb_tests2.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2cb_large.log 2> gpurun_out/r2cb_large.err; tail -c 300 gpurun_out/r2cb_large.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-defer-reduce > gpurun_out/r2cb_large_nodefer.log 2> gpurun_out/r2cb_large_nodefer.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2cb_*.log")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"],1), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), d.get("gpu_launches"))
    except Exception as e: print(f, "ERR", str(e)[:60])
PY
