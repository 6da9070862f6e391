set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused_ln.py tests/test_gpu_models.py tests/test_gpu_deferred_reduce.py -x -q -m gpu > gpurun_out/r2dc_tests.log 2>&1; tail -5 gpurun_out/r2dc_tests.log | cut -c1-200
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2dc_large.log 2> gpurun_out/r2dc_large.err
CM_NO_LN_GELU_EPILOGUE=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2dc_large_noepi.log 2> gpurun_out/r2dc_large_noepi.err
timeout 600 python tools/step_profile.py --graphed --top 70 > gpurun_out/r2dc_step_large.txt 2>&1; grep -i "layernorm\|gelu\|total CUDA" gpurun_out/r2dc_step_large.txt | cut -c1-150
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2dc_large*.log")):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(d["value"],1), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), d.get("gpu_launches"))
PY
