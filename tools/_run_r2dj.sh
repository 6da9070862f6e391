set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused_ln.py tests/test_gpu_models.py -x -q -m gpu > gpurun_out/r2dj_tests.log 2>&1; tail -3 gpurun_out/r2dj_tests.log | cut -c1-200
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2dj_large.log 2> gpurun_out/r2dj_large.err
CM_DROPOUT_REGEN=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2dj_large_regen.log 2> gpurun_out/r2dj_large_regen.err
timeout 600 python tools/step_profile.py --graphed --top 12 > gpurun_out/r2dj_step_large.txt 2>&1; grep -i "gelu\|total CUDA" gpurun_out/r2dj_step_large.txt | cut -c1-150
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2dj_large*.log")):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(d["value"],1), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), d.get("gpu_launches"))
PY
