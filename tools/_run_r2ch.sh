set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_deferred_reduce.py tests/test_gpu_models.py -x -q -m gpu > gpurun_out/r2ch_tests.log 2>&1; tail -3 gpurun_out/r2ch_tests.log
timeout 600 python tools/step_profile.py --graphed --top 60 > gpurun_out/r2ch_step_large.txt 2>&1; grep -i "reduce\|total CUDA" gpurun_out/r2ch_step_large.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2ch_large.log 2> gpurun_out/r2ch_large.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2ch_large*.log")):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(d["value"],1), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), d.get("gpu_launches"))
PY
