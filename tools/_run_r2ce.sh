set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv_mamba_fbank.py -x -q -m gpu -k "depthwise" > gpurun_out/r2ce_tests.log 2>&1; tail -3 gpurun_out/r2ce_tests.log
timeout 600 python tools/step_profile.py --graphed --top 45 > gpurun_out/r2ce_step_large.txt 2>&1; grep -i "dwconv\|glu\|total CUDA" gpurun_out/r2ce_step_large.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2ce_large.log 2> gpurun_out/r2ce_large.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2ce_large.log").read().strip().splitlines()[-1])
print(round(d["value"],1), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), d.get("gpu_launches"))
PY
