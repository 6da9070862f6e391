set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused_ln.py tests/test_gpu_models.py tests/test_gpu_conv_mamba_fbank.py -x -q -m gpu > gpurun_out/r2df_tests.log 2>&1; tail -3 gpurun_out/r2df_tests.log | cut -c1-200
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2df_large.log 2> gpurun_out/r2df_large.err
timeout 600 python tools/step_profile.py --graphed --top 12 > gpurun_out/r2df_step_large.txt 2>&1; grep -i "gelu\|total CUDA\|layernorm_bwd2" gpurun_out/r2df_step_large.txt | cut -c1-150
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2df_large.log").read().strip().splitlines()[-1])
print(round(d["value"],1), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), d.get("gpu_launches"))
PY
