# round 2, call CD: GLU fused into the depthwise-conv kernels: parity, kernel timing, bench step A/B
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv_mamba_fbank.py tests/test_gpu_models.py tests/test_gpu_deferred_reduce.py -x -q -m gpu > gpurun_out/r2cd_tests.log 2>&1; tail -8 gpurun_out/r2cd_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2cd_large.log 2> gpurun_out/r2cd_large.err; tail -c 300 gpurun_out/r2cd_large.err
CM_NO_FUSE_GLU=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2cd_large_noglu.log 2> gpurun_out/r2cd_large_noglu.err
timeout 600 python tools/step_profile.py --graphed --top 45 > gpurun_out/r2cd_step_large.txt 2>&1; grep -i "dwconv\|glu\|total CUDA\|reduce" gpurun_out/r2cd_step_large.txt
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2cd_large*.log")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"],1), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), d.get("gpu_launches"))
    except Exception as e: print(f, "ERR", str(e)[:60])
PY
