timeout 900 python -m pytest tests/test_gpu_fused_ln.py tests/test_gpu_models.py -q -m gpu -x 2>&1 | tail -8
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s3_gd_small.log 2> gpurun_out/s3_gd_small.err
timeout 600 python bench.py --steps 5 --warmup 3 --workload conmamba_large_ctc_fwdbwd_b64x20s --no-cpu-baseline > gpurun_out/s3_gd_large.log 2> gpurun_out/s3_gd_large.err
python - <<'PY'
import json
for f in ["s3_gd_small","s3_gd_large"]:
    try:
        d=json.loads(open("gpurun_out/%s.log"%f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["ms_per_step"],2), round(d["e2e"]["value"]), d["gpu_launches"], d["loss"], d["clocks"])
    except Exception as e: print(f, "ERR", e); print(open("gpurun_out/%s.err"%f).read()[-1500:])
PY
