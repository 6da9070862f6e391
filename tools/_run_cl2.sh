timeout 900 python -m pytest tests/test_gpu_models.py -q -m gpu -x 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/s3_cl2_small.log 2> gpurun_out/s3_cl2_small.err
timeout 600 python bench.py --steps 8 --warmup 3 --workload conmamba_large_ctc_fwdbwd_b64x20s --no-cpu-baseline > gpurun_out/s3_cl2_large.log 2> gpurun_out/s3_cl2_large.err
python - <<'PY'
import json
for f in ["s3_cl2_small","s3_cl2_large"]:
    try:
        d=json.loads(open("gpurun_out/%s.log"%f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["ms_per_step"],2), round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],2), d["gpu_launches"], d["loss"])
    except Exception as e: print(f, "ERR", e); print(open("gpurun_out/%s.err"%f).read()[-1500:])
PY
timeout 600 python tools/step_profile.py --top 60 --workload conmamba_large_ctc_fwdbwd_b64x20s > gpurun_out/s3_prof_large_cl2.log 2>&1; grep -v Warn gpurun_out/s3_prof_large_cl2.log | grep -i "nchw\|nhwc\|direct_copy\|cudnn\|total\|GammaBeta\|layer_norm" | cut -c1-150
