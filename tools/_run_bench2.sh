timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s3_fl_small.log 2> gpurun_out/s3_fl_small.err
timeout 600 python bench.py --steps 5 --warmup 3 --workload conmamba_large_ctc_fwdbwd_b64x20s --no-cpu-baseline > gpurun_out/s3_fl_large.log 2> gpurun_out/s3_fl_large.err
python - <<'PY'
import json
for f in ["s3_fl_small","s3_fl_large"]:
    try:
        d=json.loads(open("gpurun_out/%s.log"%f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["ms_per_step"],2), round(d["e2e"]["value"]), d["gpu_launches"], d["loss"], d["config"]["launch"][:40])
    except Exception as e: print(f, "ERR", e); print(open("gpurun_out/%s.err"%f).read()[-1500:])
PY
timeout 600 python tools/step_profile.py --top 30 --workload conmamba_large_ctc_fwdbwd_b64x20s > gpurun_out/s3_prof_large_fl.log 2>&1; grep -v Warn gpurun_out/s3_prof_large_fl.log | head -34 | cut -c1-150
