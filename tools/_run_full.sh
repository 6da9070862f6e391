timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s3_f2_small.log 2> gpurun_out/s3_f2_small.err
timeout 600 python bench.py --steps 5 --warmup 3 --workload conmamba_large_ctc_fwdbwd_b64x20s --no-cpu-baseline > gpurun_out/s3_f2_large.log 2> gpurun_out/s3_f2_large.err
CM_NO_FUSE_ADD_NORM=1 timeout 600 python bench.py --steps 5 --warmup 3 --workload conmamba_large_ctc_fwdbwd_b64x20s --no-cpu-baseline > gpurun_out/s3_f2_large_nofuse.log 2> gpurun_out/s3_f2_large_nofuse.err
python - <<'PY'
import json
for f in ["s3_f2_small","s3_f2_large","s3_f2_large_nofuse"]:
    try:
        d=json.loads(open("gpurun_out/%s.log"%f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["ms_per_step"],2), round(d["e2e"]["value"]), d["gpu_launches"], d["loss"])
    except Exception as e: print(f, "ERR", e); print(open("gpurun_out/%s.err"%f).read()[-1500:])
PY
