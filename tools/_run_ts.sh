timeout 900 python -m pytest tests/test_gpu_conv_mamba_fbank.py tests/test_gpu_models.py tests/test_gpu_fused_ln.py -q -m gpu -x 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/s3_ts2_small.log 2> gpurun_out/s3_ts2_small.err
CM_NO_TSMM=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/s3_ts2_small_off.log 2> gpurun_out/s3_ts2_small_off.err
timeout 600 python bench.py --steps 8 --warmup 3 --workload conmamba_large_ctc_fwdbwd_b64x20s --no-cpu-baseline > gpurun_out/s3_ts2_large.log 2> gpurun_out/s3_ts2_large.err
CM_NO_TSMM=1 timeout 600 python bench.py --steps 8 --warmup 3 --workload conmamba_large_ctc_fwdbwd_b64x20s --no-cpu-baseline > gpurun_out/s3_ts2_large_off.log 2> gpurun_out/s3_ts2_large_off.err
python - <<'PY'
import json
for f in ["s3_ts2_small","s3_ts2_small_off","s3_ts2_large","s3_ts2_large_off"]:
    try:
        d=json.loads(open("gpurun_out/%s.log"%f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["ms_per_step"],2), round(d["e2e"]["value"]), d["gpu_launches"], d["loss"])
    except Exception as e: print(f, "ERR", e); print(open("gpurun_out/%s.err"%f).read()[-1500:])
PY
