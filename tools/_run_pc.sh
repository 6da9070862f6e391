timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_conv_mamba_fbank.py -q -m gpu -x 2>&1 | tail -5
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s3_pc_small.log 2> gpurun_out/s3_pc_small.err; tail -c 300 gpurun_out/s3_pc_small.log; tail -3 gpurun_out/s3_pc_small.err | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-param-cache > gpurun_out/s3_nopc_small.log 2> gpurun_out/s3_nopc_small.err
timeout 600 python bench.py --steps 5 --warmup 3 --workload conmamba_large_ctc_fwdbwd_b64x20s --no-cpu-baseline > gpurun_out/s3_pc_large.log 2> gpurun_out/s3_pc_large.err
python - <<'PY'
import json
for f in ["s3_pc_small","s3_nopc_small","s3_pc_large"]:
    try:
        d=json.loads(open("gpurun_out/%s.log"%f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["ms_per_step"],2), round(d["e2e"]["value"]), d["gpu_launches"], round(d["kernel_time_share_ms"].get("cm_reduce_multi",0)/3,2), d["loss"])
    except Exception as e: print(f, "ERR", e)
PY
