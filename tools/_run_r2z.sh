# round 2, call Z: full GPU suite + default bench after the conv / dwconv / fbank kernels
set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2z_bench_large.log 2> gpurun_out/r2z_bench_large.err; tail -c 300 gpurun_out/r2z_bench_large.err
python - <<'PY'
import json
for f in ["r2z_bench_large"]:
    try:
        d=json.loads(open("gpurun_out/%s.log"%f).read().strip().splitlines()[-1])
        r=d.get("roofline") or {}
        print(f, round(d["value"],1), round(d["ms_per_step"],2), round(d["e2e"]["value"],1), d.get("gpu_launches"), r.get("kernel"), r.get("frac"), r.get("traffic"))
        print({k:round(v,2) for k,v in sorted(d["kernel_time_share_ms"].items(), key=lambda x:-x[1])[:14]})
    except Exception as e: print(f, "ERR", e)
PY
