# round 2, call I: full GPU suite after parking the lc kernels behind switches; flat optimizer; sp vs lc forward by shape
set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2i_bench_large.log 2> gpurun_out/r2i_bench_large.err; tail -c 300 gpurun_out/r2i_bench_large.err
for lc in 0 1; do echo "== CM_SCAN_LC=$lc"; CM_SCAN_LC=$lc timeout 300 python tools/prof_kernels.py --cfg 1,2,3,4 --only scan_fwd 2>&1 | cut -c1-150; done
CM_SCAN_LC=1 timeout 300 python tools/prof_kernels.py --cfg 3,4 --only scan_fwd --dtype f32 2>&1 | cut -c1-150
CM_SCAN_LC=0 timeout 300 python tools/prof_kernels.py --cfg 3,4 --only scan_fwd --dtype f32 2>&1 | cut -c1-150
python - <<'PY'
import json
for f in ["r2i_bench_large"]:
    try:
        d=json.loads(open("gpurun_out/%s.log"%f).read().strip().splitlines()[-1])
        r=d.get("roofline") or {}
        print(f, round(d["value"],1), round(d["ms_per_step"],2), round(d["e2e"]["value"],1), d.get("gpu_launches"), r.get("kernel"), r.get("frac"), d["config"]["optimizer"][:50])
        print({k:round(v,2) for k,v in sorted(d["kernel_time_share_ms"].items(), key=lambda x:-x[1])[:14]})
    except Exception as e: print(f, "ERR", e)
PY
