# round 2, call DV: last validation - smoke(), default bench with the CPU baseline, the reference arm, small / S2S workloads
set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2dv_smoke.log 2>&1; tail -1 gpurun_out/r2dv_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2dv_large.log 2> gpurun_out/r2dv_large.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2dv_ref.log 2> gpurun_out/r2dv_ref.err
for wl in conmamba_small_ctc_fwdbwd_b32x15s conmambamamba_large_s2s_fwdbwd_b64x20s; do
  timeout 600 python bench.py --steps 10 --warmup 3 --workload $wl --no-cpu-baseline > gpurun_out/r2dv_$wl.log 2> gpurun_out/r2dv_$wl.err
done
timeout 600 python tools/step_profile.py --graphed --top 40 > gpurun_out/r2dv_step_large.txt 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2dv_*.log")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d.get("roofline") or {}
        print(f.split("r2dv_")[1][:40], round(d["value"],1), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), d.get("gpu_launches"), r.get("kernel"), r.get("frac"), (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e: print(f, "ERR", str(e)[:60])
PY
