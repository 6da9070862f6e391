# round 2, call AA: every bench workload on one B200 (lines kept under profiles/r02_bench_lines.jsonl) + smoke + step profile
set -x
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2aa_large.log 2> gpurun_out/r2aa_large.err; tail -c 200 gpurun_out/r2aa_large.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2aa_ref.log 2> gpurun_out/r2aa_ref.err
for wl in conmamba_small_ctc_fwdbwd_b32x15s conmambamamba_large_s2s_fwdbwd_b64x20s conmamba_small_ctc_fwd_b8x10s; do
  timeout 600 python bench.py --steps 10 --warmup 3 --workload $wl --no-cpu-baseline > gpurun_out/r2aa_$wl.log 2> gpurun_out/r2aa_$wl.err
done
timeout 600 python bench.py --steps 5 --warmup 3 --workload conmamba_large_ctc_infer_b4x300s --no-cpu-baseline --sweep-L > gpurun_out/r2aa_cfg5.log 2> gpurun_out/r2aa_cfg5.err
timeout 600 python tools/step_profile.py --top 60 --workload conmamba_large_ctc_fwdbwd_b64x20s > gpurun_out/r2aa_step_large.log 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2aa_*.log")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d.get("roofline") or {}
        print(f.split("r2aa_")[1][:40], round(d["value"],1), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), d.get("gpu_launches"), r.get("kernel"), r.get("frac"), (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e: print(f, "ERR", e)
PY
