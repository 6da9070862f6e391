# round 2, call DD: re-validation of the final tree after the LayerNorm GELU epilogue - whole GPU suite, smoke(), every bench
# workload, graphed step profile, ncu of the LayerNorm kernels with the epilogue
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2dd_tests.log 2>&1; tail -3 gpurun_out/r2dd_tests.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2dd_smoke.log 2>&1; tail -1 gpurun_out/r2dd_smoke.log
timeout 900 python bench.py > gpurun_out/r2dd_large.log 2> gpurun_out/r2dd_large.err; tail -c 200 gpurun_out/r2dd_large.err
for wl in conmamba_small_ctc_fwdbwd_b32x15s conmambamamba_large_s2s_fwdbwd_b64x20s conmamba_small_ctc_fwd_b8x10s; do
  timeout 600 python bench.py --steps 10 --warmup 3 --workload $wl --no-cpu-baseline > gpurun_out/r2dd_$wl.log 2> gpurun_out/r2dd_$wl.err
done
timeout 600 python bench.py --steps 5 --warmup 3 --workload conmamba_large_ctc_infer_b4x300s --no-cpu-baseline --sweep-L > gpurun_out/r2dd_cfg5.log 2> gpurun_out/r2dd_cfg5.err
timeout 600 python tools/step_profile.py --graphed --top 70 > gpurun_out/r2dd_step_large.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --graph-profiling node -k regex:layernorm_bwd2_kernel --launch-skip 40 -c 3 -f -o gpurun_out/r2dd_layernorm_bwd2 python tools/step_profile.py --graphed --top 1 > gpurun_out/r2dd_ncu_ln.log 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2dd_*.log")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d.get("roofline") or {}
        print(f.split("r2dd_")[1][:40], round(d["value"],1), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), d.get("gpu_launches"), r.get("kernel"), r.get("frac"), (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e: pass
PY
