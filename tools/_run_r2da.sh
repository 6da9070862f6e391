# round 2, call DA: evidence on the final tree, one B200 - the whole GPU suite, smoke(), every bench workload, the reference
# arm, the step profiles of the graphed step, the ncu launch list of the eager bench command
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2da_tests.log 2>&1; tail -3 gpurun_out/r2da_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2da_smoke.log 2>&1; tail -2 gpurun_out/r2da_smoke.log
timeout 900 python bench.py > gpurun_out/r2da_large.log 2> gpurun_out/r2da_large.err; tail -c 200 gpurun_out/r2da_large.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2da_ref.log 2> gpurun_out/r2da_ref.err
for wl in conmamba_small_ctc_fwdbwd_b32x15s conmambamamba_large_s2s_fwdbwd_b64x20s conmamba_small_ctc_fwd_b8x10s; do
  timeout 600 python bench.py --steps 10 --warmup 3 --workload $wl --no-cpu-baseline > gpurun_out/r2da_$wl.log 2> gpurun_out/r2da_$wl.err
done
timeout 600 python bench.py --steps 5 --warmup 3 --workload conmamba_large_ctc_infer_b4x300s --no-cpu-baseline --sweep-L > gpurun_out/r2da_cfg5.log 2> gpurun_out/r2da_cfg5.err
timeout 600 python tools/step_profile.py --graphed --top 70 --workload conmamba_large_ctc_fwdbwd_b64x20s > gpurun_out/r2da_step_large.txt 2>&1
timeout 600 python tools/step_profile.py --graphed --top 50 --workload conmambamamba_large_s2s_fwdbwd_b64x20s > gpurun_out/r2da_step_s2s.txt 2>&1
timeout 300 python tools/prof_kernels.py --cfg 2,3,4 > gpurun_out/r2da_kernels.txt 2>&1
timeout 300 python tools/prof_kernels.py --cfg 3 --only aux > gpurun_out/r2da_aux.txt 2>&1
# ncu launch list of the eager bench command (after it exited 0 without ncu)
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/r2da_nograph.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r2da_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/r2da_ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --graph-profiling node -k regex:reduce_batch_kernel --launch-skip 30 -c 4 -f -o gpurun_out/r2da_reduce_batch_v2 python tools/step_profile.py --graphed --top 1 > gpurun_out/r2da_ncu_reduce.log 2>&1
ls -la gpurun_out/r2da_launches.csv gpurun_out/r2da_*.ncu-rep
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2da_*.log")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d.get("roofline") or {}
        print(f.split("r2da_")[1][:40], round(d["value"],1), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), d.get("gpu_launches"), r.get("kernel"), r.get("frac"), (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e: print(f, "ERR", str(e)[:60])
PY
