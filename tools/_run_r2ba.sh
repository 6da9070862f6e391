# round 2, call BA: final evidence on one B200 - every bench workload, the reference arm, the step profile, the ncu launch
# list of the eager bench command and `ncu --set full` captures of the kernels added in this session
set -x
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2ba_large.log 2> gpurun_out/r2ba_large.err; tail -c 200 gpurun_out/r2ba_large.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2ba_ref.log 2> gpurun_out/r2ba_ref.err
for wl in conmamba_small_ctc_fwdbwd_b32x15s conmambamamba_large_s2s_fwdbwd_b64x20s conmamba_small_ctc_fwd_b8x10s; do
  timeout 600 python bench.py --steps 10 --warmup 3 --workload $wl --no-cpu-baseline > gpurun_out/r2ba_$wl.log 2> gpurun_out/r2ba_$wl.err
done
timeout 600 python bench.py --steps 5 --warmup 3 --workload conmamba_large_ctc_infer_b4x300s --no-cpu-baseline --sweep-L > gpurun_out/r2ba_cfg5.log 2> gpurun_out/r2ba_cfg5.err
timeout 600 python tools/step_profile.py --top 70 --workload conmamba_large_ctc_fwdbwd_b64x20s > gpurun_out/r2ba_step_large.txt 2>&1
timeout 600 python tools/step_profile.py --top 50 --workload conmambamamba_large_s2s_fwdbwd_b64x20s > gpurun_out/r2ba_step_s2s.txt 2>&1
timeout 300 python tools/prof_elementwise.py > gpurun_out/r2ba_elementwise.txt 2>&1
timeout 300 python tools/prof_ctc.py > gpurun_out/r2ba_ctc.txt 2>&1
# ncu launch list of the eager bench command (after it exited 0 without ncu)
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/r2ba_nograph.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r2ba_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/r2ba_ncu_list.log 2>&1
for k in ctc_kernel stem_fwd_kernel stem_bwd_kernel add_ln_fwd_q_kernel add_ln_bwd_q_kernel gelu_dropout_bwd_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 3 -c 1 -f -o gpurun_out/r2ba_$k python tools/step_profile.py --top 1 > gpurun_out/r2ba_ncu_$k.log 2>&1
done
ls -la gpurun_out/r2ba_*.ncu-rep gpurun_out/r2ba_launches.csv
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2ba_*.log")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d.get("roofline") or {}
        print(f.split("r2ba_")[1][:40], round(d["value"],1), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), d.get("gpu_launches"), r.get("kernel"), r.get("frac"), (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e: print(f, "ERR", str(e)[:60])
PY
