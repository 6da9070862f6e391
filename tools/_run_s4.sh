# session 4: batched A = -exp(A_log) - full GPU suite, smoke, step timing
timeout 300 python -m pytest tests -x -q -m gpu > gpurun_out/s4c_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s4c_pytest.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s4c_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/s4c_smoke.log
for wl in conmamba_small_ctc_fwdbwd_b32x15s conmamba_large_ctc_fwdbwd_b64x20s; do
  timeout 100 python bench.py --workload $wl --no-cpu-baseline > gpurun_out/s4c_$wl.log 2> gpurun_out/s4c_$wl.err; echo "bench $wl rc=$?"
  python - <<P
import json
for f in ("gpurun_out/s4c_$wl.log",):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, "ms/step %.3f  value %.0f  e2e %.0f  launches %d" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["gpu_launches"]))
    except Exception as e:
        print(f, "FAILED", e)
P
done
