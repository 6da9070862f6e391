# session 4: ln_act with GELU / pre-norm bias - parity and step timing
timeout 250 python -m pytest tests/test_gpu_fused_ln.py tests/test_gpu_models.py -x -q -m gpu > gpurun_out/s4b_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/s4b_pytest.log
for wl in conmamba_small_ctc_fwdbwd_b32x15s conmamba_large_ctc_fwdbwd_b64x20s; do
  timeout 100 python bench.py --workload $wl --no-cpu-baseline > gpurun_out/s4b_$wl.log 2> gpurun_out/s4b_$wl.err; echo "bench $wl rc=$?"
  python - <<P
import json
for f in ("gpurun_out/s4b_$wl.log",):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, "ms/step %.3f  value %.0f  e2e %.0f  launches %d" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["gpu_launches"]))
    except Exception as e:
        print(f, "FAILED", e)
P
done
