# session 4 final: full GPU suite, smoke, bench lines (all workloads + reference arm), step profiles, kernel timings, ncu launch list
timeout 300 python -m pytest tests -x -q -m gpu > gpurun_out/s4d_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s4d_pytest.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s4d_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/s4d_smoke.log
timeout 150 python bench.py > gpurun_out/s4d_bench_small.log 2> gpurun_out/s4d_bench_small.err; echo "bench small rc=$?"
CM_NO_WGRAD_SPLIT=1 timeout 100 python bench.py --no-cpu-baseline > gpurun_out/s4d_bench_small_nosplit.log 2>/dev/null; echo "bench small nosplit rc=$?"
timeout 100 python bench.py --workload conmamba_large_ctc_fwdbwd_b64x20s --no-cpu-baseline > gpurun_out/s4d_bench_large.log 2> gpurun_out/s4d_bench_large.err; echo "bench large rc=$?"
CM_NO_WGRAD_SPLIT=1 timeout 100 python bench.py --workload conmamba_large_ctc_fwdbwd_b64x20s --no-cpu-baseline > gpurun_out/s4d_bench_large_nosplit.log 2>/dev/null; echo "bench large nosplit rc=$?"
timeout 100 python bench.py --workload conmambamamba_large_s2s_fwdbwd_b64x20s --no-cpu-baseline > gpurun_out/s4d_bench_s2s.log 2> gpurun_out/s4d_bench_s2s.err; echo "bench s2s rc=$?"
timeout 100 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s4d_bench_ref.log 2>&1; echo "bench ref rc=$?"
python - <<P
import json, glob
for f in sorted(glob.glob("gpurun_out/s4d_bench_*.log")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, "ms/step %.3f  value %.1f  e2e %.1f" % (d["ms_per_step"], d["value"], d["e2e"]["value"]), d.get("roofline", {}).get("frac"))
    except Exception as e:
        print(f, "FAILED", e)
P
timeout 120 python tools/step_profile.py --top 90 --workload conmamba_large_ctc_fwdbwd_b64x20s > gpurun_out/s4d_step_large.log 2>&1
timeout 120 python tools/step_profile.py --top 90 > gpurun_out/s4d_step_small.log 2>&1
timeout 120 python tools/step_profile.py --top 60 --workload conmambamamba_large_s2s_fwdbwd_b64x20s > gpurun_out/s4d_step_s2s.log 2>&1
grep -h "total CUDA" gpurun_out/s4d_step_*.log
