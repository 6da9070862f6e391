# Round-end validation on one B200: GPU tests, smoke, the three bench workloads, the reference arm, kernel timings.
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/final_bench_small.log 2> gpurun_out/final_bench_small.err; tail -c 300 gpurun_out/final_bench_small.log
timeout 600 python bench.py --steps 8 --warmup 3 --workload conmamba_large_ctc_fwdbwd_b64x20s --no-cpu-baseline > gpurun_out/final_bench_large.log 2> gpurun_out/final_bench_large.err
timeout 600 python bench.py --steps 5 --warmup 3 --workload conmambamamba_large_s2s_fwdbwd_b64x20s --no-cpu-baseline > gpurun_out/final_bench_s2s.log 2> gpurun_out/final_bench_s2s.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_ref.log 2>&1
timeout 300 python tools/prof_kernels.py --cfg 2,3 --only scan_fwd,scan_bwd,conv_fwd,conv_bwd,aux > gpurun_out/final_prof.log 2>&1
python - <<'PY'
import json
for f in ["final_bench_small","final_bench_large","final_bench_s2s","final_bench_ref"]:
    try:
        d=json.loads(open("gpurun_out/%s.log"%f).read().strip().splitlines()[-1])
        r=d.get("roofline") or {}
        print(f, round(d["value"],1), round(d["ms_per_step"],2), round(d["e2e"]["value"],1), d.get("gpu_launches"), r.get("kernel"), r.get("frac"), d.get("clocks"))
    except Exception as e: print(f, "ERR", e)
PY
cut -c1-150 gpurun_out/final_prof.log
