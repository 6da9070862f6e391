# round 2, call A: validate phase-0 changes on one B200 and take the baseline numbers of the new default workload
set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench_large.log 2> gpurun_out/r2a_bench_large.err; tail -c 400 gpurun_out/r2a_bench_large.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-optimizer --no-cpu-baseline > gpurun_out/r2a_bench_large_noopt.log 2> gpurun_out/r2a_bench_large_noopt.err
timeout 600 python bench.py --steps 10 --warmup 3 --workload conmamba_small_ctc_fwd_b8x10s --no-cpu-baseline > gpurun_out/r2a_bench_cfg1.log 2> gpurun_out/r2a_bench_cfg1.err; tail -c 300 gpurun_out/r2a_bench_cfg1.err
timeout 600 python bench.py --steps 5 --warmup 3 --workload conmamba_large_ctc_infer_b4x300s --no-cpu-baseline --sweep-L > gpurun_out/r2a_bench_cfg5.log 2> gpurun_out/r2a_bench_cfg5.err; tail -c 300 gpurun_out/r2a_bench_cfg5.err
timeout 600 python bench.py --steps 5 --warmup 3 --workload conmambamamba_large_s2s_fwdbwd_b64x20s --no-cpu-baseline > gpurun_out/r2a_bench_s2s.log 2> gpurun_out/r2a_bench_s2s.err; tail -c 300 gpurun_out/r2a_bench_s2s.err
timeout 300 python tools/prof_kernels.py --cfg 2,3 --only scan_fwd,scan_bwd,conv_fwd,conv_bwd > gpurun_out/r2a_prof.log 2>&1
python - <<'PY'
import json
for f in ["r2a_bench_large","r2a_bench_large_noopt","r2a_bench_cfg1","r2a_bench_cfg5","r2a_bench_s2s"]:
    try:
        d=json.loads(open("gpurun_out/%s.log"%f).read().strip().splitlines()[-1])
        r=d.get("roofline") or {}
        print(f, round(d["value"],1), round(d["ms_per_step"],2), round(d["e2e"]["value"],1), d.get("gpu_launches"), r.get("kernel"), r.get("frac"), d["config"]["launch"][:60])
        if "scan_sweep" in d:
            for e in d["scan_sweep"]: print("   ", e["dtype"], e["L"], round(e["ms"],3), round(e["frac"],3))
    except Exception as e: print(f, "ERR", e)
PY
cut -c1-170 gpurun_out/r2a_prof.log
