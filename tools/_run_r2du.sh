# round 2, call DU: cm_layernorm_bwd routed through the quad / staged kernels of fused_ln.cu - full GPU suite, timing, bench A/B
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2du_tests.log 2>&1; tail -3 gpurun_out/r2du_tests.log | cut -c1-300
echo "== routed" > gpurun_out/r2du_prof.log
timeout 300 python tools/prof_elementwise.py 2>&1 | grep -i "layernorm" >> gpurun_out/r2du_prof.log
echo "== pair kernels (CM_LN_NO_ROUTE=1)" >> gpurun_out/r2du_prof.log
CM_LN_NO_ROUTE=1 timeout 300 python tools/prof_elementwise.py 2>&1 | grep -i "layernorm" >> gpurun_out/r2du_prof.log
cat gpurun_out/r2du_prof.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2du_large.log 2> gpurun_out/r2du_large.err
CM_LN_NO_ROUTE=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2du_large_noroute.log 2> gpurun_out/r2du_large_noroute.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2du_large*.log")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"],1), d.get("ms_per_step"), d["kernel_time_share_ms"].get("cm_layernorm_bwd"), d.get("loss"))
    except Exception as e: print(f, "ERR", e)
PY
tail -5 gpurun_out/r2du_large.err | cut -c1-300
