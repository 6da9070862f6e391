# round 2, call DS: staged (bulk-copy) add_ln backward against the register-load form - parity tests, kernel timing, bench
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fused_ln.py -x -q -m gpu > gpurun_out/r2ds_tests.log 2>&1; tail -3 gpurun_out/r2ds_tests.log | cut -c1-300
echo "== staged" > gpurun_out/r2ds_prof.log
timeout 300 python tools/prof_elementwise.py 2>&1 | grep -i "add_ln" >> gpurun_out/r2ds_prof.log
echo "== register loads (CM_ADD_LN_NO_STAGE=1)" >> gpurun_out/r2ds_prof.log
CM_ADD_LN_NO_STAGE=1 timeout 300 python tools/prof_elementwise.py 2>&1 | grep -i "add_ln" >> gpurun_out/r2ds_prof.log
cat gpurun_out/r2ds_prof.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2ds_large.log 2> gpurun_out/r2ds_large.err
CM_ADD_LN_NO_STAGE=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2ds_large_nostage.log 2> gpurun_out/r2ds_large_nostage.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2ds_large*.log")):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(d["value"],1), d.get("ms_per_step"), d["kernel_time_share_ms"].get("cm_add_ln_bwd"))
PY
