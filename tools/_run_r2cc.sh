# round 2, call CC: time segments of scan_bwd_wg: parity (bit-identity), timing per nseg, bench step
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_scan.py -x -q -m gpu -k "warpgroup_backward" > gpurun_out/r2cc_tests.log 2>&1; tail -5 gpurun_out/r2cc_tests.log
for n in 1 2 3 4; do
  CM_SCAN_BWD_NSEG=$n timeout 300 python tools/prof_kernels.py --cfg 3,4 --only scan_bwd > gpurun_out/r2cc_bwd_nseg$n.txt 2>&1; grep scan_bwd gpurun_out/r2cc_bwd_nseg$n.txt
done
timeout 300 python tools/prof_kernels.py --cfg 2,3,4 --only scan_bwd > gpurun_out/r2cc_bwd_auto.txt 2>&1; grep scan_bwd gpurun_out/r2cc_bwd_auto.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2cc_large.log 2> gpurun_out/r2cc_large.err; tail -c 300 gpurun_out/r2cc_large.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2cc_large.log").read().strip().splitlines()[-1])
print(round(d["value"],1), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), d.get("gpu_launches"), d.get("roofline"))
PY
