timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -8 > gpurun_out/pytest.log
cat gpurun_out/pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench9.log 2> gpurun_out/bench9.err; python - <<EOF2
import json
l=[x for x in open("gpurun_out/bench9.log") if x.startswith("{")]
d=json.loads(l[-1]); print({k:d[k] for k in ("value","ms_per_step","e2e","gpu_launches","loss")}); print(d["roofline"])
EOF2
tail -2 gpurun_out/bench9.err | cut -c1-200
timeout 900 python bench.py --steps 5 --warmup 3 --workload conmamba_large_ctc_fwdbwd_b64x20s --no-cpu-baseline > gpurun_out/bench_large.log 2> gpurun_out/bench_large.err; python - <<EOF2
import json
l=[x for x in open("gpurun_out/bench_large.log") if x.startswith("{")]
d=json.loads(l[-1]); print({k:d[k] for k in ("value","ms_per_step","e2e","gpu_launches","loss")}); print(d["kernel_time_share_ms"])
EOF2
