timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -8 > gpurun_out/pytest.log
cat gpurun_out/pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench7.log 2> gpurun_out/bench7.err; python - <<EOF2
import json
l=[x for x in open("gpurun_out/bench7.log") if x.startswith("{")]
d=json.loads(l[-1]); print({k:d[k] for k in ("value","ms_per_step","e2e","gpu_launches","loss","kernel_time_share_ms")}); print(d["config"]["launch"])
EOF2
tail -3 gpurun_out/bench7.err | cut -c1-200
timeout 300 python tools/step_profile.py --top 14 2>&1 | tail -16
