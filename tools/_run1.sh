timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4
for wl in conmamba_small_ctc_fwdbwd_b32x15s conmamba_large_ctc_fwdbwd_b64x20s; do
timeout 900 python bench.py --steps 8 --warmup 3 --workload $wl --no-cpu-baseline > gpurun_out/bench_x.log 2> gpurun_out/bench_x.err; python - <<EOF2
import json
l=[x for x in open("gpurun_out/bench_x.log") if x.startswith("{")]
d=json.loads(l[-1]); print({k:d[k] for k in ("value","ms_per_step","gpu_launches","loss")})
EOF2
done
