timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/s3_8gpu.log 2> gpurun_out/s3_8gpu.err; tail -c 400 gpurun_out/s3_8gpu.log; tail -2 gpurun_out/s3_8gpu.err | cut -c1-200
python - <<'PY'
import json
d=json.loads(open("gpurun_out/s3_8gpu.log").read().strip().splitlines()[-1])
print(d["n_gpus"], round(d["value"]), round(d["ms_per_step"],2), round(d["e2e"]["value"]), d["config"]["parallelism"], d["clocks"])
PY
