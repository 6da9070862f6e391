# round 2, call DB: N GPUs on the final tree - default training step (configs[2]) and S2S (configs[3]); N from $1
set -x
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29531 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2db_large_n$N.log 2> gpurun_out/r2db_large_n$N.err
timeout 600 $TR --master-port 29532 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --workload conmambamamba_large_s2s_fwdbwd_b64x20s > gpurun_out/r2db_s2s_n$N.log 2> gpurun_out/r2db_s2s_n$N.err
for f in large s2s; do tail -1 gpurun_out/r2db_${f}_n$N.log | cut -c1-200; tail -c 300 gpurun_out/r2db_${f}_n$N.err; done
