# round 2, call AD: 8 GPUs - default (configs[2]), S2S training (configs[3]) and long-form inference (configs[4]); lines -> profiles/r02_bench_lines_8gpu.jsonl
set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2ad_large_n8.log 2> gpurun_out/r2ad_large_n8.err
timeout 600 $TR --master-port 29522 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline --workload conmambamamba_large_s2s_fwdbwd_b64x20s > gpurun_out/r2ad_s2s_n8.log 2> gpurun_out/r2ad_s2s_n8.err
timeout 600 $TR --master-port 29523 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline --workload conmamba_large_ctc_infer_b4x300s > gpurun_out/r2ad_cfg5_n8.log 2> gpurun_out/r2ad_cfg5_n8.err
for f in large s2s cfg5; do tail -1 gpurun_out/r2ad_${f}_n8.log | cut -c1-260; tail -c 200 gpurun_out/r2ad_${f}_n8.err; done
