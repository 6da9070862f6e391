timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/s3_bench_small.log 2> gpurun_out/s3_bench_small.err; tail -c 400 gpurun_out/s3_bench_small.log
timeout 600 python bench.py --steps 5 --warmup 3 --workload conmamba_large_ctc_fwdbwd_b64x20s --no-cpu-baseline > gpurun_out/s3_bench_large.log 2> gpurun_out/s3_bench_large.err; tail -c 300 gpurun_out/s3_bench_large.log
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s3_bench_ref.log 2>&1; tail -c 300 gpurun_out/s3_bench_ref.log
timeout 300 python tools/prof_kernels.py --cfg 2,3 2>&1 | tee gpurun_out/s3_prof.log | cut -c1-160
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
