# round 2, call S: forward wg tests + large step profile
set -x
timeout 900 python -m pytest tests/test_gpu_scan.py -m gpu -x -q -k "warpgroup" 2>&1 | tail -4
timeout 600 python tools/step_profile.py --top 70 --workload conmamba_large_ctc_fwdbwd_b64x20s > gpurun_out/r2s_step_large.log 2>&1
head -75 gpurun_out/r2s_step_large.log | cut -c1-150
