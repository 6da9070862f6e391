timeout 600 python tools/step_profile.py --top 70 --workload conmamba_large_ctc_fwdbwd_b64x20s > gpurun_out/final_step_large.log 2>&1
timeout 600 python tools/step_profile.py --top 70 > gpurun_out/final_step_small.log 2>&1
timeout 600 python tools/step_profile.py --top 50 --workload conmambamamba_large_s2s_fwdbwd_b64x20s > gpurun_out/final_step_s2s.log 2>&1
head -3 gpurun_out/final_step_*.log | cut -c1-120
