timeout 600 python tools/step_profile.py --top 200 --workload conmamba_large_ctc_fwdbwd_b64x20s > gpurun_out/s3_prof_large_A.log 2>&1
CM_NO_FUSE_ADD_NORM=1 timeout 600 python tools/step_profile.py --top 200 --workload conmamba_large_ctc_fwdbwd_b64x20s > gpurun_out/s3_prof_large_B.log 2>&1
head -3 gpurun_out/s3_prof_large_A.log gpurun_out/s3_prof_large_B.log | cut -c1-120
