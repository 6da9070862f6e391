timeout 600 python tools/step_profile.py --top 45 > gpurun_out/s3_prof_small_pc.log 2>&1; grep -v Warn gpurun_out/s3_prof_small_pc.log | head -50 | cut -c1-170
