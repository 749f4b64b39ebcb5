bash scripts/gpu_profile_b.sh > gpurun_out/r2_profile_b.log 2>&1; tail -8 gpurun_out/r2_profile_b.log
bash scripts/gpu_sanitize.sh > gpurun_out/r2_sanitize.log 2>&1; tail -14 gpurun_out/r2_sanitize.log
