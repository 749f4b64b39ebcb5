mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/r2_pytest_final.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_final.log
grep -E "passed|failed|^FAILED|rc=" gpurun_out/r2_pytest_final.log | tail -8
bash scripts/gpu_profile.sh > gpurun_out/r2_profile_run.log 2>&1; tail -12 gpurun_out/r2_profile_run.log
