mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_models_baseline_gpu.py tests/test_beit_gpu.py tests/test_zoe_gpu.py tests/test_gpu_parity.py tests/test_funnel_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/r2_pytest_fix.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_fix.log
grep -E "passed|failed|^FAILED|^ERROR|rc=" gpurun_out/r2_pytest_fix.log | tail -6
bash scripts/gpu_profile_a.sh > gpurun_out/r2_profile_a.log 2>&1; tail -12 gpurun_out/r2_profile_a.log
