mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_boost_gpu.py tests/test_funnel_gpu.py -m gpu -q -p no:cacheprovider -k "funnel" > gpurun_out/r2_pytest38.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest38.log
grep -E "passed|failed|^FAILED|^ERROR|rc=|Error" gpurun_out/r2_pytest38.log | tail -8
