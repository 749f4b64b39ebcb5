mkdir -p gpurun_out

timeout 1200 python -m pytest tests/test_boost_gpu.py -m gpu -q -s -p no:cacheprovider > gpurun_out/r2_pytest21.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest21.log
grep -E "precision|passed|failed|^FAILED|^ERROR|rc=|Error" gpurun_out/r2_pytest21.log | tail -20
