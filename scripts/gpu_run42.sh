mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_boost_gpu.py tests/test_leres_gpu.py -m gpu -q -x -p no:cacheprovider -k "estimateboost or modelholder or leres_vs_oracle or funnel" > gpurun_out/r2_pytest42.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest42.log
grep -E "passed|failed|^FAILED|^ERROR|rc=" gpurun_out/r2_pytest42.log | tail -6
