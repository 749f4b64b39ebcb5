mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_boost_gpu.py tests/test_leres_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/r2_pytest41.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest41.log
grep -E "passed|failed|^FAILED|^ERROR|rc=" gpurun_out/r2_pytest41.log | tail -6
timeout 300 python tools/bench_boost_parts.py > gpurun_out/r02_boost_parts_v4.txt 2>&1; grep -E "LeReS|one patch" gpurun_out/r02_boost_parts_v4.txt
