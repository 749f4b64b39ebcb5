mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_model_cabi_gpu.py tests/test_zoe_gpu.py tests/test_dav2_gpu.py tests/test_beit_gpu.py tests/test_funnel_gpu.py -q -s -p no:cacheprovider > gpurun_out/r2_cabi.log 2>&1; echo "rc=$?" >> gpurun_out/r2_cabi.log
grep -E "^\[latency\]|passed|failed|^FAILED|Error|rc=|^E  " gpurun_out/r2_cabi.log | tail -40
