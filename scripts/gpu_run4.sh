mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_vit_ops_gpu.py -q -k attention -p no:cacheprovider > gpurun_out/r2_attn4b_tests.log 2>&1; tail -3 gpurun_out/r2_attn4b_tests.log
timeout 300 python tools/bench_attention.py both > gpurun_out/r2_attn4b_bench.log 2>&1; cat gpurun_out/r2_attn4b_bench.log
timeout 900 python -m pytest tests/test_zoe_gpu.py tests/test_video_gpu.py -q -s -p no:cacheprovider > gpurun_out/r2_zoe2.log 2>&1; echo "rc=$?" >> gpurun_out/r2_zoe2.log
grep -E "^\[precision\]|passed|failed|^FAILED|Error|rc=" gpurun_out/r2_zoe2.log | tail -30
timeout 600 python -m pytest tests/test_funnel_gpu.py -q -p no:cacheprovider > gpurun_out/r2_funnel.log 2>&1; tail -15 gpurun_out/r2_funnel.log
