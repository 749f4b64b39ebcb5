mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_vit_ops_gpu.py -q -k attention -p no:cacheprovider > gpurun_out/r2_attn4c_tests.log 2>&1; tail -3 gpurun_out/r2_attn4c_tests.log
timeout 300 python tools/bench_attention.py both > gpurun_out/r2_attn4c_bench.log 2>&1
DEPTHMAP_B200_ATTN_TOKEN=0 timeout 300 python tools/bench_attention.py both >> gpurun_out/r2_attn4c_bench.log 2>&1
cat gpurun_out/r2_attn4c_bench.log
