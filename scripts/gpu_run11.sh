mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_vit_ops_gpu.py -q -k attention -p no:cacheprovider -x > gpurun_out/r2_attn4d_tests.log 2>&1; tail -4 gpurun_out/r2_attn4d_tests.log
DEPTHMAP_B200_ATTN_PTMEM=0 timeout 600 python -m pytest tests/test_vit_ops_gpu.py -q -k attention -p no:cacheprovider > gpurun_out/r2_attn4d_tests_smem.log 2>&1; tail -2 gpurun_out/r2_attn4d_tests_smem.log
timeout 300 python tools/bench_attention.py both > gpurun_out/r2_attn4d_bench.log 2>&1
DEPTHMAP_B200_ATTN_PTMEM=0 timeout 300 python tools/bench_attention.py both >> gpurun_out/r2_attn4d_bench.log 2>&1
cat gpurun_out/r2_attn4d_bench.log
