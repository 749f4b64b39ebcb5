mkdir -p gpurun_out
for P in 0 1 2; do
  DEPTHMAP_B200_ATTN_POLY=$P timeout 300 python tools/bench_attention.py both 2>&1 | sed "s/^/poly=$P /" >> gpurun_out/r2_attn4f_bench.log
done
DEPTHMAP_B200_ATTN_POLY=1 timeout 600 python -m pytest tests/test_vit_ops_gpu.py -q -k attention -p no:cacheprovider > gpurun_out/r2_attn4f_tests1.log 2>&1; tail -2 gpurun_out/r2_attn4f_tests1.log
DEPTHMAP_B200_ATTN_POLY=2 timeout 600 python -m pytest tests/test_vit_ops_gpu.py -q -k attention -p no:cacheprovider > gpurun_out/r2_attn4f_tests2.log 2>&1; tail -2 gpurun_out/r2_attn4f_tests2.log
cat gpurun_out/r2_attn4f_bench.log
