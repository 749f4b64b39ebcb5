mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_vit_ops_gpu.py -q -k attention -p no:cacheprovider > gpurun_out/r2_attn4_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_attn4_tests.log
tail -5 gpurun_out/r2_attn4_tests.log
timeout 300 python tools/bench_attention.py both > gpurun_out/r2_attn4_bench.log 2>&1
DEPTHMAP_B200_ATTN_FWD3=1 timeout 300 python tools/bench_attention.py both >> gpurun_out/r2_attn4_bench.log 2>&1
cat gpurun_out/r2_attn4_bench.log
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider -s > gpurun_out/r2_pytest_full.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_full.log
grep -E "^\[precision\]|passed|failed|^FAILED|rc=" gpurun_out/r2_pytest_full.log | tail -60
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_beit_a.json 2> gpurun_out/r2_bench_beit_a.err; tail -c 1500 gpurun_out/r2_bench_beit_a.json
