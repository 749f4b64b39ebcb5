mkdir -p gpurun_out
timeout 300 python tools/bench_attention.py both > gpurun_out/r2_attn_bench18.log 2>&1; cat gpurun_out/r2_attn_bench18.log | tail -4
timeout 1500 python -m pytest tests/test_zoe_gpu.py tests/test_vit_ops_gpu.py tests/test_beit_gpu.py tests/test_dav2_gpu.py tests/test_models_baseline_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/r2_pytest18.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest18.log
grep -E "passed|failed|^FAILED|^ERROR|rc=" gpurun_out/r2_pytest18.log | tail -8
timeout 900 python bench.py --workload zoedepth_nk768 --no-sub --no-funnel > gpurun_out/r2_bench18_zoe.json 2> gpurun_out/r2_bench18_zoe.err; tail -c 1500 gpurun_out/r2_bench18_zoe.json
timeout 900 python bench.py --no-sub --no-funnel > gpurun_out/r2_bench18_beit.json 2> gpurun_out/r2_bench18_beit.err; tail -c 1200 gpurun_out/r2_bench18_beit.json
