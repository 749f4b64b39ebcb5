mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_beit_gpu.py tests/test_model_cabi_gpu.py -q -s -p no:cacheprovider > gpurun_out/r2_run8_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_run8_tests.log
grep -E "^\[precision\]|^\[latency\]|passed|failed|^FAILED|rc=|^E  " gpurun_out/r2_run8_tests.log | tail -30
timeout 900 python bench.py --workload zoedepth_nk768 --steps 3 --warmup 3 --no-sub --no-funnel > gpurun_out/r2_bench_zoe_a.json 2> gpurun_out/r2_bench_zoe_a.err; tail -c 2500 gpurun_out/r2_bench_zoe_a.json; tail -5 gpurun_out/r2_bench_zoe_a.err
