mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_cabi_gpu.py tests/test_gpu_parity.py tests/test_dav2_gpu.py -q -s -p no:cacheprovider > gpurun_out/r2_run6.log 2>&1; echo "rc=$?" >> gpurun_out/r2_run6.log
grep -E "^\[latency\]|passed|failed|^FAILED|rc=|^E  " gpurun_out/r2_run6.log | tail -30
timeout 600 python bench.py --workload stereo2048 --steps 5 --warmup 3 --no-sub --no-funnel > gpurun_out/r2_bench_stereo_a.json 2> gpurun_out/r2_bench_stereo_a.err; tail -c 1200 gpurun_out/r2_bench_stereo_a.json; tail -3 gpurun_out/r2_bench_stereo_a.err
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_default_a.json 2> gpurun_out/r2_bench_default_a.err; tail -c 4000 gpurun_out/r2_bench_default_a.json; tail -5 gpurun_out/r2_bench_default_a.err
