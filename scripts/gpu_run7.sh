mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_default_b.json 2> gpurun_out/r2_bench_default_b.err; tail -c 600 gpurun_out/r2_bench_default_b.err
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:stereo_row -c 1 -f -o gpurun_out/prof_r2_stereo python tools/profile_step.py stereo2048 4 > gpurun_out/r2_ncu_stereo.log 2>&1; tail -2 gpurun_out/r2_ncu_stereo.log
timeout 900 python -m pytest tests/test_zoe_gpu.py tests/test_funnel_gpu.py tests/test_video_gpu.py -q -s -p no:cacheprovider > gpurun_out/r2_run7_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_run7_tests.log
grep -E "passed|failed|^FAILED|rc=" gpurun_out/r2_run7_tests.log | tail
