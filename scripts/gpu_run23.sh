mkdir -p gpurun_out
timeout 900 python bench.py --workload boost_res101_2048 --no-sub --steps 3 --warmup 3 > gpurun_out/r2_bench23_boost.json 2> gpurun_out/r2_bench23_boost.err; tail -c 2500 gpurun_out/r2_bench23_boost.json; tail -5 gpurun_out/r2_bench23_boost.err
timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_pytest23_full.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest23_full.log
grep -E "passed|failed|^FAILED|^ERROR|rc=" gpurun_out/r2_pytest23_full.log | tail -12
