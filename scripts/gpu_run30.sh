mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_boost_gpu.py tests/test_leres_gpu.py -m gpu -q -s -p no:cacheprovider > gpurun_out/r2_pytest30.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest30.log
grep -E "precision\] boost|passed|failed|^FAILED|^ERROR|rc=|Error|capture failed" gpurun_out/r2_pytest30.log | tail -10
timeout 900 python bench.py --workload boost_res101_2048 --no-sub --steps 5 --warmup 3 > gpurun_out/r02_bench_boost_n1.json 2> gpurun_out/r02_bench_boost_n1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_boost_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['gpu_launches'], d['clocks'], d['roofline']['frac'], d['cpu_baseline']['value'])
PY
tail -3 gpurun_out/r02_bench_boost_n1.err
