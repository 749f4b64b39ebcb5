mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_boost_gpu.py -m gpu -q -s -p no:cacheprovider > gpurun_out/r2_pytest32.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest32.log
grep -E "precision\] (boost|merge)|passed|failed|^FAILED|^ERROR|rc=|Error|capture failed" gpurun_out/r2_pytest32.log | tail -10
timeout 600 python tools/bench_boost_parts.py > gpurun_out/r02_boost_parts_v2.txt 2>&1; grep -E "U-Net|one patch" gpurun_out/r02_boost_parts_v2.txt
timeout 900 python bench.py --workload boost_res101_2048 --no-sub --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/r02_bench_boost_n1_v2.json 2> gpurun_out/r02_bench_boost_n1_v2.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_boost_n1_v2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['roofline']['frac'])
PY
tail -3 gpurun_out/r02_bench_boost_n1_v2.err
