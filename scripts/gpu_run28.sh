mkdir -p gpurun_out
timeout 1500 python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_default.json').read().strip().splitlines()[-1])
def show(x, name):
    print(name, 'value', round(x.get('value',0),2), 'ms', round(x.get('ms_per_step',0),2), 'e2e', round(x.get('e2e',{}).get('value',0),2), 'funnel', (x.get('e2e_funnel') or {}).get('value'), 'roof', round((x.get('roofline') or {}).get('frac',0),3), 'cpu', (x.get('cpu_baseline') or {}).get('value'), x.get('clocks'), x.get('error'))
show(d,'main')
for s in d.get('sub_benchmarks',[]): show(s, s.get('workload'))
PY
tail -3 gpurun_out/r02_bench_default.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err; tail -c 600 gpurun_out/r02_bench_reference_arm.json
S="compute-sanitizer --print-limit 10"
timeout 900 $S --tool memcheck python -m pytest tests/test_boost_gpu.py -q -x -k "modelholder_boost_api or (unet and True-1024)" -p no:cacheprovider > gpurun_out/r02_memcheck_boost.log 2>&1
timeout 600 $S --tool memcheck python -m pytest tests/test_zoe_gpu.py -q -x -k "tiny and 3-hw2" -p no:cacheprovider > gpurun_out/r02_memcheck_zoe_v2.log 2>&1
timeout 600 $S --tool racecheck python -m pytest tests/test_zoe_gpu.py -q -x -k "tiny and 3-hw2" -p no:cacheprovider > gpurun_out/r02_racecheck_zoe_v2.log 2>&1
timeout 400 $S --tool memcheck python -m pytest tests/test_vit_ops_gpu.py -q -x -k "test_attention and (257 or 16-16 or 8-6)" -p no:cacheprovider > gpurun_out/r02_memcheck_attention_v2.log 2>&1
for f in gpurun_out/r02_memcheck_boost.log gpurun_out/r02_memcheck_zoe_v2.log gpurun_out/r02_racecheck_zoe_v2.log gpurun_out/r02_memcheck_attention_v2.log; do echo "== $f"; grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY" $f | tail -2; done
