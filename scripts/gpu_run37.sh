mkdir -p gpurun_out
timeout 1500 python bench.py > gpurun_out/r02_bench_default_final.json 2> gpurun_out/r02_bench_default_final.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_default_final.json').read().strip().splitlines()[-1])
def show(x, name):
    print(name, 'value', round(x.get('value',0),2), 'ms', round(x.get('ms_per_step',0),2), 'e2e', round(x.get('e2e',{}).get('value',0),2), 'funnel', (x.get('e2e_funnel') or {}).get('value'), 'roof', round((x.get('roofline') or {}).get('frac',0),3), 'cpu', (x.get('cpu_baseline') or {}).get('value'), x.get('clocks'), x.get('error'))
show(d,'main')
for s in d.get('sub_benchmarks',[]): show(s, s.get('workload'))
PY
tail -3 gpurun_out/r02_bench_default_final.err
