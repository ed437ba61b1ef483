#!/bin/bash
mkdir -p gpurun_out
echo "== groupby/join tests (K8t hot-bin path)"; timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_sanitize_shapes.py -m gpu -q --timeout=600 -p no:cacheprovider -k "groupby or join" > gpurun_out/pytest_gb.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_gb.log | cut -c1-250
echo "== bench"; S=$(date +%s); timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$? wall $(( $(date +%s) - S )) s"; tail -3 gpurun_out/bench_n1.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_n1.json') if l.startswith('{')][-1])
print('N=1 value', d['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'launches', d['gpu_launches'], 'clocks', d['clocks'])
print('e2e', d['e2e']['value'], d['e2e'].get('columnar_host_table',{}).get('value'))
print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cpu_best']['value'], d['cpu_baseline']['cores'])
for k,v in d['queries'].items():
    if 'error' in v: print(k, v['error'][:200]); continue
    print(k, round(v['ms'],2), 'ms', round(v['rows_per_s']/1e9,1),'Grows/s frac', round(v['roofline']['frac'],3), v['check_ok'], {a:round(b['rows_per_s']/1e6) for a,b in (v.get('cpu_baseline') or {}).items() if isinstance(b,dict) and 'rows_per_s' in b})
PY
