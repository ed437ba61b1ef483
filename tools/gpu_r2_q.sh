#!/bin/bash
# round 2, call Q: what the driver runs at round end on one GPU: smoke, pytest -m gpu, bench (both arms)
mkdir -p gpurun_out
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -9 gpurun_out/smoke.log
echo "== pytest"; timeout 1800 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest.log | cut -c1-200
echo "== bench ref"; timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2> /dev/null; echo "rc=$?"; cut -c1-200 gpurun_out/bench_ref.json
echo "== bench"; /usr/bin/time -v timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; grep -E "Elapsed|Maximum resident" gpurun_out/bench_n1.err
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
