#!/bin/bash
# round 2, call AK: the driver's N=1 bench (default flags) + the reference arm, final state
mkdir -p gpurun_out
echo "== bench (no flags)"; S=$(date +%s); timeout 1500 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$? wall $(( $(date +%s) - S )) s"; tail -3 gpurun_out/bench_n1.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_n1.json') if l.startswith('{')][-1])
print('N=1 value', d['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'traffic', d['roofline']['traffic'], 'launches', d['gpu_launches'], 'clocks', d['clocks'])
print('e2e', d['e2e']['value'], d['e2e'].get('columnar_host_table',{}).get('value'))
print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cpu_best']['value'], d['cpu_baseline']['cores'])
for k,v in d['queries'].items():
    if not isinstance(v, dict): continue
    if 'error' in v: print(k, v['error'][:200]); continue
    print(k, round(v['ms'],2), 'ms', v.get('ms_all'), round(v['rows_per_s']/1e9,1),'Grows/s frac', round(v['roofline']['frac'],3), v['check_ok'], {a:round(b['rows_per_s']/1e6) for a,b in (v.get('cpu_baseline') or {}).items() if isinstance(b,dict) and 'rows_per_s' in b})
PY
echo "== reference arm"; S=$(date +%s); timeout 900 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$? wall $(( $(date +%s) - S )) s"; cut -c1-600 gpurun_out/bench_ref.json
