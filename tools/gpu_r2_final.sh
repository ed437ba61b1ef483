#!/bin/bash
# round 2, final verification: full GPU suite, sanitizer, smoke, the driver's bench (both arms), launch list of the bench command
mkdir -p gpurun_out
echo "== full pytest"; timeout 1800 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_final.log | cut -c1-250
echo "== sanitize"; timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_sanitize_shapes.py -m gpu -q -x -p no:cacheprovider > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/sanitize_memcheck.log | cut -c1-200
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 3 python -m pytest tests/test_gpu_sanitize_shapes.py -m gpu -q -x -p no:cacheprovider > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/sanitize_racecheck.log | cut -c1-200
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
echo "== bench (no flags)"; S=$(date +%s); timeout 1500 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$? wall $(( $(date +%s) - S )) s"; tail -2 gpurun_out/bench_n1.err
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
echo "== reference arm"; timeout 900 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?"; cut -c1-300 gpurun_out/bench_ref.json
echo "== launch list of the bench command"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 1 --no-queries --no-cpu > gpurun_out/ncu_final.log 2>&1; echo "rc=$?"
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/final_launches.csv')) if len(r)>10 and r[0].isdigit()]
agg=collections.OrderedDict()
for r in rows:
    name=r[4].split('(')[0][:80]; v=float(r[-1].replace(',',''))
    unit=r[-2]
    if unit=='us' or unit=='usecond': v/=1e3
    elif unit in('ns','nsecond'): v/=1e6
    a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=v
for k,(c,t) in sorted(agg.items(), key=lambda x:-x[1][1])[:10]: print(f"{t:10.3f} ms {c:5d}x {k}")
PY
