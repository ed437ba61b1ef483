#!/bin/bash
# round 2, call E: full parity suite, query timings, per-kernel launch list
mkdir -p gpurun_out
echo "== full pytest"; timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest.log
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_e.json 2> gpurun_out/bench_e.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_e.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_e.json'))
for k,v in d['queries'].items():
    if 'error' in v: print(k, v['error'][:200]); continue
    print(k, round(v['ms'],2), 'ms', round(v['rows_per_s']/1e9,1),'Grows/s frac', round(v['roofline']['frac'],3), 'kernel_ms', round(v['roofline'].get('kernel_ms',0),2), v['check_ok'])
PY
echo "== ncu launch list (groupby, join)"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_ops.csv python tools/ops_bench.py --ops groupby,join --reps 1 > gpurun_out/ncu_ops.log 2>&1; echo "rc=$?"
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/launches_ops.csv')) if len(r)>10 and r[0].isdigit()]
agg=collections.OrderedDict()
for r in rows:
    name=r[4].split('(')[0][:70]; v=float(r[-1].replace(',',''))
    unit=r[-2]
    if unit=='us' or unit=='usecond': v/=1e3
    elif unit in('ns','nsecond'): v/=1e6
    a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=v
for k,(c,t) in sorted(agg.items(), key=lambda x:-x[1][1])[:16]: print(f"{t:10.3f} ms {c:5d}x {k}")
PY
