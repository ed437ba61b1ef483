#!/bin/bash
mkdir -p gpurun_out
echo "== sort tests"; timeout 1200 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout=600 -p no:cacheprovider -k "orderby or sort or sweep16" > gpurun_out/pytest_sort.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_sort.log | cut -c1-250
timeout 600 python tools/ops_bench.py --ops orderby --reps 3 2>&1 | grep '"op"' | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print(d['op'], round(d['total_ms'],2), 'ms', round(d['rows_per_s']/1e9,2), 'Grows/s', d.get('check_ok'), d.get('sort'))"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_sort.csv python tools/ops_bench.py --ops orderby --reps 1 > gpurun_out/ncu_sort_l.log 2>&1; echo "rc=$?"
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/launches_sort.csv')) if len(r)>10 and r[0].isdigit()]
agg=collections.OrderedDict()
for r in rows:
    name=r[4].split('(')[0][:70]; v=float(r[-1].replace(',',''))
    unit=r[-2]
    if unit=='us' or unit=='usecond': v/=1e3
    elif unit in('ns','nsecond'): v/=1e6
    a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=v
for k,(c,t) in sorted(agg.items(), key=lambda x:-x[1][1])[:9]: print(f"{t:10.3f} ms {c:5d}x {k}")
PY
