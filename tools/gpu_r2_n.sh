#!/bin/bash
# round 2, call N: partitioned hash build parity + sparse join launch list; full GPU suite
mkdir -p gpurun_out
echo "== full pytest"; timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest.log | cut -c1-250
echo "== sparse join A/B"
for opt in "join.build_partition=1" "join.build_partition=0"; do
  echo "-- $opt"; timeout 600 python tools/ops_bench.py --ops join_sparse --reps 3 --opt $opt 2>&1 | grep '"op"' | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print(d['op'], round(d['total_ms'],2), 'ms', round(d['rows_per_s']/1e9,2), 'Grows/s', d.get('check_ok'))"
done | tee gpurun_out/r02_join_sparse_ab.txt
echo "== launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_sparse.csv python tools/ops_bench.py --ops join_sparse --reps 1 > gpurun_out/ncu_sparse_l.log 2>&1; echo "rc=$?"
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/launches_sparse.csv')) if len(r)>10 and r[0].isdigit()]
agg=collections.OrderedDict()
for r in rows:
    name=r[4].split('(')[0][:70]; v=float(r[-1].replace(',',''))
    unit=r[-2]
    if unit=='us' or unit=='usecond': v/=1e3
    elif unit in('ns','nsecond'): v/=1e6
    a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=v
for k,(c,t) in sorted(agg.items(), key=lambda x:-x[1][1])[:10]: print(f"{t:10.3f} ms {c:5d}x {k}")
PY
