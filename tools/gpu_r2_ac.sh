#!/bin/bash
# round 2, call AC: the skewed GROUP BY (Zipf keys, one group = 5 % of the rows) under the partition / accumulator variants
mkdir -p gpurun_out
for opt in "dense.part_impl=0" "dense.part_impl=1" "dense.spec=0" "dense.part_impl=1 --opt dense.spec=0"; do
  echo "-- $opt"; timeout 600 python tools/ops_bench.py --ops groupby_zipf,groupby --reps 3 --opt $opt 2>&1 | grep '"op"' | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print(d['op'], round(d['total_ms'],2), 'ms', round(d['kernel_ms'],2), d.get('check_ok'), d.get('launches'))"
done | tee gpurun_out/r02_zipf_ab.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_zipf.csv python tools/ops_bench.py --ops groupby_zipf --reps 1 > gpurun_out/ncu_zipf_l.log 2>&1; echo "rc=$?"
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/launches_zipf.csv')) if len(r)>10 and r[0].isdigit()]
agg=collections.OrderedDict()
for r in rows:
    name=r[4].split('(')[0][:70]; v=float(r[-1].replace(',',''))
    unit=r[-2]
    if unit=='us' or unit=='usecond': v/=1e3
    elif unit in('ns','nsecond'): v/=1e6
    a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=v
for k,(c,t) in sorted(agg.items(), key=lambda x:-x[1][1])[:8]: print(f"{t:10.3f} ms {c:5d}x {k}")
PY
