#!/bin/bash
# round 2, call AB: K3b without the trailing barrier (own-row counter zeroing, copies issued by all warps) + barrier-free tie scan
mkdir -p gpurun_out
echo "== sort tests"; timeout 1200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_sanitize_shapes.py -m gpu -q --timeout=600 -p no:cacheprovider -k "orderby or sort or sweep16 or groupby_sort or trunc" > gpurun_out/pytest_sort.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_sort.log | cut -c1-250
for opt in "sort.fix_fast=1" "sort.fix_fast=0"; do
  echo "-- $opt"; timeout 600 python tools/ops_bench.py --ops orderby --reps 3 --opt $opt 2>&1 | grep '"op"' | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print(d['op'], round(d['total_ms'],2), 'ms', round(d['rows_per_s']/1e9,2), 'Grows/s', d.get('check_ok'), d.get('sort'))"
done | tee gpurun_out/r02_orderby_ab6.txt
echo "== memcheck sort"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_sanitize_shapes.py -m gpu -q -x -p no:cacheprovider -k orderby > gpurun_out/san_mem_sort.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/san_mem_sort.log | cut -c1-250
echo "== racecheck sort"; timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 3 python -m pytest tests/test_gpu_sanitize_shapes.py -m gpu -q -x -p no:cacheprovider -k orderby > gpurun_out/san_race_sort.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/san_race_sort.log | cut -c1-250
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
for k,(c,t) in sorted(agg.items(), key=lambda x:-x[1][1])[:6]: print(f"{t:10.3f} ms {c:5d}x {k}")
PY
