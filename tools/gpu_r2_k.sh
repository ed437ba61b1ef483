#!/bin/bash
# round 2, call K: sweep16 parity + ncu of the sweep kernels (0.5e9 rows)
mkdir -p gpurun_out
echo "== sweep tests"; timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout=600 -p no:cacheprovider -k "sweep16" > gpurun_out/pytest_sweep.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest_sweep.log | cut -c1-250
echo "== launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_sweep.csv python tools/ops_bench.py --ops orderby --scale 0.25 --reps 1 > gpurun_out/ncu_sweep_l.log 2>&1; echo "rc=$?"
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/launches_sweep.csv')) if len(r)>10 and r[0].isdigit()]
agg=collections.OrderedDict()
for r in rows:
    name=r[4].split('(')[0][:70]; v=float(r[-1].replace(',',''))
    unit=r[-2]
    if unit=='us' or unit=='usecond': v/=1e3
    elif unit in('ns','nsecond'): v/=1e6
    a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=v
for k,(c,t) in sorted(agg.items(), key=lambda x:-x[1][1])[:14]: print(f"{t:10.3f} ms {c:5d}x {k}")
PY
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:"hk_sweep16_kernel" -s 6 -c 2 -f -o gpurun_out/r02_sweep16 python tools/ops_bench.py --ops orderby --scale 0.25 --reps 1 > gpurun_out/ncu_sweep.log 2>&1; echo "rc=$?"
python tools/ncu_summary.py gpurun_out/r02_sweep16.ncu-rep > gpurun_out/r02_sweep16_ncu.txt 2>&1; cat gpurun_out/r02_sweep16_ncu.txt
