#!/bin/bash
mkdir -p gpurun_out
echo "== sweep tests"; timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout=600 -p no:cacheprovider -k "sweep16" > gpurun_out/pytest_sweep.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_sweep.log | cut -c1-250
for opt in "sort.sweep16=1" "sort.sweep16=0"; do
  echo "-- $opt"; timeout 600 python tools/ops_bench.py --ops orderby --reps 3 --opt $opt 2>&1 | grep '"op"' | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print(d['op'], round(d['total_ms'],2), 'ms', round(d['rows_per_s']/1e9,2), 'Grows/s', d.get('check_ok'), d.get('sort'))"
done | tee gpurun_out/r02_orderby_ab3.txt
timeout 600 ncu --set full --clock-control none -k regex:"hk_sweep16_kernel" -s 2 -c 1 -f -o gpurun_out/r02_sweep16c python tools/ops_bench.py --ops orderby --scale 0.25 --reps 1 > gpurun_out/ncu_sweep.log 2>&1; echo "rc=$?"
python tools/ncu_summary.py gpurun_out/r02_sweep16c.ncu-rep > gpurun_out/r02_sweep16c_ncu.txt 2>&1; cat gpurun_out/r02_sweep16c_ncu.txt; rm -f gpurun_out/r02_sweep16c.ncu-rep
