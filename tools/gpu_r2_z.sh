#!/bin/bash
# round 2, call Z: K3b with the integer digit specialisation (FAST) vs the generic digit; sort tests + sanitizer
mkdir -p gpurun_out
echo "== sort tests"; timeout 1200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_sanitize_shapes.py -m gpu -q --timeout=600 -p no:cacheprovider -k "orderby or sort or sweep16" > gpurun_out/pytest_sort.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_sort.log | cut -c1-250
for opt in "sort.sweep16_fast=1" "sort.sweep16_fast=0" "sort.sweep16_fast=1"; do
  echo "-- $opt"; timeout 600 python tools/ops_bench.py --ops orderby --reps 3 --opt $opt 2>&1 | grep '"op"' | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print(d['op'], round(d['total_ms'],2), 'ms', round(d['rows_per_s']/1e9,2), 'Grows/s', d.get('check_ok'), d.get('sort'))"
done | tee gpurun_out/r02_orderby_ab5.txt
echo "== racecheck sort"; timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 3 python -m pytest tests/test_gpu_sanitize_shapes.py -m gpu -q -x -p no:cacheprovider -k orderby > gpurun_out/san_race_sort.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/san_race_sort.log | cut -c1-250
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"hk_sweep16_kernel" -s 6 -c 1 -f -o gpurun_out/r02_sweep16c python tools/ops_bench.py --ops orderby --scale 0.25 --reps 1 > gpurun_out/ncu_sweepc.log 2>&1; echo "rc=$?"
python tools/ncu_summary.py gpurun_out/r02_sweep16c.ncu-rep > gpurun_out/r02_sweep16_fast_ncu.txt 2>&1; cat gpurun_out/r02_sweep16_fast_ncu.txt
ncu -i gpurun_out/r02_sweep16c.ncu-rep --page source --csv > gpurun_out/sweep16_fast_sass.csv 2>/dev/null; ls -la gpurun_out/sweep16_fast_sass.csv
rm -f gpurun_out/r02_sweep16c.ncu-rep
