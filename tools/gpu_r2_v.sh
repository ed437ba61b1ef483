#!/bin/bash
mkdir -p gpurun_out
echo "== sort/join tests"; timeout 1200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_sanitize_shapes.py -m gpu -q --timeout=600 -p no:cacheprovider -k "orderby or sort or sweep16 or join_vs or join_golden or join_ex or join_shapes" > gpurun_out/pytest_sort.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/pytest_sort.log | cut -c1-250
for opt in "sort.straddle=1" "sort.straddle=0"; do
  echo "-- $opt"; timeout 600 python tools/ops_bench.py --ops orderby --reps 3 --opt $opt 2>&1 | grep '"op"' | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print(d['op'], round(d['total_ms'],2), 'ms', round(d['rows_per_s']/1e9,2), 'Grows/s', d.get('check_ok'), d.get('sort'))"
done | tee gpurun_out/r02_orderby_ab4.txt
timeout 600 python tools/ops_bench.py --ops join_entry,join_hash --reps 3 2>&1 | grep '"op"' | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print(d['op'], round(d['ms'],2), 'ms', d.get('check_ok'))"
