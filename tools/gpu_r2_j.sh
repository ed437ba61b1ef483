#!/bin/bash
# round 2, call J: K3b sweep16 + FUSE2 parity, ORDER BY A/B timings
mkdir -p gpurun_out
echo "== orderby tests"; timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_sanitize_shapes.py -m gpu -q -x --timeout=600 -p no:cacheprovider -k "orderby or sort or join or groupby_u32" > gpurun_out/pytest_sort.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_sort.log | cut -c1-250
echo "== A/B"
for opt in "sort.sweep16=1" "sort.sweep16=0 --opt sort.fuse2=1" "sort.sweep16=0 --opt sort.fuse2=0"; do
  echo "-- $opt"; timeout 600 python tools/ops_bench.py --ops orderby --reps 3 --opt $opt 2>&1 | grep '"op"' | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print(d['op'], round(d['total_ms'],2), 'ms', round(d['rows_per_s']/1e9,2), 'Grows/s', d.get('check_ok'), d.get('sort'))"
done | tee gpurun_out/r02_orderby_ab.txt
