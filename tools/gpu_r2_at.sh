#!/bin/bash
# round 2, call AT: dealt units with the first bin from the bin share (no row prefix) vs from the row prefix
mkdir -p gpurun_out
echo "== groupby tests"; timeout 1200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_sanitize_shapes.py tests/test_gpu_sql.py -m gpu -q --timeout=600 -p no:cacheprovider -k "groupby or dense or agg or sql or dealt" > gpurun_out/pytest_gb.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_gb.log | cut -c1-250
for opt in "dense.home_by_rows=0" "dense.home_by_rows=1"; do
  echo "-- $opt"; timeout 600 python tools/ops_bench.py --ops groupby_zipf,groupby,groupby_f32 --reps 5 --opt $opt 2>&1 | grep '"op"' | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print(d['op'], round(d['total_ms'],2), 'ms', round(d['kernel_ms'],2), d.get('check_ok'), d.get('launches'))"
done | tee gpurun_out/r02_home_ab.txt
