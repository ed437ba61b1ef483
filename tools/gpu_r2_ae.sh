#!/bin/bash
# round 2, call AE: K2 over tiles with per-bin ticket dealing (dense.dynamic) vs the static row split
mkdir -p gpurun_out
echo "== groupby tests"; timeout 1200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_sanitize_shapes.py tests/test_gpu_sql.py -m gpu -q --timeout=600 -p no:cacheprovider -k "groupby or dense or agg or sql" > gpurun_out/pytest_gb.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_gb.log | cut -c1-250
for opt in "dense.dynamic=1" "dense.dynamic=0"; do
  echo "-- $opt"; timeout 600 python tools/ops_bench.py --ops groupby_zipf,groupby,groupby_f32 --reps 5 --opt $opt 2>&1 | grep '"op"' | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print(d['op'], round(d['total_ms'],2), 'ms', round(d['kernel_ms'],2), d.get('check_ok'), d.get('launches'))"
done | tee gpurun_out/r02_dynamic_ab.txt
echo "== racecheck groupby"; timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 3 python -m pytest tests/test_gpu_sanitize_shapes.py -m gpu -q -x -p no:cacheprovider -k groupby > gpurun_out/san_race_gb.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/san_race_gb.log | cut -c1-250
