#!/bin/bash
# round 2, call AO: hash join with the unique-key fast path (stop at the first match, one walk in the expand)
mkdir -p gpurun_out
echo "== join tests"; timeout 1200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_sanitize_shapes.py tests/test_gpu_sql.py tests/test_sql_ext_golden.py -m gpu -q --timeout=600 -p no:cacheprovider -k "join" > gpurun_out/pytest_join.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_join.log | cut -c1-250
timeout 600 python tools/ops_bench.py --ops join_entry,join_hash,join_hash_i64 --reps 3 2>/dev/null | grep '"op"' | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print(d['op'], round(d.get('total_ms', d.get('ms', 0)),2), 'ms', d.get('ms_all'), d.get('check_ok'), (d.get('roofline') or {}).get('kernel_ms'))"
echo "== racecheck joins"; timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 3 python -m pytest tests/test_gpu_sanitize_shapes.py -m gpu -q -x -p no:cacheprovider -k join > gpurun_out/san_race_join.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/san_race_join.log | cut -c1-250
