#!/bin/bash
# round 2, call AQ: new tests (dealt units, block cache, hash-join unique/duplicate, carried columns) + join timings
mkdir -p gpurun_out
echo "== new tests + joins"; timeout 1500 python -m pytest tests/test_gpu_ops.py tests/test_gpu_sanitize_shapes.py tests/test_gpu_sql.py tests/test_sql_ext_golden.py -m gpu -q --timeout=600 -p no:cacheprovider -k "join or dealt or block_cache or carries" > gpurun_out/pytest_join.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/pytest_join.log | cut -c1-250
for opt in "join.carry=1"; do
echo "-- $opt"; timeout 600 python tools/ops_bench.py --ops join_entry --reps 3 --opt $opt 2>/dev/null | grep '"op"' | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print(d['op'], round(d.get('total_ms', d.get('ms', 0)),2), 'ms', d.get('ms_all'), d.get('check_ok'), (d.get('roofline') or {}).get('kernel_ms'))"
done | tee gpurun_out/r02_join_entry_after_ranges.txt
echo "== racecheck joins"; timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 3 python -m pytest tests/test_gpu_sanitize_shapes.py -m gpu -q -x -p no:cacheprovider -k join > gpurun_out/san_mem_join.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/san_mem_join.log | cut -c1-250
