#!/bin/bash
# round 2, call AJ: full GPU suite with the block cache; then quick ops numbers
mkdir -p gpurun_out
echo "== full pytest"; timeout 1800 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest.log | cut -c1-250
timeout 900 python tools/ops_bench.py --ops groupby,groupby_zipf,orderby,join,join_sparse,join_entry,join_hash --reps 3 2>&1 | grep '"op"' | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print(d['op'], round(d.get('total_ms', d.get('ms', 0)),2), 'ms', d.get('check_ok'))"
