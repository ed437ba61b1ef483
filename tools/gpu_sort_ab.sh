#!/bin/bash
# A/B of the sort scatter variants on config 4 + the operator parity tests.  Usage: gpu_sort_ab.sh "opt1=v opt2=v" ...
mkdir -p gpurun_out
echo "== pytest ops"; timeout 1500 python -m pytest tests/test_gpu_ops.py tests/test_gpu_sharded.py -q --timeout=900 -p no:cacheprovider -x -k "not full_size" > gpurun_out/pytest_ops.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_ops.log
i=0
for opts in "$@"; do
  args=""; for o in $opts; do args="$args --opt $o"; done
  echo "== orderby [$opts]"; timeout 600 python tools/ops_bench.py --ops orderby --reps 3 $args > gpurun_out/orderby_ab_$i.log 2>&1; echo "rc=$?"
  python - <<PY
import json
for l in open("gpurun_out/orderby_ab_$i.log"):
    if l.startswith("{"):
        d=json.loads(l); print({k:d.get(k) for k in ("total_ms","kernel_ms","check_ok","sort","error")})
PY
  i=$((i+1))
done
