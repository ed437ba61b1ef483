#!/bin/bash
# A/B of context options on the full-size operator benchmarks.  Usage: gpu_ops_ab.sh <ops> "opt=v ..." "opt=v ..."
mkdir -p gpurun_out
OPS=$1; shift
i=0
for opts in "$@"; do
  args=""; for o in $opts; do [ "$o" != "-" ] && args="$args --opt $o"; done
  echo "== $OPS [$opts]"; timeout 600 python tools/ops_bench.py --ops $OPS --reps 3 $args > gpurun_out/ops_ab_$i.log 2>&1; echo "rc=$?"
  python - <<PY
import json
for l in open("gpurun_out/ops_ab_$i.log"):
    if l.startswith("{"):
        d=json.loads(l); print(d.get("op"), {k:d.get(k) for k in ("total_ms","kernel_ms","check_ok","error") if d.get(k) is not None})
PY
  i=$((i+1))
done
