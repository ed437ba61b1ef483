#!/bin/bash
# round 2, call AU: K8t with 3 resident CTAs per SM (40 registers) vs 2
mkdir -p gpurun_out
for opt in "tpart.ctas_per_sm=3" "tpart.ctas_per_sm=2"; do
  echo "-- $opt"; timeout 600 python tools/ops_bench.py --ops groupby,groupby_f32,groupby_zipf,join --reps 5 --opt $opt 2>&1 | grep '"op"' | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print(d['op'], round(d['total_ms'],2), 'ms', round(d['kernel_ms'],2), d.get('check_ok'), d.get('launches'))"
done | tee gpurun_out/r02_tpart_occ_ab.txt
