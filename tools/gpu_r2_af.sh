#!/bin/bash
# round 2, call AF: config 5 with the lookup probed in place (one slice = no K8t pass) vs 16 MB slices; 64 MB slices
mkdir -p gpurun_out
for opt in "join.lut_slice_bytes=16777216" "join.lut_slice_bytes=4000000000" "join.lut_slice_bytes=67108864" "join.lut_slice_bytes=33554432"; do
  echo "-- $opt"; timeout 600 python tools/ops_bench.py --ops join --reps 3 --opt $opt 2>&1 | grep '"op"' | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print(d['op'], round(d['total_ms'],2), 'ms', round(d['kernel_ms'],2), d.get('check_ok'), d.get('launches'))"
done | tee gpurun_out/r02_join_slices_ab.txt
