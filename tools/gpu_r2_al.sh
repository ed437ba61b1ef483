#!/bin/bash
# round 2, call AL: joins on libhark's own stream vs on torch's (legacy default) stream
mkdir -p gpurun_out
for flag in "" "--torch-stream"; do
  echo "-- flag: $flag"; HARK_TRACE_ALLOC=1 timeout 600 python tools/ops_bench.py --ops join_entry,join_hash,groupby $flag --reps 3 2> gpurun_out/al.err | grep '"op"' | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print(d['op'], round(d.get('total_ms', d.get('ms', 0)),2), 'ms', d.get('ms_all'), d.get('check_ok'), (d.get('roofline') or {}).get('kernel_ms'), (d.get('roofline') or {}).get('entry_ms'))"
  grep -c dalloc gpurun_out/al.err; grep dalloc gpurun_out/al.err | tail -4 | cut -c1-200
done
