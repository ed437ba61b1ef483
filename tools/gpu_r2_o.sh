#!/bin/bash
mkdir -p gpurun_out
echo "== join tests"; timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_sanitize_shapes.py -m gpu -q --timeout=600 -p no:cacheprovider -k "join" > gpurun_out/pytest_join.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_join.log | cut -c1-250
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_o.json 2> gpurun_out/bench_o.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_o.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_o.json') if l.startswith('{')][-1])
for k,v in d['queries'].items():
    if 'error' in v: print(k, v['error'][:200]); continue
    print(k, round(v['ms'],2), 'ms', round(v['rows_per_s']/1e9,1),'Grows/s frac', round(v['roofline']['frac'],3), 'kernel_ms', round(v['roofline'].get('kernel_ms',0),2), v['check_ok'])
PY
