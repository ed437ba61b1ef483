#!/bin/bash
# round 2, call B: K8t tile partition + hash join parity, then the query suite timings
mkdir -p gpurun_out
echo "== new tests"; timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q -x --timeout=600 -p no:cacheprovider -k "tile_partition or hash_build or many_slices" > gpurun_out/pytest_new.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/pytest_new.log
echo "== full pytest"; timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest.log
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_b.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_b.json'))
for k,v in d['queries'].items():
    if 'error' in v: print(k, v['error'][:200]); continue
    print(k, round(v['ms'],2), 'ms', round(v['rows_per_s']/1e9,1),'Grows/s frac', round(v['roofline']['frac'],3), 'kernel_ms', round(v['roofline'].get('kernel_ms',0),2), v['check_ok'])
PY
