#!/bin/bash
mkdir -p gpurun_out
echo "== join tests"; timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_sanitize_shapes.py -m gpu -q --timeout=600 -p no:cacheprovider -k "join" > gpurun_out/pytest_join.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_join.log | cut -c1-250
for op in join join_sparse; do timeout 600 python tools/ops_bench.py --ops $op --reps 3 2>&1 | grep '"op"' | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print(d['op'], round(d['total_ms'],2), 'ms', round(d['rows_per_s']/1e9,2), 'Grows/s', d.get('check_ok'))"; done
timeout 900 ncu --set full --clock-control none -k regex:"hk_dagg_tiles_kernel|hk_hash_build" -s 0 -c 2 -f -o gpurun_out/r02_k2_hash2 python tools/ops_bench.py --ops join_sparse --reps 1 > gpurun_out/r02_k2_hash2.log 2>&1
python tools/ncu_summary.py gpurun_out/r02_k2_hash2.ncu-rep > gpurun_out/r02_k2_hash2_ncu.txt 2>&1; rm -f gpurun_out/r02_k2_hash2.ncu-rep
grep -E "^==|time_duration|dram__bytes_read|hit_rate|inst_executed.sum" gpurun_out/r02_k2_hash2_ncu.txt | cut -c1-150
