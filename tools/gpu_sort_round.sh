#!/bin/bash
# One gpurun call for the K3t work: operator parity tests, full-size ORDER BY timing with and without truncation, launch list.
mkdir -p gpurun_out
echo "== pytest ops"; timeout 1500 python -m pytest tests/test_gpu_ops.py -q --timeout=900 -p no:cacheprovider -x > gpurun_out/pytest_ops.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_ops.log
echo "== orderby trunc"; timeout 600 python tools/ops_bench.py --ops orderby --reps 3 --out gpurun_out/orderby_trunc.json > gpurun_out/orderby_trunc.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/orderby_trunc.log
echo "== orderby full"; timeout 600 python tools/ops_bench.py --ops orderby --reps 1 --opt sort.trunc=0 --out gpurun_out/orderby_notrunc.json > gpurun_out/orderby_notrunc.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/orderby_notrunc.log
echo "== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:hk_ -c 200 --csv --log-file gpurun_out/orderby_launches.csv python tools/ops_bench.py --ops orderby --scale 0.25 --reps 1 > gpurun_out/ncu_launch.log 2>&1; echo "rc=$?"
ls -la gpurun_out
