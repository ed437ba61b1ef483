#!/bin/bash
mkdir -p gpurun_out
echo "== microbench: small cp.async.bulk copies"; timeout 300 tools/micro/bulk_small > gpurun_out/r02_micro_bulk_small.txt 2>&1; cat gpurun_out/r02_micro_bulk_small.txt
echo "== spawn tests"; timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q -p no:cacheprovider -k "two_ranks" > gpurun_out/pytest_spawn.log 2>&1; echo "rc=$?"; tail -60 gpurun_out/pytest_spawn.log | cut -c1-300
