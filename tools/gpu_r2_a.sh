#!/bin/bash
# round 2, call A: parity suite + the new bench line (queries object) on one GPU
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/nvsmi.txt 2>&1; nproc >> gpurun_out/nvsmi.txt; free -g >> gpurun_out/nvsmi.txt
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -x --timeout=900 -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest.log
echo "== bench"; timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 6000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
