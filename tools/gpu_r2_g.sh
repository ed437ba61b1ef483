#!/bin/bash
# round 2, call G: the sanitizer shapes as a plain test, then under compute-sanitizer (memcheck, racecheck); SQL golden on the GPU
mkdir -p gpurun_out
echo "== plain"; timeout 600 python -m pytest tests/test_gpu_sanitize_shapes.py tests/test_sql_ext_golden.py tests/test_gpu_sql.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_small.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/pytest_small.log
echo "== memcheck"; timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_sanitize_shapes.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r02_sanitize_memcheck.txt 2>&1; echo "rc=$?"; tail -6 gpurun_out/r02_sanitize_memcheck.txt
echo "== racecheck"; timeout 1800 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 3 python -m pytest tests/test_gpu_sanitize_shapes.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r02_sanitize_racecheck.txt 2>&1; echo "rc=$?"; tail -12 gpurun_out/r02_sanitize_racecheck.txt
