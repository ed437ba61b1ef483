#!/bin/bash
# One gpurun call: smoke, GPU parity tests, short bench, optional ncu.  Usage: run_gpu_round.sh [quick|full]
MODE=${1:-full}
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvsmi.txt 2>&1; nproc >> gpurun_out/nvsmi.txt; free -g >> gpurun_out/nvsmi.txt
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -8 gpurun_out/smoke.log
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider ${PYTEST_ARGS} > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest.log
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 ${BENCH_ARGS} > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
if [ "$MODE" = "full" ]; then
echo "== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launch.log 2>&1; echo "rc=$?"
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:hk_filter -s 3 -c 2 -f -o gpurun_out/prof_filter python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"
fi
ls -la gpurun_out
