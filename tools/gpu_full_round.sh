#!/bin/bash
# One gpurun call: smoke, the whole GPU parity suite, bench.py, full-size operator timings.
mkdir -p gpurun_out
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
echo "== pytest"; timeout 1800 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider --durations=8 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest.log
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
echo "== ops"; timeout 900 python tools/ops_bench.py --ops groupby,orderby,join --reps 3 --out gpurun_out/ops.json > gpurun_out/ops.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/ops.log
ls -la gpurun_out
