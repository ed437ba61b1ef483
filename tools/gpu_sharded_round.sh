#!/bin/bash
# Multi-GPU round (gpurun --gpus N): configs 3 and 5 at full per-GPU size, config 4 at 0.5e9 rows per GPU; sharded parity tests.
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
echo "== sharded tests"; HARK_PEER=1 timeout 900 $TR -m pytest tests/test_gpu_sharded.py -q -x -p no:cacheprovider > gpurun_out/sharded_tests_n$N.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/sharded_tests_n$N.log
echo "== groupby,join"; HARK_SHARD_TRACE=1 timeout 900 $TR tools/sharded_bench.py --ops groupby,join --reps 3 --out gpurun_out/sharded_n${N}_full.json > gpurun_out/sharded_n${N}_full.log 2>&1; echo "rc=$?"; grep '^{' gpurun_out/sharded_n${N}_full.log | cut -c1-400
echo "== orderby"; HARK_SHARD_TRACE=1 timeout 900 $TR tools/sharded_bench.py --ops orderby --scale 0.25 --reps 3 --out gpurun_out/sharded_n${N}_orderby.json > gpurun_out/sharded_n${N}_orderby.log 2>&1; echo "rc=$?"; grep '^{' gpurun_out/sharded_n${N}_orderby.log | cut -c1-600
echo "== bench"; timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?"; cut -c1-300 gpurun_out/bench_n$N.json
