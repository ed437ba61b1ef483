#!/bin/bash
# round 2, call AG (8 GPUs): diagnosis of join_groupby strong, peer arm (reps 2+ took 50-70 ms instead of 9)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 tools/debug_join_strong.py --peer 1 > gpurun_out/debug_join_strong.jsonl 2> gpurun_out/debug_join_strong.err; echo "rc=$?"
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/debug_join_strong.err | tail -5
cat gpurun_out/debug_join_strong.jsonl | cut -c1-600
