#!/bin/bash
# round 2, call AI (4 GPUs): why the NCCL-arm ORDER BY got 2x slower between two N=8 bench runs
mkdir -p gpurun_out
HARK_TRACE_ALLOC=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 tools/debug_orderby_nccl.py > gpurun_out/debug_orderby_nccl.jsonl 2> gpurun_out/debug_orderby_nccl.err; echo "rc=$?"
grep -c "dalloc" gpurun_out/debug_orderby_nccl.err; grep "dalloc" gpurun_out/debug_orderby_nccl.err | grep "device 0" | tail -16 | cut -c1-260
cat gpurun_out/debug_orderby_nccl.jsonl | cut -c1-500
