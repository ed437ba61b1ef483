#!/usr/bin/env python
"""Where the sharded ORDER BY spends its time: exchange and local sort timed separately (CUDA events, per rank),
with the local sort's own entry statistics and pass counters.  torchrun --nproc-per-node N tools/sharded_orderby_diag.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    from harkdb_b200.sharded import HarkEngine, ShardedEnv
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
    eng = HarkEngine(int(os.environ.get("LOCAL_RANK", "0")))
    senv = ShardedEnv(eng)
    env = eng.env
    per = int(float(sys.argv[1])) if len(sys.argv) > 1 else 500_000_000
    specs = [dict(kind=0, lo=-(2 ** 19), range=2 ** 20), dict(kind=0, lo=0, range=0)]
    t = env.synth(per, [2, 2], specs, seed=42, row0=rank * per)
    out = []
    for rep in range(3):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev[0].record()
        recv = senv.repartition(t, [0, 1], [0, 0]) if world > 1 else t
        ev[1].record()
        r = eng.query_orderby(recv, [0, 1], [0, 1], [0, 0])
        ev[2].record()
        torch.cuda.synchronize()
        st = env.stats()
        out.append({"rows_local": r.shape[0], "exchange_ms": round(ev[0].elapsed_time(ev[1]), 2),
                    "local_sort_ms": round(ev[1].elapsed_time(ev[2]), 2), "entry_total_ms": round(st["total_ms"], 2),
                    "entry_kernel_ms": round(st["kernel_ms"], 2),
                    "sort": {k: env.get_option("sort.last_" + k) for k in ("passes", "truncated", "fix_runs", "fallback")}})
        r.free()
        if world > 1:
            recv.free()
    print(json.dumps({"rank": rank, "reps": out}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
