"""N > 1 diagnosis of one sharded query: per-rank, per-repetition device time (CUDA events) and host wall time of
join_groupby (strong split of config 5), fresh and after an ORDER BY exchange, with the phase trace of a late run.
torchrun --nproc-per-node N tools/debug_join_strong.py [--peer 0|1]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--peer", type=int, default=1)
    ap.add_argument("--reps", type=int, default=6)
    args = ap.parse_args()
    rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lrank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
    from harkdb_b200.sharded import ShardedEnv, HarkEngine
    from tools.query_suite import Suite, I32, AGG_SUM, AGG_COUNT, GEN_AFFINE, GEN_UNIFORM, odd_coprime
    eng = HarkEngine(lrank)
    env = eng.env
    senv = ShardedEnv(eng, peer=bool(args.peer))
    su = Suite(env, world, rank, senv, 6650.0, 3, 1.0)

    def gather(x):
        t = torch.tensor([float(x)], device="cuda", dtype=torch.float64)
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [round(float(o.item()), 2) for o in out]

    def join_reps(tag, reps):
        per = 4 * 10 ** 9 // world
        nd = 10 ** 8
        nd_per = nd // world
        a = odd_coprime(2654435761, nd)
        dim = env.synth(nd_per, [I32, I32], [dict(kind=GEN_AFFINE, a=a, b=12345, range=nd), dict(kind=0, lo=0, range=1024)], seed=7, row0=rank * nd_per)
        fact = env.synth(per, [I32, I32], [dict(kind=GEN_UNIFORM, lo=0, range=nd), dict(kind=0, lo=0, range=1000)], seed=42, row0=rank * per)
        sf, sd = su._shard(fact), su._shard(dim)
        run = lambda: senv.join_groupby(sf, sd, 0, 0, 1, [1, 1], [AGG_SUM, AGG_COUNT])
        for _ in range(2):
            run().free()
        rows = []
        r = None
        for i in range(reps):
            traced = i >= reps - 2
            if r is not None:
                r.free()
            senv.trace_on = traced
            senv.pop_trace()
            su.barrier()
            free0 = torch.cuda.mem_get_info()[0]
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            r = run()
            e1.record()
            t_host = (time.perf_counter() - t0) * 1e3
            torch.cuda.synchronize()
            t_sync = (time.perf_counter() - t0) * 1e3
            su.barrier()
            dev = e0.elapsed_time(e1)
            tr = senv.pop_trace() if traced else {}
            senv.trace_on = False
            g_dev, g_host, g_sync = gather(dev), gather(t_host), gather(t_sync)
            if rank == 0:
                print(json.dumps({"tag": tag, "rep": i, "traced": traced, "device_ms": g_dev, "host_enqueue_ms": g_host,
                                  "host_to_sync_ms": g_sync, "free_gb_rank0": round(free0 / 2 ** 30, 1),
                                  "phases_rank0": {k: round(v, 2) for k, v in tr.items()}}), flush=True)
        if r is not None:
            r.free()
        dim.free(); fact.free()
        env.trim(); torch.cuda.empty_cache()

    join_reps("fresh", args.reps)
    d = su.orderby("strong")
    if rank == 0:
        print(json.dumps({"tag": "orderby_strong", "ms_all": d["ms_all"]}), flush=True)
    env.trim(); torch.cuda.empty_cache()
    join_reps("after_orderby", args.reps)
    d = su.join_groupby("weak")
    if rank == 0:
        print(json.dumps({"tag": "join_weak", "ms_all": d["ms_all"]}), flush=True)
    env.trim(); torch.cuda.empty_cache()
    join_reps("after_join_weak", args.reps)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
