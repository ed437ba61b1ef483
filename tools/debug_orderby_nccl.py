"""N > 1 diagnosis of the NCCL-arm ORDER BY (K8b partition + all_to_all): fresh, after a pool trim + empty_cache, and after a
peer arena was created, used and closed in the same process.  torchrun --nproc-per-node N tools/debug_orderby_nccl.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist


def main():
    rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lrank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
    from harkdb_b200.sharded import ShardedEnv, HarkEngine
    from tools.query_suite import Suite
    eng = HarkEngine(lrank)
    env = eng.env

    def run(tag, peer):
        senv = ShardedEnv(eng, peer=peer)
        su = Suite(env, world, rank, senv, 6650.0, 3, 1.0)
        st0 = torch.cuda.memory_stats()
        d = su.orderby("strong")
        st1 = torch.cuda.memory_stats()
        if rank == 0:
            print(json.dumps({"tag": tag, "torch_segments_allocated": st1["segment.all.allocated"] - st0["segment.all.allocated"],
                              "torch_segments_freed": st1["segment.all.freed"] - st0["segment.all.freed"],
                              "torch_reserved_gb": round(st1["reserved_bytes.all.current"] / 2 ** 30, 1), "peer": bool(senv.peer), "ms_all": d["ms_all"], "host_ms_all": d.get("host_ms_all"),
                              "phases": d["phases_ms"], "free_gb": round(torch.cuda.mem_get_info()[0] / 2 ** 30, 1)}), flush=True)
        if senv.peer:
            eng.env.peer_arena_close()
        del senv

    run("nccl_fresh", False)
    run("nccl_again", False)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
