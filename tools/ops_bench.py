#!/usr/bin/env python
"""Operator timings for BASELINE configs 3-5 (GROUP BY, ORDER BY, JOIN+GROUP BY) on one GPU, with cheap
on-device result checks (size-independent properties).  Not the driver's bench line (that is bench.py, config 2);
this is the measurement behind DESIGN.md §roofline and profiles/*_ops.json.

  python tools/ops_bench.py [--ops groupby,orderby,join] [--scale 1.0] [--reps 3] [--out gpurun_out/ops.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

I32, U32, I64, F32, F64 = 0, 1, 2, 3, 4
GT = 0
AGG_SUM, AGG_COUNT, AGG_AVG = 2, 5, 6


class _CAI:
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


_TS = {0: "<i4", 1: "<i4", 2: "<i8", 3: "<f4", 4: "<f8"}


def as_torch(table, col):
    import torch
    table._env.sync()    # libhark launches on its own stream; torch must not read the column before it is written
    return torch.as_tensor(_CAI(table.column_ptr(col), table.shape[0], _TS[table.dtypes[col]]), device="cuda")


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def timed(env, fn, reps):
    """One untimed warm-up (grows the context's memory pool, loads the kernels), then `reps` timed runs; every
    result but the last is freed before the next run so that the pool never has to grow inside a timed run."""
    import torch
    fn().free()
    best, r = None, None
    for i in range(reps):
        if r is not None:
            r.free()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = fn()
        env.sync()
        wall = (time.perf_counter() - t0) * 1e3
        st = env.stats()
        st["wall_ms"] = wall
        if best is None or st["total_ms"] < best["total_ms"]:
            best = st
    return best, r


def line(name, n, st, extra=None):
    pk = peak()
    gbs = st["alg_bytes"] / (st["total_ms"] * 1e-3) / 1e9 if st["total_ms"] > 0 else 0.0
    d = {"op": name, "rows": n, "total_ms": st["total_ms"], "kernel_ms": st["kernel_ms"], "wall_ms": st["wall_ms"],
         "alg_bytes": st["alg_bytes"], "rows_out": st["rows_out"], "launches": st["launches"],
         "rows_per_s": n / (st["total_ms"] * 1e-3), "alg_gbs": gbs, "frac_of_measured_peak": gbs / pk, "peak_gbs": pk}
    if extra:
        d.update(extra)
    print(json.dumps(d), flush=True)
    return d


def run_groupby(env, scale, reps):
    import torch
    n = int(10 ** 9 * scale)
    specs = [dict(kind=0, lo=0, range=1 << 20), dict(kind=0, lo=0, range=1000)]
    t = env.synth(n, [I32, I32], specs, seed=42)
    ops = [AGG_SUM, AGG_COUNT, AGG_AVG]
    st, r = timed(env, lambda: env.query_groupby_ex(t, 0, [1, 1, 1], ops), reps)
    keys, sums, cnts, avgs = r.columns()
    ok = (len(keys) == min(1 << 20, len(keys)) and np.all(np.diff(keys.astype(np.int64)) > 0) and int(cnts.sum()) == n)
    total = int(as_torch(t, 1).sum(dtype=torch.int64).item())
    ok = ok and (int(sums.view(np.uint32).astype(np.uint64).sum()) - total) % (1 << 32) == 0
    med = int(np.median(cnts))
    st2, r2 = timed(env, lambda: env.query_groupby_ex(t, 0, [1, 1, 1], ops, having=[(2, GT, med, 0.0)]), 1)
    ok = ok and r2.shape[0] == int((cnts > med).sum())
    # the reference-pinned entry (main.fut:9, u32 SUM) on the same table
    st3, r3 = timed(env, lambda: env.query_groupby(t, 0, [1], [2]), 1)
    ok = ok and r3.shape == (len(keys), 2) and np.array_equal(r3.column(1).view(np.int32), sums)
    r3.free()
    d = line("groupby_cfg3", n, st, {"groups": len(keys), "check_ok": bool(ok), "having_total_ms": st2["total_ms"],
                                     "pinned_query_groupby_sum_ms": st3["total_ms"]})
    r.free(); r2.free(); t.free()
    return d


def run_groupby_f32(env, scale, reps):
    """Config 3's f32 variant: val f32 uniform [0,1); float SUM accumulates in f64 and rounds once."""
    import torch
    n = int(10 ** 9 * scale)
    specs = [dict(kind=0, lo=0, range=1 << 20), dict(kind=0, flo=0.0, fhi=1.0)]
    t = env.synth(n, [I32, F32], specs, seed=42)
    ops = [AGG_SUM, AGG_COUNT, AGG_AVG]
    st, r = timed(env, lambda: env.query_groupby_ex(t, 0, [1, 1, 1], ops), reps)
    keys, sums, cnts, avgs = r.columns()
    total = float(as_torch(t, 1).sum(dtype=torch.float64).item())
    ok = (np.all(np.diff(keys.astype(np.int64)) > 0) and int(cnts.sum()) == n and sums.dtype == np.float32
          and abs(float(sums.astype(np.float64).sum()) - total) <= 1e-5 * total
          and np.allclose(avgs, sums.astype(np.float64) / cnts, rtol=1e-5))
    d = line("groupby_cfg3_f32", n, st, {"groups": len(keys), "check_ok": bool(ok)})
    r.free(); t.free()
    return d


def run_join_half(env, scale, reps):
    """Config 5's 50 %-match run: fk uniform over twice the dimension's key range."""
    import torch
    nf, nd = int(4 * 10 ** 9 * scale), int(10 ** 8 * min(1.0, scale * 4))
    a = 2654435761
    while np.gcd(a, nd) != 1:
        a += 2
    dim = env.synth(nd, [I32, I32], [dict(kind=1, a=a, b=12345, range=nd), dict(kind=0, lo=0, range=1024)], seed=7)
    fact = env.synth(nf, [I32, I32], [dict(kind=0, lo=0, range=2 * nd), dict(kind=0, lo=0, range=1000)], seed=42)
    ops = [AGG_SUM, AGG_COUNT]
    st, r = timed(env, lambda: env.join_groupby(fact, dim, 0, 0, 1, [1, 1], ops), reps)
    keys, sums, cnts = r.columns()
    fk, val = as_torch(fact, 0), as_torch(fact, 1)
    hit = fk < nd
    matched = int(hit.sum().item())
    total = int(torch.where(hit, val, torch.zeros_like(val)).sum(dtype=torch.int64).item())
    del hit
    ok = (len(keys) == 1024 and int(cnts.sum()) == matched
          and (int(sums.view(np.uint32).astype(np.uint64).sum()) - total) % (1 << 32) == 0)
    d = line("join_groupby_cfg5_half_match", nf, st, {"dim_rows": nd, "matched_rows": matched, "check_ok": bool(ok)})
    r.free(); dim.free(); fact.free()
    return d


def run_groupby_zipf(env, scale, reps):
    """Config 3 with skewed keys (HARK_GEN_LOGUNIFORM: P(k) ~ 1/(k+1), key 0 holds 5 % of the rows)."""
    import torch
    n = int(10 ** 9 * scale)
    specs = [dict(kind=3, lo=0, range=1 << 20), dict(kind=0, lo=0, range=1000)]
    t = env.synth(n, [I32, I32], specs, seed=42)
    ops = [AGG_SUM, AGG_COUNT, AGG_AVG]
    st, r = timed(env, lambda: env.query_groupby_ex(t, 0, [1, 1, 1], ops), reps)
    keys, sums, cnts, avgs = r.columns()
    total = int(as_torch(t, 1).sum(dtype=torch.int64).item())
    ok = (bool(np.all(np.diff(keys.astype(np.int64)) > 0)) and int(cnts.sum()) == n
          and (int(sums.view(np.uint32).astype(np.uint64).sum()) - total) % (1 << 32) == 0)
    env.set_option("groupby.impl", 1)
    st_sort, r_sort = timed(env, lambda: env.query_groupby_ex(t, 0, [1, 1, 1], ops), 1)
    env.set_option("groupby.impl", 0)
    ok = ok and all(np.array_equal(a, b) for a, b in zip(r.columns(), r_sort.columns()))   # K2 == sort path, bit for bit
    d = line("groupby_cfg3_zipf", n, st, {"groups": len(keys), "check_ok": bool(ok), "max_group_rows": int(cnts.max()),
                                          "sort_path_total_ms": st_sort["total_ms"]})
    r.free(); r_sort.free(); t.free()
    return d


def run_filter_sweep(env, scale, reps):
    """Config 2 at selectivities 1 / 10 / 25 / 50 / 90 %: col2 > t AND col5 < u, (1 - t) * u = sigma."""
    n = int(10 ** 9 * scale)
    t = env.synth(n, [F32] * 8, [dict(kind=0)] * 8, seed=42)
    out = []
    for sigma in (0.01, 0.10, 0.25, 0.50, 0.90):
        u = sigma ** 0.5
        preds = [(1, GT, 0, 1.0 - u), (4, 2, 0, u)]
        st, r = timed(env, lambda: env.query_filter(t, [0, 2], preds), reps)
        got = r.shape[0] / n
        gbs = st["alg_bytes"] / (st["kernel_ms"] * 1e-3) / 1e9
        out.append({"sigma_target": sigma, "sigma": got, "kernel_ms": st["kernel_ms"], "alg_bytes": st["alg_bytes"],
                    "alg_gbs": gbs, "frac_of_measured_peak": gbs / peak()})
        r.free()
    d = {"op": "filter_sweep_cfg2", "rows": n, "sweep": out, "check_ok": all(abs(x["sigma"] - x["sigma_target"]) < 0.01 for x in out)}
    print(json.dumps(d), flush=True)
    t.free()
    return d


def run_orderby(env, scale, reps):
    import torch
    n = int(2 * 10 ** 9 * scale)
    specs = [dict(kind=0, lo=-(2 ** 19), range=2 ** 20), dict(kind=0, lo=0, range=0)]
    t = env.synth(n, [I64, I64], specs, seed=42)
    a0, b0 = as_torch(t, 0), as_torch(t, 1)
    s1, s2 = int(a0.sum().item()), int(b0.sum().item())
    x2 = int(torch.bitwise_xor(a0 * 0x9E3779B97F4A7C15 + b0, b0 >> 7).sum().item())   # pairing-sensitive multiset hash
    torch.cuda.empty_cache()
    st, r = timed(env, lambda: env.query_orderby(t, [0, 1], [0, 1]), reps)
    a, b = as_torch(r, 0), as_torch(r, 1)
    sums_ok = int(a.sum().item()) == s1 and int(b.sum().item()) == s2
    pair_ok = int(torch.bitwise_xor(a * 0x9E3779B97F4A7C15 + b, b >> 7).sum().item()) == x2
    torch.cuda.empty_cache()
    sorted_ok = True
    chunk = 1 << 28
    for lo in range(0, n - 1, chunk):
        hi = min(n - 1, lo + chunk)
        okv = (a[lo:hi] < a[lo + 1:hi + 1]) | ((a[lo:hi] == a[lo + 1:hi + 1]) & (b[lo:hi] <= b[lo + 1:hi + 1]))
        sorted_ok = sorted_ok and bool(okv.all().item())
        del okv
    sort_info = {k: env.get_option("sort.last_" + k) for k in ("passes", "truncated", "fix_runs", "fallback")}
    d = line("orderby_cfg4", n, st, {"check_ok": bool(sums_ok and pair_ok and sorted_ok), "sums_ok": sums_ok,
                                     "pair_ok": pair_ok, "sorted_ok": sorted_ok, "sort": sort_info})
    del a, b, a0, b0
    r.free(); t.free()
    return d


def run_join(env, scale, reps, sparse=False):
    import torch
    nf, nd = int(4 * 10 ** 9 * scale), int(10 ** 8 * min(1.0, scale * 4))
    # dim: pk = (a*r+b) mod nd with gcd(a, nd)=1 -> a permutation of 0..nd-1 (unique); attr uniform over 1024
    a = 2654435761
    while np.gcd(a, nd) != 1 and not sparse:
        a += 2
    if sparse:      # pk = a*j + b truncated to i32 (spread over the whole 32-bit range), fk = a*U + b, U uniform over [0, nd)
        dim = env.synth(nd, [I32, I32], [dict(kind=1, a=a, b=12345, range=0), dict(kind=0, lo=0, range=1024)], seed=7)
        fact = env.synth(nf, [I32, I32], [dict(kind=4, a=a, b=12345, range=nd), dict(kind=0, lo=0, range=1000)], seed=42)
    else:
        dim = env.synth(nd, [I32, I32], [dict(kind=1, a=a, b=12345, range=nd), dict(kind=0, lo=0, range=1024)], seed=7)
        fact = env.synth(nf, [I32, I32], [dict(kind=0, lo=0, range=nd), dict(kind=0, lo=0, range=1000)], seed=42)
    ops = [AGG_SUM, AGG_COUNT]
    st, r = timed(env, lambda: env.join_groupby(fact, dim, 0, 0, 1, [1, 1], ops), reps)
    keys, sums, cnts = r.columns()
    total = int(as_torch(fact, 1).sum(dtype=torch.int64).item())
    ok = (len(keys) == 1024 and np.array_equal(keys, np.arange(1024, dtype=np.int32)) and int(cnts.sum()) == nf
          and (int(sums.view(np.uint32).astype(np.uint64).sum()) - total) % (1 << 32) == 0)
    d = line("join_groupby_cfg5_sparse_pk" if sparse else "join_groupby_cfg5", nf, st, {"dim_rows": nd, "groups": len(keys), "check_ok": bool(ok)})
    r.free(); dim.free(); fact.free()
    return d


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ops", default="groupby,orderby,join")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--join-scale", type=float, default=None)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--out", default="")
    ap.add_argument("--opt", action="append", default=[], help="key=value context option")
    ap.add_argument("--own-stream", action="store_true",
                    help="let libhark create its own stream; default: torch's current stream (what bench.py and ShardedEnv do), "
                         "so that the torch events tools/query_suite.py times with see the kernels")
    args = ap.parse_args()
    from harkdb_b200 import hark_ffi
    env = hark_ffi.Futhark() if args.own_stream else hark_ffi.Futhark(stream=hark_ffi.torch_stream_handle())
    for kv in args.opt:
        k, v = kv.split("=")
        env.set_option(k, int(v))
    res = []
    for op in args.ops.split(","):
        try:
            if op == "groupby":
                res.append(run_groupby(env, args.scale, args.reps))
            elif op == "orderby":
                res.append(run_orderby(env, args.scale, args.reps))
            elif op == "groupby_f32":
                res.append(run_groupby_f32(env, args.scale, args.reps))
            elif op == "join_half":
                res.append(run_join_half(env, args.join_scale or args.scale, args.reps))
            elif op == "groupby_zipf":
                res.append(run_groupby_zipf(env, args.scale, args.reps))
            elif op == "filter_sweep":
                res.append(run_filter_sweep(env, args.scale, args.reps))
            elif op == "join":
                res.append(run_join(env, args.join_scale or args.scale, args.reps))
            elif op in ("join_entry", "join_hash", "join_hash_i64"):
                from tools.query_suite import Suite
                su = Suite(env, reps=args.reps, scale=args.scale)
                d = su.join_entry() if op == "join_entry" else su.join_hash(i64=op.endswith("i64"))
                d["op"] = op
                print(json.dumps(d), flush=True)
                res.append(d)
            elif op == "join_sparse":
                res.append(run_join(env, args.join_scale or args.scale, args.reps, sparse=True))
        except Exception as e:  # keep going: one OOM must not hide the other operators' numbers
            print(json.dumps({"op": op, "error": repr(e)[:400]}), flush=True)
            res.append({"op": op, "error": repr(e)[:400]})
        try:            # the next operator has other sizes: start it from an empty pool
            env.trim()
            import torch
            torch.cuda.empty_cache()
        except Exception:
            pass
    if args.out:
        json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
