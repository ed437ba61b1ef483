"""The per-query measurements behind bench.py's `queries` object: BASELINE configs 3, 4 and 5 (GROUP BY, ORDER BY,
JOIN + GROUP BY) plus the variants SURVEY.md §8(d) names, timed through the C-ABI (1 GPU) or through
harkdb_b200.sharded.ShardedEnv (torchrun, N > 1: weak and strong scaling, K8c peer exchange and the NCCL exchange).

Nothing here touches oracle/: results are checked with size-independent properties computed on the device (row counts,
wrap-around sums, pairing hashes, sortedness across rank boundaries) and, at N > 1, by comparing the sharded result of a
reduced-size run bit for bit with the same query on one GPU (`parity_ok`).

Timing of one query: two untimed warm-ups (kernel loading, memory-pool growth), then `reps` runs, each bracketed by a
barrier + torch.cuda.synchronize() on both sides and timed with CUDA events on the stream libhark launches on; the MAX
over ranks of every run is taken, `ms` is the median of the runs.  Inputs are resident and (except the small group /
dimension tables) far larger than the 126 MB L2.
"""
from __future__ import annotations

import gc
import statistics
import time

import numpy as np

I32, U32, I64, F32, F64 = 0, 1, 2, 3, 4
GT = 0
AGG_SUM, AGG_COUNT, AGG_AVG = 2, 5, 6
GEN_UNIFORM, GEN_AFFINE, GEN_AFFINE_UNIFORM = 0, 1, 4
NVLINK_NOMINAL_GBS = 900.0      # per direction per GPU
NVLINK_MEASURED_GBS = 770.0     # B200_PROFILING.md: peer copy per direction

_TS = {0: "<i4", 1: "<i4", 2: "<i8", 3: "<f4", 4: "<f8"}


class _CAI:
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def as_torch(table, col):
    import torch
    n = table.shape[0]
    if n == 0:
        return torch.empty(0, device="cuda", dtype={0: torch.int32, 1: torch.int32, 2: torch.int64, 3: torch.float32,
                                                    4: torch.float64}[table.dtypes[col]])
    return torch.as_tensor(_CAI(table.column_ptr(col), n, _TS[table.dtypes[col]]), device="cuda")


def odd_coprime(a, n):
    while np.gcd(a, n) != 1:
        a += 2
    return a


class Suite:
    """world == 1: `env` (hark_ffi.Futhark) is driven directly.  world > 1: `senv` (ShardedEnv) over the same env."""

    def __init__(self, env, world=1, rank=0, senv=None, peak_gbs=6650.0, reps=3, scale=1.0):
        import torch
        self.torch = torch
        self.env, self.senv = env, senv
        self.world, self.rank = world, rank
        self.peak = peak_gbs
        self.reps = reps
        self.scale = scale
        self.dist = None
        self.last_phases = None
        self.last_host_ms = []
        if world > 1:
            import torch.distributed as dist
            self.dist = dist

    # ---- plumbing ----
    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def allsum(self, x):
        t = self.torch.tensor([int(x)], device="cuda", dtype=self.torch.int64)
        if self.world > 1:
            self.dist.all_reduce(t)
        return int(t.item())

    def wrapsum(self, t):
        t = t.reshape(1).clone()
        if self.world > 1:
            self.dist.all_reduce(t)
        return int(t.item())

    def allmax(self, x):
        t = self.torch.tensor([float(x)], device="cuda", dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, fn, reps=None, warm=2):
        """-> (median ms, all ms, last result, stats of the last run [1 GPU only]).  At N > 1 one extra, traced run
        follows the warm-ups (self.last_phases): no result is alive then, so the memory pool does not grow under it."""
        torch = self.torch
        for _ in range(warm):
            fn().free()
        self.last_phases = self._trace(fn)
        ms_all, r, st = [], None, None
        self.last_host_ms = []   # host time of fn() per run, max over ranks: tells a launch-side stall from a device one
        # Python's cyclic collector stays off inside the timed runs (as timeit does): at 8 GPUs a collection that landed in
        # one rank's run — finalizers of the previous query's tensors and table handles — stalled that rank's launches for
        # ~50 ms and, through the collectives, every rank's device time (profiles/r02_s_bench_n8.json, join strong).
        gc.collect()
        gc_was_on = gc.isenabled()
        gc.disable()
        try:
            for _ in range(reps or self.reps):
                if r is not None:
                    r.free()
                self.barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0 = time.perf_counter()
                e0.record()
                r = fn()
                e1.record()
                host_ms = (time.perf_counter() - t0) * 1e3
                self.barrier()
                ms_all.append(self.allmax(e0.elapsed_time(e1)))
                self.last_host_ms.append(round(self.allmax(host_ms), 3))
                if self.world == 1:
                    st = self.env.stats()
        finally:
            if gc_was_on:
                gc.enable()
        return statistics.median(ms_all), [round(x, 3) for x in ms_all], r, st

    def roofline(self, alg_bytes, ms, st, kernel):
        gbs = alg_bytes / (ms * 1e-3) / 1e9
        d = {"bound": "hbm", "alg_bytes": int(alg_bytes), "achieved": gbs, "peak": self.peak, "unit": "GB/s",
             "frac": gbs / self.peak, "frac_of_nominal_8TBs": gbs / 8000.0, "kernel": kernel, "traffic": None}
        if st is not None:
            d["kernel_ms"] = st["kernel_ms"]
            d["entry_ms"] = st["total_ms"]
            d["launches"] = st["launches"]
        return d

    def rows_for(self, cfg_rows, mode):
        """(rows per rank, total rows).  weak: every rank holds the whole single-GPU configuration; strong: the
        configuration's rows are divided over the ranks."""
        total = int(cfg_rows * self.scale)
        if mode == "weak":
            return total, total * self.world
        per = total // self.world
        return per, per * self.world

    def _trace(self, fn):
        """One extra run with phase tracing on (host wall time per phase, a device sync on both sides of each)."""
        if self.senv is None or self.world == 1:
            return None
        self.senv.trace_on = True
        self.senv.pop_trace()
        self.barrier()
        fn().free()
        self.barrier()
        tr = self.senv.pop_trace()
        self.senv.trace_on = False
        out = {}
        for k, v in tr.items():
            out[k] = round(self.allmax(v), 3)
        return out

    # ---- config 3: GROUP BY key SUM/COUNT/AVG HAVING COUNT > k ----
    def groupby(self, mode="single", f32=False, zipf=False):
        torch, env = self.torch, self.env
        per, total = self.rows_for(10 ** 9, "weak" if mode == "single" else mode)
        vdt = F32 if f32 else I32
        kspec = dict(kind=3 if zipf else 0, lo=0, range=1 << 20)
        vspec = dict(kind=0, flo=0.0, fhi=1.0) if f32 else dict(kind=0, lo=0, range=1000)
        t = env.synth(per, [I32, vdt], [kspec, vspec], seed=42, row0=self.rank * per)
        ops = [AGG_SUM, AGG_COUNT, AGG_AVG]
        sharded = self.world > 1
        tt = self._shard(t) if sharded else t
        run0 = (lambda: self.senv.query_groupby_ex(tt, 0, [1, 1, 1], ops)) if sharded else \
               (lambda: env.query_groupby_ex(t, 0, [1, 1, 1], ops))
        # the HAVING constant: k = median group count of the un-filtered result (config: "HAVING COUNT > k")
        r0 = run0()
        loc = r0.local if sharded else r0
        keys, sums, cnts, avgs = loc.columns()
        n_groups = self.allsum(len(keys))
        if sharded:
            allc = [None] * self.world
            self.dist.all_gather_object(allc, cnts)
            allc = np.concatenate(allc)
        else:
            allc = cnts
        k = int(np.median(allc)) if len(allc) else 0
        expect_having = int((allc > k).sum())
        ok = self.allsum(int(cnts.sum())) == total and bool(np.all(np.diff(keys.astype(np.int64)) > 0))
        v = as_torch(t, 1)
        env.sync()
        if f32:
            tot = self._allsum_f(float(v.sum(dtype=torch.float64).item()))
            got = self._allsum_f(float(sums.astype(np.float64).sum()))
            ok = ok and abs(got - tot) <= 1e-5 * abs(tot) and bool(np.allclose(avgs, sums.astype(np.float64) / cnts, rtol=1e-5))
        else:
            tot = self.wrapsum(v.sum(dtype=torch.int64))
            got = self.allsum(int(sums.view(np.uint32).astype(np.uint64).sum()))
            ok = ok and (got - tot) % (1 << 32) == 0
        ok = ok and self._rank_order_ok(keys)
        r0.free()
        having = [(2, GT, k, 0.0)]
        run = (lambda: self.senv.query_groupby_ex(tt, 0, [1, 1, 1], ops, having)) if sharded else \
              (lambda: env.query_groupby_ex(t, 0, [1, 1, 1], ops, having=having))
        ms, ms_all, r, st = self.timed(run)
        loc = r.local if sharded else r
        n_out = self.allsum(loc.shape[0])
        ok = ok and n_out == expect_having
        alg = total * 8 + n_groups * (4 + 4 + 8 + 8)
        d = {"rows": total, "rows_per_gpu": per, "ms": ms, "ms_all": ms_all, "rows_per_s": total / (ms * 1e-3),
             "groups": n_groups, "having_k": k, "rows_out": n_out, "check_ok": bool(ok),
             "roofline": self.roofline(alg, ms, st, "hk_tpart_kernel + hk_dagg_kernel (K8t + K2)")}
        if sharded:
            d["phases_ms"] = self.last_phases
            d["host_ms_all"] = self.last_host_ms
            d["nvlink_bytes_per_gpu"] = int(n_groups * 24 * (self.world - 1) / self.world)
        r.free()
        t.free()
        return d

    def _allsum_f(self, x):
        t = self.torch.tensor([float(x)], device="cuda", dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t)
        return float(t.item())

    def _shard(self, t):
        from harkdb_b200.sharded import ShardTable
        return ShardTable(self.senv, t)

    def _rank_order_ok(self, keys):
        """rank r's keys all precede rank r+1's (concatenation in rank order = global key order)."""
        if self.world == 1:
            return True
        b = [None] * self.world
        self.dist.all_gather_object(b, (int(keys[0]), int(keys[-1])) if len(keys) else None)
        flat = [x for x in b if x is not None]
        return all(flat[i][1] < flat[i + 1][0] for i in range(len(flat) - 1))

    # ---- config 4: ORDER BY col1, col2 ----
    def orderby(self, mode="single", weak_rows=None):
        torch, env = self.torch, self.env
        cfg = 2 * 10 ** 9 if (mode != "weak" or weak_rows is None) else weak_rows
        per, total = self.rows_for(cfg, "weak" if mode == "single" else mode)
        specs = [dict(kind=0, lo=-(2 ** 19), range=2 ** 20), dict(kind=0, lo=0, range=0)]
        t = env.synth(per, [I64, I64], specs, seed=42, row0=self.rank * per)
        a0, b0 = as_torch(t, 0), as_torch(t, 1)
        env.sync()
        s_in = (self.wrapsum(a0.sum()), self.wrapsum(b0.sum()),
                self.wrapsum(torch.bitwise_xor(a0 * 0x9E3779B97F4A7C15 + b0, b0 >> 7).sum()))
        del a0, b0
        torch.cuda.empty_cache()
        sharded = self.world > 1
        tt = self._shard(t) if sharded else t
        run = (lambda: self.senv.query_orderby(tt, [0, 1], [0, 1], [0, 0])) if sharded else \
              (lambda: env.query_orderby(t, [0, 1], [0, 1]))
        ms, ms_all, r, st = self.timed(run)
        loc = r.local if sharded else r
        a, b = as_torch(loc, 0), as_torch(loc, 1)
        env.sync()
        n_out = self.allsum(a.shape[0])
        s_out = (self.wrapsum(a.sum()), self.wrapsum(b.sum()),
                 self.wrapsum(torch.bitwise_xor(a * 0x9E3779B97F4A7C15 + b, b >> 7).sum()))
        srt = True
        chunk = 1 << 28
        for lo in range(0, a.shape[0] - 1, chunk):
            hi = min(a.shape[0] - 1, lo + chunk)
            okv = (a[lo:hi] < a[lo + 1:hi + 1]) | ((a[lo:hi] == a[lo + 1:hi + 1]) & (b[lo:hi] <= b[lo + 1:hi + 1]))
            srt = srt and bool(okv.all().item())
            del okv
        if sharded:
            ends = [None] * self.world
            mine = ((int(a[0]), int(b[0])), (int(a[-1]), int(b[-1]))) if a.shape[0] else None
            self.dist.all_gather_object(ends, mine)
            flat = [e for e in ends if e is not None]
            srt = srt and all(flat[i][1] <= flat[i + 1][0] for i in range(len(flat) - 1))
        ok = n_out == total and s_in == s_out and self.allsum(int(srt)) == self.world
        sort_info = {k: env.get_option("sort.last_" + k) for k in ("passes", "truncated", "fix_runs", "fallback")}
        npass = max(1, sort_info["passes"])
        d = {"rows": total, "rows_per_gpu": per, "ms": ms, "ms_all": ms_all, "rows_per_s": total / (ms * 1e-3),
             "check_ok": bool(ok), "sort_rank0": sort_info,
             "roofline": self.roofline(2 * total * 16, ms, st, "hk_lsd_scatter_kernel x passes (K3/K3t)")}
        d["roofline"]["pass_level_floor_ms"] = npass * (per * 40) / (self.peak * 1e9) * 1e3
        d["roofline"]["pass_level_note"] = ("a radix sort cannot read once + write once: floor = passes x (8 B histogram read "
                                            "+ 16 B read + 16 B write) per row at the measured peak")
        if sharded:
            d["phases_ms"] = self.last_phases
            d["host_ms_all"] = self.last_host_ms
            d["rank0_load_vs_even"] = a.shape[0] / max(per, 1)
            d["nvlink_bytes_per_gpu"] = int(per * 16 * (self.world - 1) / self.world)
            self._nvlink(d, ("peer_scatter", "exchange"))
        del a, b
        r.free()
        t.free()
        return d

    def _nvlink(self, d, phase_names):
        ph = d.get("phases_ms") or {}
        x = sum(ph.get(k, 0.0) for k in phase_names)
        if x > 0:
            gbs = d["nvlink_bytes_per_gpu"] / (x * 1e-3) / 1e9
            d["nvlink"] = {"exchange_ms": x, "gbs_per_gpu_per_direction": gbs, "frac_of_900": gbs / NVLINK_NOMINAL_GBS,
                           "frac_of_measured_770": gbs / NVLINK_MEASURED_GBS}

    # ---- config 5: fact JOIN dim ON fk = pk GROUP BY attr ----
    def join_groupby(self, mode="single", sparse=False, half=False):
        torch, env = self.torch, self.env
        per, total = self.rows_for(4 * 10 ** 9, "weak" if mode == "single" else mode)
        nd = int(10 ** 8 * min(1.0, self.scale * 4))
        nd_per = nd // self.world
        nd = nd_per * self.world
        if sparse:
            # pk = a*j + b truncated to i32, a large and odd: unique keys spread over the whole 32-bit range
            # (span / rows ~ 43 at 1e8 rows); fk = a*U + b with U uniform over [0, nd)
            a = 2654435761
            dspec = dict(kind=GEN_AFFINE, a=a, b=12345, range=0)
            fspec = dict(kind=GEN_AFFINE_UNIFORM, a=a, b=12345, range=nd * (2 if half else 1))
        else:
            a = odd_coprime(2654435761, nd)
            dspec = dict(kind=GEN_AFFINE, a=a, b=12345, range=nd)
            fspec = dict(kind=GEN_UNIFORM, lo=0, range=nd * (2 if half else 1))
        dim = env.synth(nd_per, [I32, I32], [dspec, dict(kind=0, lo=0, range=1024)], seed=7, row0=self.rank * nd_per)
        fact = env.synth(per, [I32, I32], [fspec, dict(kind=0, lo=0, range=1000)], seed=42, row0=self.rank * per)
        ops = [AGG_SUM, AGG_COUNT]
        sharded = self.world > 1
        if sharded:
            sf, sd = self._shard(fact), self._shard(dim)
            run = lambda: self.senv.join_groupby(sf, sd, 0, 0, 1, [1, 1], ops)
        else:
            run = lambda: env.join_groupby(fact, dim, 0, 0, 1, [1, 1], ops)
        ms, ms_all, r, st = self.timed(run)
        loc = r.local if sharded else r
        keys, sums, cnts = loc.columns()
        n_groups = self.allsum(len(keys))
        val = as_torch(fact, 1)
        env.sync()
        if half:
            # a fact row matches iff its uniform draw fell into the dimension's half of the doubled range
            fk = as_torch(fact, 0)
            if sparse:
                hit = ((fk.to(torch.int64) - 12345) * pow(a, -1, 1 << 32) & 0xFFFFFFFF) < nd
            else:
                hit = fk < nd
            matched = self.allsum(int(hit.sum().item()))
            tot = self.wrapsum(torch.where(hit, val, torch.zeros_like(val)).sum(dtype=torch.int64))
            del hit, fk
        else:
            matched = total
            tot = self.wrapsum(val.sum(dtype=torch.int64))
        got = self.allsum(int(sums.view(np.uint32).astype(np.uint64).sum()))
        ok = n_groups == 1024 and self.allsum(int(cnts.sum())) == matched and (got - tot) % (1 << 32) == 0
        ok = ok and self._rank_order_ok(keys)
        d = {"rows": total, "rows_per_gpu": per, "dim_rows": nd, "ms": ms, "ms_all": ms_all, "rows_per_s": total / (ms * 1e-3),
             "groups": n_groups, "matched_rows": matched, "check_ok": bool(ok),
             "build": "hash (sparse pk)" if sparse else "direct-address lookup (dense pk)",
             "roofline": self.roofline(8 * total + 8 * nd, ms, st, "hk_tpart_kernel + hk_dagg_kernel (probe + aggregate)")}
        if sharded:
            d["phases_ms"] = self.last_phases
            d["host_ms_all"] = self.last_host_ms
            d["nvlink_bytes_per_gpu"] = int(nd_per * 8 * (self.world - 1))
            self._nvlink(d, ("allgather_dim",))
        r.free()
        dim.free()
        fact.free()
        return d

    # ---- the reference's own join entry (join.fut:52): result ordered by key, r1, r2 ----
    def join_entry(self, n1=1 << 27, n2=1 << 24):
        torch, env = self.torch, self.env
        n1, n2 = max(1024, int(n1 * self.scale)), max(256, int(n2 * self.scale))
        a = odd_coprime(2654435761, n2)
        db2 = env.synth(n2, [U32, U32], [dict(kind=GEN_AFFINE, a=a, b=7, range=n2), dict(kind=0, lo=0, range=1024)], seed=7)
        db1 = env.synth(n1, [U32, U32], [dict(kind=0, lo=0, range=n2), dict(kind=0, lo=0, range=1000)], seed=42)
        run = lambda: env.join(db1, db2, 0, 0, [0, 1], [1])
        ms, ms_all, r, st = self.timed(run)
        k, v, at = as_torch(r, 0), as_torch(r, 1), as_torch(r, 2)
        env.sync()
        ok = r.shape[0] == n1 and bool((k[:-1] <= k[1:]).all().item())
        ok = ok and int(v.sum(dtype=torch.int64).item()) == int(as_torch(db1, 1).sum(dtype=torch.int64).item())
        ok = ok and int(k.sum(dtype=torch.int64).item()) == int(as_torch(db1, 0).sum(dtype=torch.int64).item())
        alg = 4 * (n1 + n2) * 2 + r.shape[0] * 12
        d = {"rows": n1 + n2, "rows_out": r.shape[0], "ms": ms, "ms_all": ms_all, "rows_per_s": (n1 + n2) / (ms * 1e-3),
             "check_ok": bool(ok), "roofline": self.roofline(alg, ms, st, "K3 sort of both sides + merge-expand")}
        del k, v, at
        r.free()
        db1.free()
        db2.free()
        return d

    # ---- typed join through the hash build + probe (multiset result, hark_entry_join_ex order = 0) ----
    def join_hash(self, n1=1 << 27, n2=1 << 24, i64=False):
        torch, env = self.torch, self.env
        n1, n2 = max(1024, int(n1 * self.scale)), max(256, int(n2 * self.scale))
        kdt = I64 if i64 else I32
        a = 2654435761 if not i64 else 0x9E3779B97F4A7C15       # odd: j -> a*j + b is injective mod 2^32 / 2^64 (sparse keys)
        db2 = env.synth(n2, [kdt, I32], [dict(kind=GEN_AFFINE, a=a, b=7, range=0), dict(kind=0, lo=0, range=1024)], seed=7)
        db1 = env.synth(n1, [kdt, I32], [dict(kind=GEN_AFFINE_UNIFORM, a=a, b=7, range=n2), dict(kind=0, lo=0, range=1000)], seed=42)
        run = lambda: env.join_ex(db1, db2, 0, 0, [0, 1], [1], 0)
        ms, ms_all, r, st = self.timed(run)
        k, v = as_torch(r, 0), as_torch(r, 1)
        env.sync()
        k0, v0 = as_torch(db1, 0), as_torch(db1, 1)
        ok = r.shape[0] == n1 and bool(torch.equal(k, k0)) and bool(torch.equal(v, v0))    # unique build keys: one match per row, row order kept
        alg = (8 if i64 else 4) * (n1 + n2) + r.shape[0] * ((8 if i64 else 4) + 8) * 2
        d = {"rows": n1 + n2, "rows_out": r.shape[0], "ms": ms, "ms_all": ms_all, "rows_per_s": (n1 + n2) / (ms * 1e-3),
             "key_dtype": "i64" if i64 else "i32", "check_ok": bool(ok),
             "roofline": self.roofline(alg, ms, st, "hk_hj_build_kernel + hk_hj_count_kernel + hk_hj_expand_kernel")}
        del k, v, k0, v0
        r.free()
        db1.free()
        db2.free()
        return d

    # ---- N > 1: sharded result == single-GPU result on a reduced size, bit for bit ----
    def parity(self):
        """Every rank builds the FULL reduced-size tables too and runs the query on its one GPU; the sharded result is
        gathered and compared.  Integer columns bit-exact; f64 AVG / f32 SUM within 1e-12 / 1e-5 relative."""
        env, senv = self.env, self.senv
        out = {}
        W, rank = self.world, self.rank
        n = (1 << 20) * W

        def same(cols_a, cols_b):
            if len(cols_a) != len(cols_b):
                return False
            for x, y in zip(cols_a, cols_b):
                if x.shape != y.shape:
                    return False
                if x.dtype.kind == "f":
                    if not np.allclose(x, y, rtol=1e-12 if x.dtype == np.float64 else 1e-5, atol=0, equal_nan=True):
                        return False
                elif not np.array_equal(x, y):
                    return False
            return True

        def allok(b):
            return self.allsum(int(bool(b))) == W

        # GROUP BY
        # f32 values are non-negative: a relative tolerance says nothing about sums that cancel to ~0
        specs = [dict(kind=0, lo=-5000, range=40000), dict(kind=0, lo=-100, range=1000), dict(kind=0, flo=0.0, fhi=1.0)]
        full = env.synth(n, [I32, I32, F32], specs, seed=11)
        part = env.synth(n // W, [I32, I32, F32], specs, seed=11, row0=rank * (n // W))
        ops, sc = [AGG_SUM, AGG_COUNT, AGG_AVG, 3, 4, AGG_SUM], [1, 1, 1, 1, 2, 2]
        hv = [(2, GT, n // 40000, 0.0)]
        one = env.query_groupby_ex(full, 0, sc, ops, having=hv)
        sh = senv.query_groupby_ex(self._shard(part), 0, sc, ops, hv)
        out["groupby"] = allok(same(one.columns(), senv.gather_columns(sh)))
        one.free(); sh.free()
        # ORDER BY (ties on the first key, DESC second key)
        one = env.query_orderby(full, [0, 1, 2], [0, 1], [0, 1])
        sh = senv.query_orderby(self._shard(part), [0, 1, 2], [0, 1], [0, 1])
        out["orderby"] = allok(same(one.columns(), senv.gather_columns(sh)))
        one.free(); sh.free(); full.free(); part.free()
        # JOIN + GROUP BY (dense and sparse build keys; 50 % of the fact rows match)
        nd = 4096 * W
        for name, dspec, fspec in (
                ("join_groupby", dict(kind=GEN_AFFINE, a=odd_coprime(48271, nd), b=3, range=nd), dict(kind=0, lo=0, range=2 * nd)),
                ("join_groupby_sparse", dict(kind=GEN_AFFINE, a=2654435761, b=3, range=0),
                 dict(kind=GEN_AFFINE_UNIFORM, a=2654435761, b=3, range=2 * nd))):
            dfull = env.synth(nd, [I32, I32], [dspec, dict(kind=0, lo=-7, range=50)], seed=5)
            dpart = env.synth(nd // W, [I32, I32], [dspec, dict(kind=0, lo=-7, range=50)], seed=5, row0=rank * (nd // W))
            ffull = env.synth(n, [I32, I32], [fspec, dict(kind=0, lo=-100, range=1000)], seed=6)
            fpart = env.synth(n // W, [I32, I32], [fspec, dict(kind=0, lo=-100, range=1000)], seed=6, row0=rank * (n // W))
            try:
                one = env.join_groupby(ffull, dfull, 0, 0, 1, [1, 1, 1], [AGG_SUM, AGG_COUNT, AGG_AVG])
                sh = senv.join_groupby(self._shard(fpart), self._shard(dpart), 0, 0, 1, [1, 1, 1], [AGG_SUM, AGG_COUNT, AGG_AVG])
                out[name] = allok(same(one.columns(), senv.gather_columns(sh)))
                one.free(); sh.free()
            except Exception as ex:      # an operator this build refuses is reported, not hidden
                out[name] = "error: " + repr(ex)[:160]
            for x in (dfull, dpart, ffull, fpart):
                x.free()
        return out
