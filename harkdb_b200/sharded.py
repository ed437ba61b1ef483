"""Multi-GPU layer: one process per GPU, tables row-range partitioned, `torch.distributed` for the plumbing.

The reference is single-process, single-context (FutharkContext.py:41); this layer is new.  It sits
ABOVE the C-ABI: every rank owns one libhark context (one GPU) and the same SPMD program runs on all ranks.

  filter / projection   shard-local, no collective; the global result is the concatenation in rank order, which
                        is the single-GPU row order.
  GROUP BY              shard-local partial aggregation (K2/K4) -> range repartition of the PARTIAL GROUPS by
                        sampled key splitters (K8b + all-to-all) -> merge (same operators; AVG travels as an f64 sum
                        and an i64 count) -> HAVING.  Rank r ends up with the r-th key range, locally key-ordered, so
                        concatenation in rank order is the single-GPU output order.
  ORDER BY              sampled splitters over the key tuples -> stable K8b partition -> all-to-all -> local stable
                        K3 sort.  Equal key tuples all land on one rank and arrive in (source rank, row) order, so
                        the global result is stable with respect to the global input row order, like one GPU.
  JOIN + GROUP BY       the dimension table is all-gathered (dim << fact), every rank probes and pre-aggregates its
                        own fact shard, partial groups merge as in GROUP BY.
  JOIN                  both sides range-repartitioned by the join key with common splitters, joined locally.

`ShardedEnv` exposes the same methods as `hark_ffi.Futhark`, so `FutharkContext.sql` runs unchanged on top of it
(`ShardedFutharkContext`).  The local operator engine is pluggable: `HarkEngine` (libhark.so, CUDA tensors, NCCL) is
the product; the CPU tests drive the identical host logic over `gloo` with a numpy stand-in engine.
"""

from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

I32, U32, I64, F32, F64 = 0, 1, 2, 3, 4
AGG_KEY, AGG_PROD, AGG_SUM, AGG_MAX, AGG_MIN, AGG_COUNT, AGG_AVG, AGG_SUMF64, AGG_SUM64 = 0, 1, 2, 3, 4, 5, 6, 7, 8
NP_DTYPES = {I32: np.dtype(np.int32), U32: np.dtype(np.uint32), I64: np.dtype(np.int64),
             F32: np.dtype(np.float32), F64: np.dtype(np.float64)}


def _torch_dtype(code):
    import torch
    return {I32: torch.int32, U32: torch.int32, I64: torch.int64, F32: torch.float32, F64: torch.float64}[code]


# ------------------------------------------------------------------------------------------------------------
# host logic shared by every engine
# ------------------------------------------------------------------------------------------------------------
def expand_partial_ops(s_cols: Sequence[int], ops: Sequence[int], pinned_u32: bool = False):
    """Aggregates of the query -> (columns, codes) of the shard-local partial aggregation.
    AVG becomes an f64 sum and a count; every other aggregate is its own partial."""
    p_s, p_ops = [], []
    for c, op in zip(s_cols, ops):
        op = int(op)
        if pinned_u32:
            op = op if op in (AGG_PROD, AGG_SUM, AGG_MAX, AGG_MIN) else AGG_MIN        # groupby.fut:41
        elif op < AGG_PROD or op > AGG_SUM64:
            op = AGG_MIN
        if op == AGG_AVG:
            p_s += [c, c]
            p_ops += [AGG_SUMF64, AGG_COUNT]
        else:
            p_s.append(c)
            p_ops.append(op)
    return p_s, p_ops


def merge_ops_for(p_ops: Sequence[int]) -> List[int]:
    """Operator that combines two partials of each partial column (counts and sums add)."""
    return [AGG_SUM if op in (AGG_SUM, AGG_COUNT, AGG_SUMF64, AGG_SUM64) else op for op in p_ops]


def final_ops_for(ops: Sequence[int], pinned_u32: bool = False) -> List[int]:
    out = []
    for op in ops:
        op = int(op)
        if pinned_u32:
            op = op if op in (AGG_PROD, AGG_SUM, AGG_MAX, AGG_MIN) else AGG_MIN
        elif op < AGG_PROD or op > AGG_SUM64:
            op = AGG_MIN
        out.append(op)
    return out


def sample_positions(n_local: int, nsamples: int) -> np.ndarray:
    """Evenly spaced sample rows of a shard (deterministic)."""
    if n_local <= 0:
        return np.zeros(0, dtype=np.int64)
    s = min(n_local, nsamples)
    return ((np.arange(s, dtype=np.float64) + 0.5) * (n_local / s)).astype(np.int64).clip(0, n_local - 1)


def pick_splitters(samples: Sequence[np.ndarray], weights: Sequence[float], nparts: int, nk: int) -> np.ndarray:
    """Weighted quantiles of the gathered sample tuples -> (nparts-1, nk) uint64 splitters, ascending.
    samples[r] is rank r's (s_r, nk) array of order keys, each row standing for weights[r] table rows."""
    rows = [np.asarray(s, dtype=np.uint64).reshape(-1, nk) for s in samples]
    allk = np.concatenate(rows, axis=0) if rows else np.zeros((0, nk), np.uint64)
    w = np.concatenate([np.full(len(r), float(wt)) for r, wt in zip(rows, weights)]) if rows else np.zeros(0)
    if len(allk) == 0:
        return np.zeros((nparts - 1, nk), dtype=np.uint64)
    order = np.lexsort(tuple(allk[:, j] for j in range(nk - 1, -1, -1)))
    allk, w = allk[order], w[order]
    cw = np.cumsum(w)
    total = cw[-1]
    out = np.empty((nparts - 1, nk), dtype=np.uint64)
    for p in range(1, nparts):
        i = int(np.searchsorted(cw, total * p / nparts, side="left"))
        out[p - 1] = allk[min(i, len(allk) - 1)]
    return out


def splitters_to_int64(sp_ordkeys: np.ndarray, dtype_code: int) -> np.ndarray:
    """Order keys (uint64) of an INTEGER column back to values whose int64 order is the column's order
    (i32: value; u32: value as a non-negative int64; i64: value)."""
    sp = np.asarray(sp_ordkeys, dtype=np.uint64)
    if dtype_code == I32:
        return (sp.astype(np.int64) - (1 << 31)).astype(np.int64)
    if dtype_code == U32:
        return sp.astype(np.int64)
    if dtype_code == I64:
        return (sp ^ np.uint64(1 << 63)).view(np.int64)
    raise ValueError("split_sorted: integer key columns only")


# ------------------------------------------------------------------------------------------------------------
# engines
# ------------------------------------------------------------------------------------------------------------
class _CAI:
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


_TYPESTR = {I32: "<i4", U32: "<i4", I64: "<i8", F32: "<f4", F64: "<f8"}     # u32 travels as its i32 bit pattern


class HarkEngine:
    """The product engine: libhark.so on this rank's GPU; tables are hark_ffi.DeviceTable."""

    def __init__(self, device: int = -1):
        import torch
        from . import hark_ffi
        if device >= 0:
            torch.cuda.set_device(device)
        self.torch = torch
        self.device = torch.device("cuda", torch.cuda.current_device())
        # one stream for torch collectives and libhark kernels: no cross-stream ordering to manage
        self.env = hark_ffi.Futhark(device=self.device.index, stream=hark_ffi.torch_stream_handle())

    # ---- tables ----
    def to_device(self, arr, dtype=None):
        return self.env.to_device(arr, dtype)

    def from_columns(self, cols):
        return self.env.from_columns(cols)

    def columns_torch(self, t):
        n, m = t.shape
        dts = t.dtypes
        out = []
        for c in range(m):
            if n == 0:
                out.append(self.torch.empty(0, dtype=_torch_dtype(dts[c]), device=self.device))
            else:
                out.append(self.torch.as_tensor(_CAI(t.column_ptr(c), n, _TYPESTR[dts[c]]), device=self.device))
        return out

    def from_torch(self, cols, dtypes):
        n = int(cols[0].shape[0]) if cols else 0
        cols = [c.contiguous() for c in cols]
        if n == 0:
            return self.env.from_columns([np.zeros(0, dtype=NP_DTYPES[d]) for d in dtypes])
        t = self.env.from_device_pointers([int(c.data_ptr()) for c in cols], list(dtypes), n)
        t._keepalive = cols          # the table borrows the tensors' memory
        return t

    def empty(self, n, dtype_code):
        return self.torch.empty(n, dtype=_torch_dtype(dtype_code), device=self.device)

    # ---- K8c: partition fused with the exchange (NVLink peer stores into the destination's arena) ----
    def peer_setup(self, senv, arena_bytes: int) -> bool:
        """Creates this rank's receive arena and maps everybody else's.  Collective; False on every rank if any rank
        cannot (no CUDA IPC in this environment, allocation failure): the NCCL path is used instead."""
        from .hark_ffi import HarkError
        ok, handle = 1, b""
        try:
            handle = self.env.peer_arena_create(arena_bytes)
        except HarkError:
            ok = 0
        got = [None] * senv.world
        senv.dist.all_gather_object(got, (ok, handle), group=senv.group)
        if not all(g[0] for g in got):
            self._peer_close_quietly()
            return False
        try:
            self.env.peer_arena_open([g[1] for g in got], senv.rank)
        except HarkError:
            ok = 0
        flags = [None] * senv.world
        senv.dist.all_gather_object(flags, ok, group=senv.group)
        if not all(flags):
            self._peer_close_quietly()
            return False
        self.peer_bytes = int(arena_bytes)
        return True

    def _peer_close_quietly(self):
        try:
            self.env.peer_arena_close()
        except Exception:
            pass

    def peer_repartition(self, senv, local, key_cols, desc, splitters):
        """Returns the rows this rank owns after the exchange (a view of its arena), or None when the rows of some
        destination would not fit its arena (decided identically on every rank from the all-gathered counts)."""
        torch = self.torch
        world = senv.world
        with senv.phase("peer_count"):
            mine = self.env.peer_scatter_count(local, key_cols, desc, splitters, world)
            t = torch.tensor(mine, dtype=torch.int64, device=self.device)
            rows = [torch.empty_like(t) for _ in range(world)]
            senv.dist.all_gather(rows, t, group=senv.group)
            cm = torch.stack(rows).cpu().numpy()                       # [src][dst]
        need = max(self.env.peer_arena_bytes_needed(local, int(cm[:, d].sum())) for d in range(world))
        if need > self.peer_bytes:
            return None
        with senv.phase("peer_scatter"):
            self.env.peer_scatter_run(local, cm)
            # "every rank's stores into my arena have landed": one stream-ordered collective after the scatter
            senv.dist.all_reduce(torch.zeros(1, device=self.device), group=senv.group)
        return self.env.peer_scatter_result()

    def split_sorted(self, t, key_col, splitters):
        """Rows per destination of a table already sorted ascending by the integer column key_col:
        destination of a row = number of splitters <= its key (partition_by_splitters' rule), found by a
        boundary search instead of a partition pass."""
        torch = self.torch
        n = t.shape[0]
        dt = t.dtypes[key_col]
        sp = splitters_to_int64(np.asarray(splitters, dtype=np.uint64).reshape(-1), dt)
        if n == 0:
            return [0] * (len(sp) + 1)
        k = self.columns_torch(t)[key_col]
        if dt == U32:
            k = k.to(torch.int64) & 0xFFFFFFFF
        b = torch.searchsorted(k, torch.from_numpy(sp).to(k.device).to(k.dtype), right=False).cpu().tolist()
        edges = [0] + [int(x) for x in b] + [n]
        return [edges[i + 1] - edges[i] for i in range(len(edges) - 1)]

    def to_numpy_columns(self, t):
        return t.columns()

    def sync(self):
        self.env.sync()

    def __getattr__(self, name):      # every operator entry is the local env's
        return getattr(self.env, name)


class ShardTable:
    """A table whose rows are spread over the ranks; `local` is this rank's shard (an engine table)."""

    def __init__(self, senv: "ShardedEnv", local):
        self._senv = senv
        self.local = local

    @property
    def dtypes(self):
        return self.local.dtypes

    @property
    def shape(self):
        return self.local.shape

    def free(self):
        if self.local is not None:
            self.local.free()
            self.local = None


class ShardedEnv:
    """Same call surface as hark_ffi.Futhark, over row-range shards (see module docstring)."""

    def __init__(self, engine, group=None, oversample: int = 64, trace: Optional[bool] = None,
                 peer: Optional[bool] = None):
        import os
        import torch.distributed as dist
        self.dist = dist
        self.engine = engine
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.oversample = oversample
        # phase tracing (HARK_SHARD_TRACE=1): host wall time per phase with a device sync on both sides
        self.trace_on = bool(int(os.environ.get("HARK_SHARD_TRACE", "0"))) if trace is None else trace
        self.trace = {}
        # partial groups of small dense key domains merge with an all-reduce (HARK_DENSE_MERGE=0: always repartition)
        self.dense_merge = os.environ.get("HARK_DENSE_MERGE", "1") != "0"
        # K8c peer-memory exchange: on when the engine offers it and CUDA IPC works (HARK_PEER=0 forces NCCL;
        # HARK_PEER_ARENA_GB sizes the receive arena, default 24)
        self.peer = False
        want_peer = os.environ.get("HARK_PEER", "1") != "0" if peer is None else bool(peer)
        if self.world > 1 and hasattr(engine, "peer_setup") and want_peer:
            gb = float(os.environ.get("HARK_PEER_ARENA_GB", "24"))
            self.peer = bool(engine.peer_setup(self, int(gb * (1 << 30))))

    class _Phase:
        def __init__(self, senv, name):
            self.senv, self.name = senv, name

        def __enter__(self):
            if self.senv.trace_on:
                import time
                self.senv.engine.sync()
                self.t0 = time.perf_counter()

        def __exit__(self, *exc):
            if self.senv.trace_on:
                import time
                self.senv.engine.sync()
                tr = self.senv.trace
                tr[self.name] = tr.get(self.name, 0.0) + (time.perf_counter() - self.t0) * 1e3
            return False

    def phase(self, name):
        return ShardedEnv._Phase(self, name)

    def pop_trace(self):
        t, self.trace = self.trace, {}
        return t

    # ---- plumbing ----
    def _wrap(self, local) -> ShardTable:
        return ShardTable(self, local)

    def _as_shard(self, db, dtype=None) -> Tuple[ShardTable, bool]:
        if isinstance(db, ShardTable):
            return db, False
        return self.to_device(db, dtype), True

    def to_device(self, arr, dtype=None) -> ShardTable:
        """Every rank passes the SAME host array; rank r keeps rows [r*n/W, (r+1)*n/W)."""
        arr = np.asarray(arr)
        n = arr.shape[0]
        lo, hi = self.rank * n // self.world, (self.rank + 1) * n // self.world
        from .hark_ffi import convert_for_entry
        if dtype is not None:
            arr = convert_for_entry(arr, dtype, "table")     # range check on the whole table, like one GPU
        return self._wrap(self.engine.to_device(np.ascontiguousarray(arr[lo:hi]), None))

    def from_columns(self, cols) -> ShardTable:
        """Every rank passes the SAME host columns (one dtype each); rank r keeps rows [r*n/W, (r+1)*n/W)."""
        cols = [np.asarray(c) for c in cols]
        n = len(cols[0]) if cols else 0
        lo, hi = self.rank * n // self.world, (self.rank + 1) * n // self.world
        return self._wrap(self.engine.from_columns([np.ascontiguousarray(c[lo:hi]) for c in cols]))

    def with_constant_key(self, t: ShardTable) -> ShardTable:
        return self._wrap(self.engine.with_constant_key(t.local))

    def all_counts(self, n_local: int) -> List[int]:
        if self.world == 1:
            return [n_local]
        import torch
        t = torch.tensor([int(n_local)], dtype=torch.int64, device=self.engine.device)
        outs = [torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(outs, t, group=self.group)
        return [int(x) for x in torch.cat(outs).cpu().tolist()]

    PACK_LIMIT_BYTES = 32 << 20      # below this an exchange is latency-bound: all columns travel in ONE all-to-all

    def exchange(self, local, counts: Sequence[int]):
        """Rows of `local` are grouped by destination (counts[d] rows for rank d, in order).  Returns the engine
        table made of what every rank sent here, in source-rank order.  Small tables (partial groups) are packed
        row-major into one byte buffer so that the whole exchange is two collectives (counts + data); large ones
        (ORDER BY shards) go column by column without a packing copy."""
        import torch
        eng, dist = self.engine, self.dist
        dts = local.dtypes
        cols = eng.columns_torch(local)
        dev = cols[0].device if cols else "cpu"
        counts = [int(c) for c in counts]
        send = torch.tensor(counts, dtype=torch.int64, device=dev)
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send, group=self.group)
        recv_l = [int(x) for x in recv.cpu().tolist()]
        n_out = sum(recv_l)
        widths = [c.element_size() for c in cols]
        rowb = sum(widths)
        if len(cols) > 1 and rowb * max(sum(counts), n_out) <= self.PACK_LIMIT_BYTES:
            n_in = sum(counts)
            buf = torch.empty((n_in, rowb), dtype=torch.uint8, device=dev)
            off = 0
            for col, w in zip(cols, widths):
                if n_in:
                    buf[:, off:off + w] = col.reshape(-1).clone().view(torch.uint8).view(n_in, w)
                off += w
            got = torch.empty((n_out, rowb), dtype=torch.uint8, device=dev)
            dist.all_to_all_single(got, buf, recv_l, counts, group=self.group)
            outs, off = [], 0
            for col, w in zip(cols, widths):
                x = torch.empty(n_out * w, dtype=torch.uint8, device=dev)      # a fresh, aligned buffer per column
                x.view(n_out, w).copy_(got[:, off:off + w])
                outs.append(x.view(col.dtype))
                off += w
            return eng.from_torch(outs, dts)
        outs = []
        for col in cols:
            out = torch.empty(n_out, dtype=col.dtype, device=dev)
            dist.all_to_all_single(out, col.contiguous(), recv_l, counts, group=self.group)
            outs.append(out)
        return eng.from_torch(outs, dts)

    def allgather_table(self, local):
        """Every rank gets all shards, concatenated in rank order."""
        import torch
        eng, dist = self.engine, self.dist
        if self.world == 1:
            return local, False
        dts = local.dtypes
        cols = eng.columns_torch(local)
        ns = self.all_counts(local.shape[0])
        mx = max(ns + [1])
        outs = []
        for col in cols:
            pad = torch.zeros(mx, dtype=col.dtype, device=col.device)
            pad[: col.shape[0]] = col
            parts = [torch.empty_like(pad) for _ in range(self.world)]
            dist.all_gather(parts, pad, group=self.group)
            outs.append(torch.cat([p[: ns[r]] for r, p in enumerate(parts)]))
        return eng.from_torch(outs, dts), True

    def _gather_samples(self, mine: np.ndarray, n_local: int, nk: int):
        """All ranks' sample tuples and weights with ONE pair of tensor all-gathers (fixed-size, padded)."""
        import torch
        S = self.oversample * self.world
        if self.world == 1:
            return [mine], [n_local / max(len(mine), 1)]
        buf = np.zeros((S + 1, nk), dtype=np.uint64)
        buf[: len(mine)] = mine
        buf[S, 0] = len(mine)
        dev = self.engine.device
        t = torch.from_numpy(buf.view(np.int64)).to(dev)
        w = torch.tensor([float(n_local)], dtype=torch.float64, device=dev)
        outs = [torch.empty_like(t) for _ in range(self.world)]
        ws = [torch.empty_like(w) for _ in range(self.world)]
        self.dist.all_gather(outs, t, group=self.group)
        self.dist.all_gather(ws, w, group=self.group)
        samples, weights = [], []
        for o, wt in zip(outs, ws):
            a = o.cpu().numpy().view(np.uint64)
            cnt = int(a[S, 0])
            samples.append(a[:cnt].copy())
            weights.append(float(wt.item()) / max(cnt, 1))
        return samples, weights

    def choose_splitters(self, local, key_cols, desc) -> np.ndarray:
        nk = len(key_cols)
        n_local = local.shape[0]
        with self.phase("sample"):
            pos = sample_positions(n_local, self.oversample * self.world)
            mine = self.engine.sample_order_keys(local, key_cols, desc, pos) if len(pos) else np.zeros((0, nk), np.uint64)
            samples, weights = self._gather_samples(mine, n_local, nk)
        return pick_splitters(samples, weights, self.world, nk)

    def repartition(self, local, key_cols, desc, splitters=None, sorted_by_key: bool = False):
        """Range-repartition a shard by key tuple; returns the rows this rank now owns (source-rank order).
        sorted_by_key: the shard is already ascending in its single integer key column (a GROUP BY result), so the
        destinations are contiguous slices and no partition pass is needed."""
        if splitters is None:
            splitters = self.choose_splitters(local, key_cols, desc)
        if sorted_by_key and len(key_cols) == 1 and not desc[0]:
            with self.phase("split_sorted"):
                counts = self.engine.split_sorted(local, key_cols[0], splitters)
            with self.phase("exchange"):
                return self.exchange(local, counts)
        if self.peer:
            got = self.engine.peer_repartition(self, local, key_cols, desc, splitters)
            if got is not None:
                return got
        with self.phase("partition"):
            part, counts = self.engine.partition_by_splitters(local, key_cols, desc, splitters, self.world)
        try:
            with self.phase("exchange"):
                return self.exchange(part, counts)
        finally:
            part.free()

    # ---- result collection ----
    def from_futhark(self, t: ShardTable) -> np.ndarray:
        """Global result as a row-major ndarray on every rank (FutharkContext.py:66,71 `from_futhark`)."""
        full, tmp = self.allgather_table(t.local)
        try:
            cols = self.engine.to_numpy_columns(full)
        finally:
            if tmp:
                full.free()
        dts = t.dtypes
        if not cols:
            return np.empty((0, 0), dtype=np.int32)
        if len(set(dts)) == 1:
            return np.ascontiguousarray(np.stack(cols, axis=1))
        return np.stack([c.astype(np.float64) for c in cols], axis=1)

    def gather_columns(self, t: ShardTable) -> List[np.ndarray]:
        full, tmp = self.allgather_table(t.local)
        try:
            return self.engine.to_numpy_columns(full)
        finally:
            if tmp:
                full.free()

    # ---- shard-local operators ----
    def query_sel(self, db, cols) -> ShardTable:
        t, tmp = self._as_shard(db, np.int32)
        try:
            return self._wrap(self.engine.query_sel(t.local, cols))
        finally:
            if tmp:
                t.free()

    def query_filter(self, db, cols, preds) -> ShardTable:
        t, tmp = self._as_shard(db)
        try:
            return self._wrap(self.engine.query_filter(t.local, cols, preds))
        finally:
            if tmp:
                t.free()

    # ---- GROUP BY ----
    DENSE_MERGE_SLOTS = 1 << 22      # key domains up to this many values merge with a reduce instead of an exchange

    def _merge_dense(self, part, p_ops, pinned_u32):
        """Partial groups of a small, dense key domain are merged with a REDUCE over NVLink instead of a repartition
        (north_star: "partial aggregates merge with an NCCL reduce"): every rank scatters its partial columns into
        dense per-slot arrays (slot = key - global min key), one all-reduce per reduction class (integer sums and
        counts as int64 — a 32-bit SUM wraps exactly like the low word of the 64-bit one —, float sums as f64, integer
        MIN / MAX as int64 with MAX negated), and rank r compacts the non-empty slots of the r-th slice of the domain:
        concatenation in rank order is key order, as after a range repartition.  Returns None when the shape does not
        fit (PROD, float MIN / MAX, a sparse or huge domain): the caller then repartitions."""
        import torch
        eng, dist = self.engine, self.dist
        dts = part.dtypes
        kdt = dts[0]
        if kdt not in (I32, U32, I64):
            return None
        for op, dt in zip(p_ops, dts[1:]):
            if op == AGG_PROD or (op in (AGG_MAX, AGG_MIN) and dt in (F32, F64)):
                return None
        cols = eng.columns_torch(part)
        dev = cols[0].device
        key = cols[0].to(torch.int64)
        if kdt == U32:
            key = key & 0xFFFFFFFF
        big = torch.iinfo(torch.int64).max
        mm = torch.stack([key.min() if key.numel() else torch.tensor(big, device=dev),
                          -key.max() if key.numel() else torch.tensor(big, device=dev)])
        dist.all_reduce(mm, op=dist.ReduceOp.MIN, group=self.group)
        kmin, nkmax = (int(x) for x in mm.cpu().tolist())
        if kmin == big:                     # no group anywhere
            return eng.from_torch([c[:0] for c in cols], dts)
        R = -nkmax - kmin + 1
        if R > self.DENSE_MERGE_SLOTS:
            return None
        slot = key - kmin
        isum, fsum, imin = [], [], []      # (partial column index, source tensor as the reduction dtype)
        for j, (op, dt) in enumerate(zip(p_ops, dts[1:]), start=1):
            c = cols[j]
            if op in (AGG_MAX, AGG_MIN):
                v = c.to(torch.int64)
                if dt == U32:
                    v = v & 0xFFFFFFFF
                imin.append((j, -v if op == AGG_MAX else v))
            elif dt in (F32, F64):
                fsum.append((j, c.to(torch.float64)))
            else:
                isum.append((j, c.to(torch.int64)))
        bi = torch.zeros((len(isum) + 1, R), dtype=torch.int64, device=dev)   # last row: presence
        for r, (_, v) in enumerate(isum):
            bi[r, slot] = v
        bi[len(isum), slot] = 1
        dist.all_reduce(bi, group=self.group)
        bf = bm = None
        if fsum:
            bf = torch.zeros((len(fsum), R), dtype=torch.float64, device=dev)
            for r, (_, v) in enumerate(fsum):
                bf[r, slot] = v
            dist.all_reduce(bf, group=self.group)
        if imin:
            bm = torch.full((len(imin), R), big, dtype=torch.int64, device=dev)
            for r, (_, v) in enumerate(imin):
                bm[r, slot] = v
            dist.all_reduce(bm, op=dist.ReduceOp.MIN, group=self.group)
        s0, s1 = R * self.rank // self.world, R * (self.rank + 1) // self.world
        idx = torch.nonzero(bi[len(isum), s0:s1] > 0).reshape(-1) + s0
        out = [None] * len(dts)
        out[0] = (idx + kmin).to(cols[0].dtype)             # u32 keys above 2^31 wrap back into their i32 bit pattern
        for r, (j, _) in enumerate(isum):
            out[j] = bi[r, idx].to(cols[j].dtype)
        for r, (j, _) in enumerate(fsum):
            out[j] = bf[r, idx].to(cols[j].dtype)
        for r, (j, _) in enumerate(imin):
            v = bm[r, idx]
            out[j] = (-v if p_ops[j - 1] == AGG_MAX else v).to(cols[j].dtype)
        return eng.from_torch(out, dts)

    def _merge_groups(self, part, ops, pinned_u32, having=()):
        """part: shard-local partial groups [key, partials...] (consumed).  Returns this rank's final groups."""
        eng = self.engine
        p_ops = expand_partial_ops(list(range(len(ops))), ops, pinned_u32)[1]
        if self.world > 1:
            merged = None
            if self.dense_merge:
                with self.phase("merge_reduce"):
                    merged = self._merge_dense(part, p_ops, pinned_u32)
            if merged is not None:
                part.free()
            else:
                recv = self.repartition(part, [0], [0], sorted_by_key=True)     # partial groups come out key-ordered
                part.free()
                m = recv.shape[1]
                with self.phase("merge"):
                    if pinned_u32:
                        merged = eng.query_groupby(recv, 0, list(range(1, m)), merge_ops_for(p_ops))
                    else:
                        merged = eng.query_groupby_ex(recv, 0, list(range(1, m)), merge_ops_for(p_ops))
                recv.free()
        else:
            merged = part
        if pinned_u32:
            return merged
        with self.phase("finalize"):
            final = eng.groupby_finalize(merged, final_ops_for(ops))
            merged.free()
            if having:
                f2 = eng.query_filter(final, list(range(final.shape[1])), list(having))
                final.free()
                final = f2
        return final

    def query_groupby(self, db, g_col, s_cols, t_cols) -> ShardTable:
        """main.fut:9 semantics (u32, codes 0-4) over shards."""
        t, tmp = self._as_shard(db, np.uint32)
        try:
            s_cols, t_cols = [int(x) for x in s_cols], [int(x) for x in t_cols]
            if len(t_cols) < len(s_cols):
                from .hark_ffi import HarkError
                raise HarkError(1, "t_cols shorter than s_cols (groupby.fut:47)")
            p_s, p_ops = expand_partial_ops(s_cols, t_cols[: len(s_cols)], pinned_u32=True)
            part = self.engine.query_groupby(t.local, g_col, p_s, p_ops)
            return self._wrap(self._merge_groups(part, t_cols[: len(s_cols)], True))
        finally:
            if tmp:
                t.free()

    def query_groupby_ex(self, db, g_col, s_cols, ops, having=()) -> ShardTable:
        t, tmp = self._as_shard(db)
        try:
            p_s, p_ops = expand_partial_ops([int(x) for x in s_cols], ops)
            with self.phase("local"):
                part = self.engine.query_groupby_ex(t.local, g_col, p_s, p_ops)
            return self._wrap(self._merge_groups(part, [int(x) for x in ops], False, having))
        finally:
            if tmp:
                t.free()

    def query_groupby_multi(self, db, g_cols, s_cols, ops, having=()) -> ShardTable:
        """GROUP BY over several key columns.  Shard-local partial aggregation by the key tuple, range repartition of
        the partial groups by sampled TUPLE splitters (the ORDER BY exchange: equal tuples land on one rank), merge
        with the same multi-key operator, finalize.  Rank order = lexicographic key order."""
        g_cols = [int(g) for g in g_cols]
        if len(g_cols) == 1:
            return self.query_groupby_ex(db, g_cols[0], s_cols, ops, having)
        if len(g_cols) > 4 and self.world > 1:
            raise NotImplementedError("the splitter exchange compares tuples of at most 4 key columns")
        t, tmp = self._as_shard(db)
        try:
            eng, ng = self.engine, len(g_cols)
            ops = [int(x) for x in ops]
            p_s, p_ops = expand_partial_ops([int(x) for x in s_cols], ops)
            with self.phase("local"):
                part = eng.query_groupby_multi(t.local, g_cols, p_s, p_ops)        # [key_1..key_ng, partials...]
            if self.world > 1:
                recv = self.repartition(part, list(range(ng)), [0] * ng)
                part.free()
                m = recv.shape[1]
                with self.phase("merge"):
                    merged = eng.query_groupby_multi(recv, list(range(ng)), list(range(ng, m)), merge_ops_for(p_ops))
                recv.free()
            else:
                merged = part
            with self.phase("finalize"):
                # groupby_finalize passes non-AVG columns through: the keys after the first ride along as such
                final = eng.groupby_finalize(merged, [AGG_MIN] * (ng - 1) + final_ops_for(ops))
                merged.free()
                if having:
                    f2 = eng.query_filter(final, list(range(final.shape[1])), list(having))
                    final.free()
                    final = f2
            return self._wrap(final)
        finally:
            if tmp:
                t.free()

    # ---- ORDER BY ----
    def query_orderby(self, db, cols, key_cols, desc=None) -> ShardTable:
        t, tmp = self._as_shard(db)
        try:
            cols, key_cols = [int(x) for x in cols], [int(x) for x in key_cols]
            desc = [int(x) for x in (desc if desc is not None else [0] * len(key_cols))]
            if self.world == 1 or not key_cols:
                return self._wrap(self.engine.query_orderby(t.local, cols, key_cols, desc))
            need = list(dict.fromkeys(cols + key_cols))          # only these columns cross NVLink
            m = t.shape[1]
            proj = t.local if need == list(range(m)) else self.engine.query_filter(t.local, need, [])
            try:
                k2 = [need.index(k) for k in key_cols]
                recv = self.repartition(proj, k2, desc)
            finally:
                if proj is not t.local:
                    proj.free()
            try:
                with self.phase("local"):
                    return self._wrap(self.engine.query_orderby(recv, [need.index(c) for c in cols], k2, desc))
            finally:
                recv.free()
        finally:
            if tmp:
                t.free()

    # ---- JOIN ----
    def join_groupby(self, fact, dim, fk_col, pk_col, g_col, s_cols, ops) -> ShardTable:
        tf, tmpf = self._as_shard(fact)
        td, tmpd = self._as_shard(dim)
        try:
            with self.phase("allgather_dim"):
                dim_full, dtmp = self.allgather_table(td.local)      # dim << fact: broadcast the build side
            try:
                p_s, p_ops = expand_partial_ops([int(x) for x in s_cols], ops)
                with self.phase("local"):
                    part = self.engine.join_groupby(tf.local, dim_full, fk_col, pk_col, g_col, p_s, p_ops)
            finally:
                if dtmp:
                    dim_full.free()
            return self._wrap(self._merge_groups(part, [int(x) for x in ops], False))
        finally:
            if tmpf:
                tf.free()
            if tmpd:
                td.free()

    def join(self, db1, db2, col1, col2, cols1, cols2) -> ShardTable:
        """join.fut:52 semantics over shards: both sides range-repartitioned by the key with common splitters."""
        t1, tmp1 = self._as_shard(db1, np.uint32)
        t2, tmp2 = self._as_shard(db2, np.uint32)
        try:
            if self.world == 1:
                return self._wrap(self.engine.join(t1.local, t2.local, col1, col2, cols1, cols2))
            for t, c in ((t1, col1), (t2, col2)):
                # join.fut orders by the UNSIGNED key; the splitter exchange compares in the column's own order.  The two
                # agree for u32 columns and for i32 columns without negative values (checked over all ranks).
                if t.dtypes[c] == U32 or (t.dtypes[c] == I32 and self._all_nonnegative(t.local, c)):
                    continue
                from .hark_ffi import HarkError
                raise HarkError(1, "sharded join: key columns must be u32, or i32 without negative values "
                                   "(the reference orders by the unsigned key)")
            # splitters from both sides' keys, so neither side can overload a rank
            s1, w1 = self._samples(t1.local, [col1])
            s2, w2 = self._samples(t2.local, [col2])
            sp = pick_splitters(s1 + s2, w1 + w2, self.world, 1)
            r1 = self.repartition(t1.local, [col1], [0], sp)
            if self.peer:   # r1 may be a view of the receive arena, which the next exchange overwrites
                keep = self.engine.query_filter(r1, list(range(r1.shape[1])), [])
                r1.free()
                r1 = keep
            r2 = self.repartition(t2.local, [col2], [0], sp)
            try:
                return self._wrap(self.engine.join(r1, r2, col1, col2, cols1, cols2))
            finally:
                r1.free()
                r2.free()
        finally:
            if tmp1:
                t1.free()
            if tmp2:
                t2.free()

    def _all_nonnegative(self, local, col) -> bool:
        import torch
        k = self.engine.columns_torch(local)[col]
        mn = torch.tensor([int(k.min().item()) if k.numel() else 0], dtype=torch.int64, device=k.device if k.numel() else self.engine.device)
        if self.world > 1:
            self.dist.all_reduce(mn, op=self.dist.ReduceOp.MIN, group=self.group)
        return int(mn.item()) >= 0

    def _samples(self, local, key_cols):
        n_local = local.shape[0]
        pos = sample_positions(n_local, self.oversample * self.world)
        k = self.engine.sample_order_keys(local, key_cols, [0] * len(key_cols), pos) if len(pos) else \
            np.zeros((0, len(key_cols)), np.uint64)
        return self._gather_samples(k, n_local, len(key_cols))

    def sync(self):
        self.engine.sync()


def ShardedFutharkContext(device: int = -1, group=None, engine=None):
    """FutharkContext (create_table / drop_table / sql) over row-range shards: same class, its FutEnv is a ShardedEnv.
    Every rank calls the same methods with the same arguments; `sql` returns the full result on every rank."""
    from .FutharkContext import FutharkContext
    fc = FutharkContext.__new__(FutharkContext)
    fc.FutEnv = ShardedEnv(engine if engine is not None else HarkEngine(device), group)
    fc.tables = {}
    fc.resident = True
    return fc
