"""numpy stand-in for the local operator engine of harkdb_b200.sharded — TEST INFRASTRUCTURE ONLY.

The product engine is `sharded.HarkEngine` (libhark.so on a B200).  This one runs every local operator through the
oracle (oracle/np_oracle.py) on CPU tensors, so that the multi-process host logic of `ShardedEnv` — splitter
selection, count exchange, all-to-all plumbing, partial-aggregate merge, result order — can be tested with the
`gloo` backend on a machine without GPUs."""

import numpy as np
import torch

from oracle import np_oracle as NO

_TORCH = {NO.I32: torch.int32, NO.U32: torch.int32, NO.I64: torch.int64, NO.F32: torch.float32, NO.F64: torch.float64}


class OTable:
    def __init__(self, cols):
        self.cols = [np.ascontiguousarray(c) for c in cols]

    @property
    def dtypes(self):
        return [NO.dtype_code(c) for c in self.cols]

    @property
    def shape(self):
        return (len(self.cols[0]) if self.cols else 0, len(self.cols))

    def columns(self):
        return list(self.cols)

    def free(self):
        pass


def _ordkey64(a, desc):
    u = NO.order_key(a, False).astype(np.uint64)        # widened first, complemented after: libhark's convention
    return ~u if desc else u


class OracleEngine:
    device = torch.device("cpu")

    def to_device(self, arr, dtype=None):
        arr = np.asarray(arr)
        if dtype is not None:
            arr = arr.astype(dtype)
        return OTable([arr[:, c].copy() for c in range(arr.shape[1])])

    def from_columns(self, cols):
        return OTable(cols)

    def columns_torch(self, t):
        out = []
        for c in t.cols:
            v = c.view(np.int32) if c.dtype == np.uint32 else c
            out.append(torch.from_numpy(v.copy()))
        return out

    def from_torch(self, cols, dtypes):
        out = []
        for c, d in zip(cols, dtypes):
            a = c.numpy().copy()
            out.append(a.view(np.uint32) if d == NO.U32 else a)
        return OTable(out)

    def to_numpy_columns(self, t):
        return list(t.cols)

    def sync(self):
        pass

    # ---- operators (oracle) ----
    def query_sel(self, t, cols):
        return OTable([t.cols[int(c)].copy() for c in cols])

    def query_filter(self, t, cols, preds):
        return OTable(NO.query_filter(t.cols, [int(c) for c in cols], list(preds)))

    def query_groupby(self, t, g_col, s_cols, t_cols):
        db = np.stack([c.view(np.uint32) if c.dtype == np.int32 else c for c in t.cols], axis=1).astype(np.uint32)
        out = NO.query_groupby(db, int(g_col), [int(x) for x in s_cols], [int(x) for x in t_cols])
        return OTable([out[:, c].copy() for c in range(out.shape[1])])

    def query_groupby_ex(self, t, g_col, s_cols, ops, having=()):
        return OTable(NO.query_groupby_ex(t.cols, int(g_col), [int(x) for x in s_cols], [int(x) for x in ops], list(having)))

    def with_constant_key(self, t):
        n = len(t.cols[0]) if t.cols else 0
        return OTable([np.zeros(n, dtype=np.int32)] + list(t.cols))

    def query_groupby_multi(self, t, g_cols, s_cols, ops, having=()):
        return OTable(NO.query_groupby_multi(t.cols, [int(g) for g in g_cols], [int(x) for x in s_cols],
                                             [int(x) for x in ops], list(having)))

    def groupby_finalize(self, t, ops):
        out, col = [t.cols[0]], 1
        for op in ops:
            if op == NO.AGG_AVG:
                out.append(t.cols[col] / t.cols[col + 1].astype(np.float64))
                col += 2
            else:
                out.append(t.cols[col])
                col += 1
        assert col == len(t.cols)
        return OTable(out)

    def query_orderby(self, t, cols, key_cols, desc=None):
        return OTable(NO.query_orderby(t.cols, [int(c) for c in cols], [int(k) for k in key_cols], desc))

    def join(self, t1, t2, col1, col2, cols1, cols2):
        d1 = np.stack(t1.cols, axis=1).astype(np.uint32) if t1.cols else np.zeros((0, 0), np.uint32)
        d2 = np.stack(t2.cols, axis=1).astype(np.uint32) if t2.cols else np.zeros((0, 0), np.uint32)
        out = NO.join(d1, d2, int(col1), int(col2), [int(c) for c in cols1], [int(c) for c in cols2])
        return OTable([out[:, c].copy() for c in range(out.shape[1])])

    def join_groupby(self, fact, dim, fk_col, pk_col, g_col, s_cols, ops):
        return OTable(NO.join_groupby(fact.cols, dim.cols, int(fk_col), int(pk_col), int(g_col),
                                      [int(x) for x in s_cols], [int(x) for x in ops]))

    def sample_order_keys(self, t, key_cols, desc, rows):
        rows = np.asarray(rows, dtype=np.int64)
        return np.stack([_ordkey64(t.cols[int(k)][rows], bool(d)) for k, d in zip(key_cols, desc)], axis=1)

    def split_sorted(self, t, key_col, splitters):
        from harkdb_b200.sharded import splitters_to_int64
        k = t.cols[int(key_col)]
        sp = splitters_to_int64(np.asarray(splitters, dtype=np.uint64).reshape(-1), NO.dtype_code(k))
        b = np.searchsorted(k.astype(np.int64), sp, side="left")
        edges = [0] + [int(x) for x in b] + [len(k)]
        return [edges[i + 1] - edges[i] for i in range(len(edges) - 1)]

    def partition_by_splitters(self, t, key_cols, desc, splitters, nparts):
        n = t.shape[0]
        sp = np.asarray(splitters, dtype=np.uint64).reshape(nparts - 1, len(key_cols))
        keys = [_ordkey64(t.cols[int(k)], bool(d)) for k, d in zip(key_cols, desc)]
        digit = np.zeros(n, dtype=np.int64)
        for s in sp:                                   # digit = number of splitters <= key tuple
            le = np.zeros(n, dtype=bool)
            eq = np.ones(n, dtype=bool)
            for j, k in enumerate(keys):
                le |= eq & (s[j] < k)
                eq &= (s[j] == k)
            digit += (le | eq)
        order = np.argsort(digit, kind="stable")
        counts = np.bincount(digit, minlength=nparts).tolist()
        return OTable([c[order] for c in t.cols]), [int(x) for x in counts]
