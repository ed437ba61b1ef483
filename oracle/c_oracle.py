"""ctypes loader for oracle/liboracle.so (oracle.c) — TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this.  Build with ``make -C oracle`` (``__graft_entry__.build()`` does).
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class HarkPred(C.Structure):
    _fields_ = [("col", C.c_int32), ("op", C.c_int32), ("ival", C.c_int64), ("fval", C.c_double)]


class HarkColspec(C.Structure):
    _fields_ = [("kind", C.c_int32), ("reserved", C.c_int32), ("lo", C.c_int64), ("range", C.c_uint64),
                ("flo", C.c_double), ("fhi", C.c_double), ("a", C.c_uint64), ("b", C.c_uint64)]


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", _HERE, "-s"])
        _LIB = C.CDLL(path)
        _LIB.oracle_segmented_reduce_add_i32.restype = C.c_int64
        _LIB.oracle_replicated_iota.restype = C.c_int64
        _LIB.oracle_expand_mul.restype = C.c_int64
        _LIB.oracle_expand_reduce_mul_add.restype = C.c_int64
        _LIB.oracle_expand_outer_reduce_mul_add.restype = C.c_int64
        _LIB.oracle_mix64.restype = C.c_uint64
        _LIB.oracle_mix64.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64]
    return _LIB


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _i32(xs) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(xs, dtype=np.int32).reshape(-1))


def segmented_scan_add(flags, xs) -> np.ndarray:
    f = np.ascontiguousarray(np.asarray(flags, dtype=np.uint8).reshape(-1))
    a = _i32(xs)
    out = np.empty(len(a), dtype=np.int32)
    lib().oracle_segmented_scan_add_i32(_p(f), _p(a), C.c_int64(len(a)), _p(out))
    return out


def segmented_reduce_add(flags, xs) -> np.ndarray:
    f = np.ascontiguousarray(np.asarray(flags, dtype=np.uint8).reshape(-1))
    a = _i32(xs)
    out = np.empty(max(len(a), 1), dtype=np.int32)
    g = lib().oracle_segmented_reduce_add_i32(_p(f), _p(a), C.c_int64(len(a)), _p(out))
    return out[:g].copy()


def replicated_iota(reps) -> np.ndarray:
    r = _i32(reps)
    out = np.empty(max(int(r.sum()) if len(r) else 0, 1), dtype=np.int32)
    t = lib().oracle_replicated_iota(_p(r), C.c_int64(len(r)), _p(out))
    return out[:t].copy()


def segmented_iota(flags) -> np.ndarray:
    f = np.ascontiguousarray(np.asarray(flags, dtype=np.uint8).reshape(-1))
    out = np.empty(len(f), dtype=np.int32)
    lib().oracle_segmented_iota(_p(f), C.c_int64(len(f)), _p(out))
    return out


def _expand_like(fn, arr, cap) -> np.ndarray:
    a = _i32(arr)
    out = np.empty(max(cap, 1), dtype=np.int32)
    t = fn(_p(a), C.c_int64(len(a)), _p(out))
    return out[:t].copy()


def expand_mul(arr) -> np.ndarray:
    return _expand_like(lib().oracle_expand_mul, arr, int(np.sum(arr)) if len(arr) else 0)


def expand_reduce_mul_add(arr) -> np.ndarray:
    return _expand_like(lib().oracle_expand_reduce_mul_add, arr, len(arr))


def expand_outer_reduce_mul_add(arr) -> np.ndarray:
    return _expand_like(lib().oracle_expand_outer_reduce_mul_add, arr, len(arr))


def query_sel(db: np.ndarray, cols: Sequence[int]) -> np.ndarray:
    db = np.ascontiguousarray(db, dtype=np.int32)
    n, m = db.shape
    c = _i32(cols)
    out = np.empty((n, len(c)), dtype=np.int32)
    rc = lib().oracle_query_sel_i32(_p(db), C.c_int64(n), C.c_int64(m), _p(c), C.c_int64(len(c)), _p(out))
    if rc:
        raise IndexError("column index out of bounds")
    return out


def query_groupby(db: np.ndarray, g_col: int, s_cols: Sequence[int], t_cols: Sequence[int]) -> np.ndarray:
    db = np.ascontiguousarray(db, dtype=np.uint32)
    n, m = db.shape
    s, t = _i32(s_cols), _i32(t_cols)
    c = len(s)
    if len(t) < c:
        raise IndexError("t_cols shorter than s_cols")
    out = np.empty((max(n, 1), c + 1), dtype=np.uint32)
    G = C.c_int64(0)
    rc = lib().oracle_query_groupby_u32(_p(db), C.c_int64(n), C.c_int64(m), C.c_int32(g_col), _p(s), _p(t),
                                        C.c_int64(c), _p(out), C.byref(G))
    if rc:
        raise IndexError("column index out of bounds")
    return out[:G.value].copy()


def join(db1: np.ndarray, db2: np.ndarray, col1: int, col2: int, cols1: Sequence[int],
         cols2: Sequence[int]) -> np.ndarray:
    db1 = np.ascontiguousarray(db1, dtype=np.uint32)
    db2 = np.ascontiguousarray(db2, dtype=np.uint32)
    c1, c2 = _i32(cols1), _i32(cols2)
    outp = C.POINTER(C.c_uint32)()
    P = C.c_int64(0)
    rc = lib().oracle_join_u32(_p(db1), C.c_int64(db1.shape[0]), C.c_int64(db1.shape[1]),
                               _p(db2), C.c_int64(db2.shape[0]), C.c_int64(db2.shape[1]),
                               C.c_int32(col1), C.c_int32(col2), _p(c1), C.c_int64(len(c1)), _p(c2),
                               C.c_int64(len(c2)), C.byref(outp), C.byref(P))
    if rc:
        raise IndexError("column index out of bounds")
    w = len(c1) + len(c2)
    res = np.ctypeslib.as_array(outp, shape=(P.value * w,)).copy().reshape(P.value, w) if P.value * w else \
        np.zeros((P.value, w), dtype=np.uint32)
    lib().oracle_free(outp)
    return res


def make_preds(preds: Sequence[Tuple[int, int, int, float]]):
    arr = (HarkPred * max(len(preds), 1))()
    for i, (c, op, iv, fv) in enumerate(preds):
        arr[i] = HarkPred(int(c), int(op), int(iv), float(fv))
    return arr


def query_filter(db: np.ndarray, cols: Sequence[int], preds: Sequence[Tuple[int, int, int, float]],
                 threads: int = 1, out: np.ndarray | None = None) -> np.ndarray:
    """Row-major table in, row-major [n_out][k] out (oracle_query_filter)."""
    from .np_oracle import DTYPE_CODES
    db = np.ascontiguousarray(db)
    n, m = db.shape
    c = _i32(cols)
    if out is None:
        out = np.empty((max(n, 1), len(c)), dtype=db.dtype)
    n_out = C.c_int64(0)
    rc = lib().oracle_query_filter(_p(db), C.c_int64(n), C.c_int64(m), C.c_int32(DTYPE_CODES[np.dtype(db.dtype)]),
                                   _p(c), C.c_int64(len(c)), make_preds(preds), C.c_int64(len(preds)), _p(out),
                                   C.byref(n_out), C.c_int32(threads))
    if rc:
        raise IndexError("bad column / predicate")
    return out[:n_out.value]


def max_threads() -> int:
    return int(lib().oracle_max_threads())


def mix64(seed: int, col: int, row: int) -> int:
    return int(lib().oracle_mix64(seed & 0xFFFFFFFFFFFFFFFF, col, row))


def synth_column(dtype: int, spec: dict, seed: int, col: int, row0: int, n: int, threads: int = 1) -> np.ndarray:
    from .np_oracle import NP_DTYPES
    cs = HarkColspec(int(spec.get("kind", 0)), 0, int(spec.get("lo", 0)), int(spec.get("range", 0)),
                     float(spec.get("flo", 0.0)), float(spec.get("fhi", 1.0)), int(spec.get("a", 1)),
                     int(spec.get("b", 0)))
    out = np.empty(n, dtype=NP_DTYPES[dtype])
    rc = lib().oracle_synth_column(C.c_int32(dtype), C.byref(cs), C.c_uint64(seed & 0xFFFFFFFFFFFFFFFF),
                                   C.c_int32(col), C.c_int64(row0), C.c_int64(n), _p(out), C.c_int32(threads))
    if rc:
        raise ValueError("bad dtype")
    return out


# ------------------------------------------------------------------------------------------------
# CPU arms timed by bench.py (oracle/cpu_arms.c): cpu-ref-mt and cpu-best.  Each returns what the query returns,
# so tests/test_cpu_arms.py can hold them to np_oracle.
# ------------------------------------------------------------------------------------------------
def arm_filter_best_f32(cols, p0, c0, p1, c1, s0, s1, threads=1, out=None):
    """SELECT s0, s1 WHERE p0 > c0 AND p1 < c1 over f32 SoA columns.  -> (out0, out1, chunk starts, chunk counts)."""
    n = len(cols[0])
    T = max(1, int(threads))
    o0, o1 = out if out is not None else (np.empty(n, np.float32), np.empty(n, np.float32))
    counts = np.zeros(T, dtype=np.int64)
    f = lib().arm_filter_best_f32
    f.restype = C.c_int64
    total = f(_p(cols[p0]), C.c_float(c0), _p(cols[p1]), C.c_float(c1), _p(cols[s0]), _p(cols[s1]), C.c_int64(n), _p(o0),
              _p(o1), _p(counts), C.c_int32(T))
    starts = np.array([n * t // T for t in range(T)], dtype=np.int64)
    return o0, o1, starts, counts, int(total)


def arm_filter_rowmajor_f32(db, p0, c0, p1, c1, s0, s1, threads=1, out=None):
    db = np.ascontiguousarray(db, dtype=np.float32)
    n, m = db.shape
    T = max(1, int(threads))
    o = out if out is not None else np.empty((n, 2), np.float32)
    counts = np.zeros(T, dtype=np.int64)
    f = lib().arm_filter_rowmajor_f32
    f.restype = C.c_int64
    total = f(_p(db), C.c_int64(n), C.c_int64(m), C.c_int32(p0), C.c_float(c0), C.c_int32(p1), C.c_float(c1), C.c_int32(s0),
              C.c_int32(s1), _p(o), _p(counts), C.c_int32(T))
    starts = np.array([n * t // T for t in range(T)], dtype=np.int64)
    return o, starts, counts, int(total)


def chunks_concat(arr, starts, counts):
    return np.concatenate([arr[s:s + c] for s, c in zip(starts, counts)]) if len(starts) else arr[:0]


def arm_groupby_best(key, val, kmin, krange, threads=1):
    """-> (count[krange] i64, sum[krange] i64 or f64) over the dense key range."""
    key = np.ascontiguousarray(key, dtype=np.int32)
    cnt = np.empty(krange, dtype=np.int64)
    if val.dtype == np.float32:
        s = np.empty(krange, dtype=np.float64)
        rc = lib().arm_groupby_best_f32(_p(key), _p(np.ascontiguousarray(val)), C.c_int64(len(key)), C.c_int32(kmin),
                                        C.c_int64(krange), _p(cnt), _p(s), C.c_int32(threads))
    else:
        s = np.empty(krange, dtype=np.int64)
        rc = lib().arm_groupby_best_i32(_p(key), _p(np.ascontiguousarray(val, dtype=np.int32)), C.c_int64(len(key)),
                                        C.c_int32(kmin), C.c_int64(krange), _p(cnt), _p(s), C.c_int32(threads))
    if rc:
        raise MemoryError("arm_groupby_best")
    return cnt, s


def arm_groupby_ref_mt(rows, t_cols, threads=1):
    """rows [n][1+c] u32 (column 0 = key) -> [G][1+c] u32, groupby.fut semantics, 32 parallelised one-bit passes."""
    rows = np.ascontiguousarray(rows, dtype=np.uint32)
    n, s = rows.shape
    t = _i32(t_cols)
    out = np.empty((max(n, 1), s), dtype=np.uint32)
    f = lib().arm_groupby_ref_mt
    f.restype = C.c_int64
    g = f(_p(rows), C.c_int64(n), C.c_int64(s), _p(t), _p(out), C.c_int32(threads))
    if g < 0:
        raise MemoryError("arm_groupby_ref_mt")
    return out[:g].copy()


def arm_orderby_i64x2(a, b, digit_bits=8, threads=1):
    """Stable ORDER BY a, b (ascending, signed); returns sorted copies."""
    a = np.array(a, dtype=np.int64)
    b = np.array(b, dtype=np.int64)
    ta, tb = np.empty_like(a), np.empty_like(b)
    rc = lib().arm_orderby_i64x2(_p(a), _p(b), C.c_int64(len(a)), _p(ta), _p(tb), C.c_int32(digit_bits), C.c_int32(threads))
    if rc:
        raise MemoryError("arm_orderby")
    return a, b


def arm_join_groupby_best(fk, val, pk, attr, threads=1):
    """-> (attr values present, sum i64, count i64), ascending attr; unmatched fact rows drop out."""
    fk, val = np.ascontiguousarray(fk, np.int32), np.ascontiguousarray(val, np.int32)
    pk, attr = np.ascontiguousarray(pk, np.int32), np.ascontiguousarray(attr, np.int32)
    pk_min, pk_span = int(pk.min()), int(pk.max()) - int(pk.min()) + 1
    g_min, g_range = int(attr.min()), int(attr.max()) - int(attr.min()) + 1
    lut = np.empty(pk_span, dtype=np.int32)
    cnt, s = np.empty(g_range, np.int64), np.empty(g_range, np.int64)
    rc = lib().arm_join_groupby_best(_p(fk), _p(val), C.c_int64(len(fk)), _p(pk), _p(attr), C.c_int64(len(pk)),
                                     C.c_int64(pk_min), C.c_int64(pk_span), C.c_int32(g_min), C.c_int64(g_range), _p(lut),
                                     _p(cnt), _p(s), C.c_int32(threads))
    if rc:
        raise MemoryError("arm_join_groupby_best")
    keep = cnt > 0
    return (np.arange(g_range, dtype=np.int64) + g_min)[keep].astype(np.int32), s[keep], cnt[keep]
