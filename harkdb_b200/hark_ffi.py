"""ctypes binding of libhark.so — the drop-in for ``futhark_ffi.Futhark(_main)``.

The reference drives its compiled operators through ``futhark_ffi`` (FutharkContext.py:31,41):
``Futhark(_main)`` exposes one method per Futhark entry (``query_sel``, ``query_groupby``) plus
``from_futhark(handle) -> ndarray`` (FutharkContext.py:65-66,70-71).  :class:`Futhark` below keeps
exactly that call shape on top of the C-ABI in ``include/hark.h``; the extra methods are the
extensions (filter, typed group-by, order-by, join).

There is no CPU fallback: if ``libhark.so`` is missing, or no B200 is usable, construction raises.
"""

from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HARK_LIB") or os.path.join(_HERE, "libhark.so")   # HARK_LIB: kernel-variant experiments

I32, U32, I64, F32, F64 = 0, 1, 2, 3, 4
NP_DTYPES = {I32: np.dtype(np.int32), U32: np.dtype(np.uint32), I64: np.dtype(np.int64),
             F32: np.dtype(np.float32), F64: np.dtype(np.float64)}
DTYPE_CODES = {v: k for k, v in NP_DTYPES.items()}

GT, GE, LT, LE, EQ, NE = 0, 1, 2, 3, 4, 5
CMP_CODES = {">": GT, ">=": GE, "<": LT, "<=": LE, "=": EQ, "==": EQ, "!=": NE, "<>": NE,
             "gt": GT, "gte": GE, "lt": LT, "lte": LE, "eq": EQ, "neq": NE}
AGG_KEY, AGG_PROD, AGG_SUM, AGG_MAX, AGG_MIN, AGG_COUNT, AGG_AVG, AGG_SUMF64, AGG_SUM64 = 0, 1, 2, 3, 4, 5, 6, 7, 8
GEN_UNIFORM, GEN_AFFINE, GEN_CONST, GEN_LOGUNIFORM, GEN_AFFINE_UNIFORM = 0, 1, 2, 3, 4

STATUS = {0: "HARK_OK", 1: "HARK_ERR_ARG", 2: "HARK_ERR_CUDA", 3: "HARK_ERR_OOM", 4: "HARK_ERR_UNSUPPORTED"}


class HarkError(Exception):
    """A libhark entry returned non-zero (the analogue of a Futhark entry's error string)."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"{STATUS.get(code, code)}: {msg}")
        self.code = code


class HarkPred(C.Structure):
    _fields_ = [("col", C.c_int32), ("op", C.c_int32), ("ival", C.c_int64), ("fval", C.c_double)]


class HarkColspec(C.Structure):
    _fields_ = [("kind", C.c_int32), ("reserved", C.c_int32), ("lo", C.c_int64), ("range", C.c_uint64),
                ("flo", C.c_double), ("fhi", C.c_double), ("a", C.c_uint64), ("b", C.c_uint64)]


class HarkStats(C.Structure):
    _fields_ = [("kernel_ms", C.c_double), ("total_ms", C.c_double), ("alg_bytes", C.c_int64),
                ("rows_in", C.c_int64), ("rows_out", C.c_int64), ("launches", C.c_int64)]


_P = C.c_void_p
_I32P = C.POINTER(C.c_int32)

# name -> (restype, argtypes): every symbol include/hark.h declares
SIGNATURES = {
    "hark_abi_version": (C.c_int, []),
    "hark_context_new": (_P, [C.c_int, _P]),
    "hark_context_free": (None, [_P]),
    "hark_context_sync": (C.c_int, [_P]),
    "hark_context_trim": (C.c_int, [_P]),
    "hark_context_get_error": (_P, [_P]),
    "hark_context_device": (C.c_int, [_P]),
    "hark_last_init_error": (C.c_char_p, []),
    "hark_table_from_host": (C.c_int, [_P, C.POINTER(_P), _P, C.c_int64, C.c_int64, C.c_int32]),
    "hark_table_from_columns": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_P), _I32P, C.c_int64, C.c_int64]),
    "hark_table_from_device": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_P), _I32P, C.c_int64, C.c_int64]),
    "hark_table_synth": (C.c_int, [_P, C.POINTER(_P), C.c_int64, C.c_int64, _I32P, C.c_uint64,
                                   C.POINTER(HarkColspec), C.c_int64]),
    "hark_table_shape": (C.c_int, [_P, _P, C.POINTER(C.c_int64)]),
    "hark_table_dtypes": (C.c_int, [_P, _P, _I32P]),
    "hark_table_to_host": (C.c_int, [_P, _P, _P]),
    "hark_table_column_to_host": (C.c_int, [_P, _P, C.c_int32, C.c_int64, C.c_int64, _P]),
    "hark_table_column_ptr": (_P, [_P, _P, C.c_int32]),
    "hark_table_free": (C.c_int, [_P, _P]),
    "hark_entry_query_sel": (C.c_int, [_P, C.POINTER(_P), _P, _I32P, C.c_int64]),
    "hark_entry_query_groupby": (C.c_int, [_P, C.POINTER(_P), _P, C.c_int32, _I32P, _I32P, C.c_int64]),
    "hark_entry_join": (C.c_int, [_P, C.POINTER(_P), _P, _P, C.c_int32, C.c_int32, _I32P, C.c_int64, _I32P,
                                  C.c_int64]),
    "hark_entry_join_ex": (C.c_int, [_P, C.POINTER(_P), _P, _P, C.c_int32, C.c_int32, _I32P, C.c_int64, _I32P,
                                     C.c_int64, C.c_int32]),
    "hark_entry_query_filter": (C.c_int, [_P, C.POINTER(_P), _P, _I32P, C.c_int64, C.POINTER(HarkPred), C.c_int64]),
    "hark_entry_query_groupby_ex": (C.c_int, [_P, C.POINTER(_P), _P, C.c_int32, _I32P, _I32P, C.c_int64,
                                              C.POINTER(HarkPred), C.c_int64]),
    "hark_entry_query_groupby_multi": (C.c_int, [_P, C.POINTER(_P), _P, _I32P, C.c_int64, _I32P, _I32P, C.c_int64,
                                                 C.POINTER(HarkPred), C.c_int64]),
    "hark_entry_query_orderby": (C.c_int, [_P, C.POINTER(_P), _P, _I32P, C.c_int64, _I32P, _I32P, C.c_int64]),
    "hark_entry_join_groupby": (C.c_int, [_P, C.POINTER(_P), _P, _P, C.c_int32, C.c_int32, C.c_int32, _I32P, _I32P,
                                          C.c_int64]),
    "hark_table_sort_by": (C.c_int, [_P, C.POINTER(_P), _P, C.c_int32]),
    "hark_table_partition_by_hash": (C.c_int, [_P, C.POINTER(_P), _P, C.c_int32, C.c_int32, C.POINTER(C.c_int64)]),
    "hark_table_partition_by_splitters": (C.c_int, [_P, C.POINTER(_P), _P, _I32P, _I32P, C.c_int64,
                                                    C.POINTER(C.c_uint64), C.c_int32, C.POINTER(C.c_int64)]),
    "hark_table_sample_order_keys": (C.c_int, [_P, _P, _I32P, _I32P, C.c_int64, C.POINTER(C.c_int64), C.c_int64,
                                               C.POINTER(C.c_uint64)]),
    "hark_entry_groupby_finalize": (C.c_int, [_P, C.POINTER(_P), _P, _I32P, C.c_int64]),
    "hark_peer_arena_create": (C.c_int, [_P, C.c_int64, _P]),
    "hark_peer_arena_open": (C.c_int, [_P, _P, C.c_int32, C.c_int32]),
    "hark_peer_arena_close": (C.c_int, [_P]),
    "hark_peer_arena_bytes_needed": (C.c_int64, [_P, _P, C.c_int64]),
    "hark_peer_scatter_count": (C.c_int, [_P, _P, _I32P, _I32P, C.c_int64, C.POINTER(C.c_uint64), C.POINTER(C.c_int64)]),
    "hark_peer_scatter_run": (C.c_int, [_P, _P, C.POINTER(C.c_int64)]),
    "hark_peer_scatter_result": (C.c_int, [_P, C.POINTER(_P)]),
    "hark_table_slice": (C.c_int, [_P, C.POINTER(_P), _P, C.c_int64, C.c_int64]),
    "hark_table_concat": (C.c_int, [_P, C.POINTER(_P), _P, _P]),
    "hark_stats_last": (C.c_int, [_P, C.POINTER(HarkStats)]),
    "hark_stats_total_launches": (C.c_int64, [_P]),
    "hark_context_set_option": (C.c_int, [_P, C.c_char_p, C.c_int64]),
    "hark_context_get_option": (C.c_int, [_P, C.c_char_p, C.POINTER(C.c_int64)]),
    "hark_debug_plan_truncation": (C.c_int, [C.c_int64, C.c_int32, _I32P, C.POINTER(C.c_uint64), C.c_int32, C.c_int32,
                                             _I32P, _I32P, _I32P, _I32P]),
    "hark_host_alloc": (_P, [C.c_int64]),
    "hark_host_free": (None, [_P]),
}

_LIB: Optional[C.CDLL] = None

LEGACY_DEFAULT_STREAM = 1   # cudaStreamLegacy: CUDA's handle for "the default stream" that is not the NULL pointer


def torch_stream_handle() -> int:
    """cudaStream_t of torch's current stream, usable as `Futhark(stream=...)`.  torch's default stream has handle
    0, which this ABI reads as "create your own stream", so it is passed as cudaStreamLegacy instead: libhark's
    kernels are then ordered with torch's kernels and with the collectives torch.distributed launches."""
    import torch
    return int(torch.cuda.current_stream().cuda_stream) or LEGACY_DEFAULT_STREAM


def load_library(path: str = LIB_PATH) -> C.CDLL:
    """dlopen libhark.so and type every exported symbol.  Raises if the library was not built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(path):
        raise ImportError(f"{path} not found: build it with `make -C harkdb_b200/csrc` "
                          f"(or `python -c 'import __graft_entry__ as g; g.build()'`); there is no CPU fallback")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the ABI header and the library disagree
        fn.restype = res
        fn.argtypes = args
    if lib.hark_abi_version() != 1:
        raise ImportError("libhark.so ABI version mismatch")
    _LIB = lib
    return lib


def _i32arr(xs) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(xs, dtype=np.int64).reshape(-1).astype(np.int32))


def _i32p(a: np.ndarray):
    return a.ctypes.data_as(_I32P)


def make_preds(preds: Sequence[Tuple[int, int, int, float]]):
    arr = (HarkPred * max(len(preds), 1))()
    for i, p in enumerate(preds):
        c, op, iv, fv = p
        arr[i] = HarkPred(int(c), int(op), int(iv), float(fv))
    return arr


def convert_for_entry(arr: np.ndarray, want: np.dtype, what: str) -> np.ndarray:
    """Value-preserving cast with a range check (SURVEY.md §8b dtype hazard: pandas hands the
    reference int64 while the Futhark entries take i32/u32)."""
    arr = np.asarray(arr)
    want = np.dtype(want)
    if arr.dtype == want:
        return arr
    if arr.dtype.kind in "iu" and want.kind in "iu":
        info = np.iinfo(want)
        if arr.size and (arr.min() < info.min or arr.max() > info.max):
            raise HarkError(1, f"{what}: values do not fit {want} (min {arr.min()}, max {arr.max()})")
        return arr.astype(want)
    if arr.dtype.kind == "f" and want.kind in "iu":
        r = np.rint(arr)
        if arr.size and (not np.all(np.isfinite(arr)) or np.any(r != arr)):
            raise HarkError(1, f"{what}: non-integral values cannot be passed to an integer entry")
        return convert_for_entry(r.astype(np.int64), want, what)
    return arr.astype(want)


class DeviceTable:
    """Owned handle of a device-resident SoA table (the analogue of an opaque futhark_*_2d)."""

    def __init__(self, env: "Futhark", handle: int):
        self._env = env
        self._h = C.c_void_p(handle)

    @property
    def handle(self):
        if self._h is None:
            raise HarkError(1, "table already freed")
        return self._h

    @property
    def shape(self) -> Tuple[int, int]:
        s = (C.c_int64 * 2)()
        self._env._check(self._env.lib.hark_table_shape(self._env.ctx, self.handle, s))
        return int(s[0]), int(s[1])

    @property
    def dtypes(self) -> List[int]:
        m = self.shape[1]
        d = (C.c_int32 * max(m, 1))()
        self._env._check(self._env.lib.hark_table_dtypes(self._env.ctx, self.handle, d))
        return [int(d[i]) for i in range(m)]

    def column_ptr(self, col: int) -> int:
        return int(self._env.lib.hark_table_column_ptr(self._env.ctx, self.handle, col) or 0)

    def column(self, col: int, row0: int = 0, nrows: Optional[int] = None) -> np.ndarray:
        n, m = self.shape
        nrows = n - row0 if nrows is None else nrows
        out = np.empty(nrows, dtype=NP_DTYPES[self.dtypes[col]])
        self._env._check(self._env.lib.hark_table_column_to_host(self._env.ctx, self.handle, col, row0, nrows,
                                                                 out.ctypes.data_as(_P)))
        return out

    def columns(self) -> List[np.ndarray]:
        return [self.column(c) for c in range(self.shape[1])]

    def to_numpy(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        """Row-major [n][m] host array.  Mixed-dtype results (COUNT/AVG columns) widen to float64."""
        n, m = self.shape
        dts = self.dtypes
        if m == 0 or len(set(dts)) == 1:
            dt = NP_DTYPES[dts[0]] if m else np.dtype(np.int32)
            if out is None:
                out = np.empty((n, m), dtype=dt)
            assert out.dtype == dt and out.shape == (n, m) and out.flags.c_contiguous
            self._env._check(self._env.lib.hark_table_to_host(self._env.ctx, self.handle, out.ctypes.data_as(_P)))
            return out
        return np.stack([c.astype(np.float64) for c in self.columns()], axis=1)

    def free(self):
        if self._h is not None and self._env.ctx:
            self._env.lib.hark_table_free(self._env.ctx, self._h)
        self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


TableLike = Union[np.ndarray, DeviceTable]


class Futhark:
    """Same role and call shape as ``futhark_ffi.Futhark(_main)`` (FutharkContext.py:41)."""

    def __init__(self, device: int = -1, stream: int = 0, lib_path: str = LIB_PATH):
        self.lib = load_library(lib_path)
        self.ctx = self.lib.hark_context_new(device, C.c_void_p(stream) if stream else None)
        if not self.ctx:
            msg = self.lib.hark_last_init_error()
            raise HarkError(2, (msg or b"context creation failed").decode())
        self.ctx = C.c_void_p(self.ctx)

    # ---- plumbing ----
    def _check(self, rc: int):
        if rc != 0:
            p = self.lib.hark_context_get_error(self.ctx)
            msg = C.string_at(p).decode() if p else "(no message)"
            if p:
                C.CDLL(None).free(C.c_void_p(p))
            raise HarkError(rc, msg)

    def close(self):
        if self.ctx:
            self.lib.hark_context_free(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        self._check(self.lib.hark_context_sync(self.ctx))

    def trim(self):
        """Hands the memory pool's free blocks back to the driver (between workloads of very different sizes)."""
        self._check(self.lib.hark_context_trim(self.ctx))

    def set_option(self, key: str, value: int):
        self._check(self.lib.hark_context_set_option(self.ctx, key.encode(), int(value)))

    def get_option(self, key: str) -> int:
        v = C.c_int64(0)
        self._check(self.lib.hark_context_get_option(self.ctx, key.encode(), C.byref(v)))
        return v.value

    def stats(self) -> dict:
        s = HarkStats()
        self._check(self.lib.hark_stats_last(self.ctx, C.byref(s)))
        return {f: getattr(s, f) for f, _ in HarkStats._fields_}

    def total_launches(self) -> int:
        return int(self.lib.hark_stats_total_launches(self.ctx))

    # ---- tables ----
    def to_device(self, arr: np.ndarray, dtype: Optional[np.dtype] = None) -> DeviceTable:
        """Row-major 2-D host array -> resident SoA table (futhark_new_*_2d)."""
        arr = np.asarray(arr)
        if arr.ndim != 2:
            raise HarkError(1, "table data must be 2-D")
        if dtype is not None:
            arr = convert_for_entry(arr, dtype, "table")
        if arr.dtype not in DTYPE_CODES:
            if arr.dtype.kind in "iub":
                arr = convert_for_entry(arr, np.int64, "table")
            elif arr.dtype.kind == "f":
                arr = arr.astype(np.float64)
            else:
                raise HarkError(1, f"unsupported table dtype {arr.dtype}")
        arr = np.ascontiguousarray(arr)
        h = C.c_void_p()
        self._check(self.lib.hark_table_from_host(self.ctx, C.byref(h), arr.ctypes.data_as(_P), arr.shape[0],
                                                  arr.shape[1], DTYPE_CODES[arr.dtype]))
        return DeviceTable(self, h.value)

    def from_columns(self, cols: Sequence[np.ndarray]) -> DeviceTable:
        cols = [np.ascontiguousarray(c) for c in cols]
        m = len(cols)
        n = len(cols[0]) if m else 0
        ptrs = (_P * max(m, 1))(*[c.ctypes.data_as(_P) for c in cols])
        dts = (C.c_int32 * max(m, 1))(*[DTYPE_CODES[c.dtype] for c in cols])
        h = C.c_void_p()
        self._check(self.lib.hark_table_from_columns(self.ctx, C.byref(h), ptrs, dts, n, m))
        return DeviceTable(self, h.value)

    def from_device_pointers(self, ptrs: Sequence[int], dtypes: Sequence[int], n: int) -> DeviceTable:
        """Borrow device column arrays (e.g. torch tensors' data_ptr()); the caller keeps them alive."""
        m = len(ptrs)
        pa = (_P * max(m, 1))(*[C.c_void_p(p) for p in ptrs])
        dts = (C.c_int32 * max(m, 1))(*dtypes)
        h = C.c_void_p()
        self._check(self.lib.hark_table_from_device(self.ctx, C.byref(h), pa, dts, n, m))
        return DeviceTable(self, h.value)

    def with_constant_key(self, t: DeviceTable) -> DeviceTable:
        """[constant i32 column 0] ++ the columns of `t` (borrowed, no copy): the table a GROUP BY-less aggregate groups."""
        n, m = t.shape
        key = self.synth(n, [I32], [dict(kind=GEN_CONST, lo=0)])
        if n == 0:
            view = self.from_columns([np.zeros(0, np.int32)] + [np.zeros(0, NP_DTYPES[d]) for d in t.dtypes])
            key.free()
            return view
        view = self.from_device_pointers([key.column_ptr(0)] + [t.column_ptr(c) for c in range(m)], [I32] + list(t.dtypes), n)
        view._keepalive = (key, t)
        return view

    def synth(self, n: int, dtypes: Sequence[int], specs: Sequence[dict], seed: int = 42, row0: int = 0) -> DeviceTable:
        m = len(dtypes)
        dts = (C.c_int32 * max(m, 1))(*dtypes)
        cs = (HarkColspec * max(m, 1))()
        for i, s in enumerate(specs):
            cs[i] = HarkColspec(int(s.get("kind", 0)), 0, int(s.get("lo", 0)), int(s.get("range", 0)),
                                float(s.get("flo", 0.0)), float(s.get("fhi", 1.0)), int(s.get("a", 1)),
                                int(s.get("b", 0)))
        h = C.c_void_p()
        self._check(self.lib.hark_table_synth(self.ctx, C.byref(h), n, m, dts, seed & 0xFFFFFFFFFFFFFFFF, cs, row0))
        return DeviceTable(self, h.value)

    def _as_table(self, db: TableLike, dtype=None) -> Tuple[DeviceTable, bool]:
        """(table, temporary?) — an ndarray is uploaded for this call only, like the reference does
        on every query (FutharkContext.py:65,70)."""
        if isinstance(db, DeviceTable):
            return db, False
        return self.to_device(db, dtype), True

    def from_futhark(self, handle: DeviceTable) -> np.ndarray:
        """futhark_ffi's from_futhark (FutharkContext.py:66,71): opaque result -> ndarray."""
        return handle.to_numpy()

    # ---- reference-pinned entries ----
    def query_sel(self, db: TableLike, cols) -> DeviceTable:
        """main.fut:7.  An ndarray `db` is converted to i32 like the Futhark entry demands."""
        t, tmp = self._as_table(db, np.int32)
        try:
            c = _i32arr(cols)
            h = C.c_void_p()
            self._check(self.lib.hark_entry_query_sel(self.ctx, C.byref(h), t.handle, _i32p(c), len(c)))
            return DeviceTable(self, h.value)
        finally:
            if tmp:
                t.free()

    def query_groupby(self, db: TableLike, g_col, s_cols, t_cols) -> DeviceTable:
        """main.fut:9.  An ndarray `db` is converted to u32 like the Futhark entry demands."""
        t, tmp = self._as_table(db, np.uint32)
        try:
            s, tc = _i32arr(s_cols), _i32arr(t_cols)
            if len(tc) < len(s):
                raise HarkError(1, "t_cols shorter than s_cols (groupby.fut:47)")
            h = C.c_void_p()
            self._check(self.lib.hark_entry_query_groupby(self.ctx, C.byref(h), t.handle, int(g_col), _i32p(s),
                                                          _i32p(tc), len(s)))
            return DeviceTable(self, h.value)
        finally:
            if tmp:
                t.free()

    def join(self, db1: TableLike, db2: TableLike, col1, col2, cols1, cols2) -> DeviceTable:
        """join.fut:52 (the reference never wires it to Python)."""
        t1, tmp1 = self._as_table(db1, np.uint32)
        t2, tmp2 = self._as_table(db2, np.uint32)
        try:
            c1, c2 = _i32arr(cols1), _i32arr(cols2)
            h = C.c_void_p()
            self._check(self.lib.hark_entry_join(self.ctx, C.byref(h), t1.handle, t2.handle, int(col1), int(col2),
                                                 _i32p(c1), len(c1), _i32p(c2), len(c2)))
            return DeviceTable(self, h.value)
        finally:
            if tmp1:
                t1.free()
            if tmp2:
                t2.free()

    def join_ex(self, db1: TableLike, db2: TableLike, col1, col2, cols1, cols2, order: int = 1) -> DeviceTable:
        """Typed join: integer key columns of one dtype, projected columns keep their dtypes.  order=1: reference order
        (key, r1, r2) by sort + merge; order=0: hash build on db2 + probe with db1 (multiset, grouped by db1 row)."""
        t1, tmp1 = self._as_table(db1)
        t2, tmp2 = self._as_table(db2)
        try:
            c1, c2 = _i32arr(cols1), _i32arr(cols2)
            h = C.c_void_p()
            self._check(self.lib.hark_entry_join_ex(self.ctx, C.byref(h), t1.handle, t2.handle, int(col1), int(col2),
                                                    _i32p(c1), len(c1), _i32p(c2), len(c2), int(order)))
            return DeviceTable(self, h.value)
        finally:
            if tmp1:
                t1.free()
            if tmp2:
                t2.free()

    # ---- extensions ----
    def query_filter(self, db: TableLike, cols, preds) -> DeviceTable:
        t, tmp = self._as_table(db)
        try:
            c = _i32arr(cols)
            h = C.c_void_p()
            self._check(self.lib.hark_entry_query_filter(self.ctx, C.byref(h), t.handle, _i32p(c), len(c),
                                                         make_preds(preds), len(preds)))
            return DeviceTable(self, h.value)
        finally:
            if tmp:
                t.free()

    def query_groupby_ex(self, db: TableLike, g_col, s_cols, ops, having=()) -> DeviceTable:
        t, tmp = self._as_table(db)
        try:
            s, o = _i32arr(s_cols), _i32arr(ops)
            h = C.c_void_p()
            self._check(self.lib.hark_entry_query_groupby_ex(self.ctx, C.byref(h), t.handle, int(g_col), _i32p(s),
                                                             _i32p(o), len(s), make_preds(having), len(having)))
            return DeviceTable(self, h.value)
        finally:
            if tmp:
                t.free()

    def query_groupby_multi(self, db: TableLike, g_cols, s_cols, ops, having=()) -> DeviceTable:
        """GROUP BY several integer columns: output = the key columns, then one column per aggregate."""
        t, tmp = self._as_table(db)
        try:
            g, s, o = _i32arr(g_cols), _i32arr(s_cols), _i32arr(ops)
            h = C.c_void_p()
            self._check(self.lib.hark_entry_query_groupby_multi(self.ctx, C.byref(h), t.handle, _i32p(g), len(g), _i32p(s),
                                                                _i32p(o), len(s), make_preds(having), len(having)))
            return DeviceTable(self, h.value)
        finally:
            if tmp:
                t.free()

    def query_orderby(self, db: TableLike, cols, key_cols, desc=None) -> DeviceTable:
        t, tmp = self._as_table(db)
        try:
            c, kc = _i32arr(cols), _i32arr(key_cols)
            d = _i32arr(desc if desc is not None else [0] * len(kc))
            h = C.c_void_p()
            self._check(self.lib.hark_entry_query_orderby(self.ctx, C.byref(h), t.handle, _i32p(c), len(c), _i32p(kc),
                                                          _i32p(d), len(kc)))
            return DeviceTable(self, h.value)
        finally:
            if tmp:
                t.free()

    def join_groupby(self, fact: TableLike, dim: TableLike, fk_col, pk_col, g_col, s_cols, ops) -> DeviceTable:
        tf, tmpf = self._as_table(fact)
        td, tmpd = self._as_table(dim)
        try:
            s, o = _i32arr(s_cols), _i32arr(ops)
            h = C.c_void_p()
            self._check(self.lib.hark_entry_join_groupby(self.ctx, C.byref(h), tf.handle, td.handle, int(fk_col),
                                                         int(pk_col), int(g_col), _i32p(s), _i32p(o), len(s)))
            return DeviceTable(self, h.value)
        finally:
            if tmpf:
                tf.free()
            if tmpd:
                td.free()

    # ---- building blocks for the multi-GPU layer ----
    def sort_by(self, t: DeviceTable, key_col: int) -> DeviceTable:
        h = C.c_void_p()
        self._check(self.lib.hark_table_sort_by(self.ctx, C.byref(h), t.handle, int(key_col)))
        return DeviceTable(self, h.value)

    def partition_by_hash(self, t: DeviceTable, key_col: int, nparts: int) -> Tuple[DeviceTable, List[int]]:
        h = C.c_void_p()
        counts = (C.c_int64 * nparts)()
        self._check(self.lib.hark_table_partition_by_hash(self.ctx, C.byref(h), t.handle, int(key_col), nparts,
                                                          counts))
        return DeviceTable(self, h.value), [int(x) for x in counts]

    def partition_by_splitters(self, t: DeviceTable, key_cols, desc, splitters: np.ndarray,
                               nparts: int) -> Tuple[DeviceTable, List[int]]:
        """Stable range partition: bucket of a row = number of splitter tuples <= its key tuple.
        `splitters`: uint64 order keys [nparts-1][len(key_cols)], ascending (see sample_order_keys)."""
        kc = _i32arr(key_cols)
        d = _i32arr(desc if desc is not None else [0] * len(kc))
        sp = np.ascontiguousarray(np.asarray(splitters, dtype=np.uint64).reshape(-1))
        if sp.size != (nparts - 1) * len(kc):
            raise HarkError(1, "partition_by_splitters: need (nparts-1) x nk splitter keys")
        h = C.c_void_p()
        counts = (C.c_int64 * nparts)()
        self._check(self.lib.hark_table_partition_by_splitters(
            self.ctx, C.byref(h), t.handle, _i32p(kc), _i32p(d), len(kc),
            sp.ctypes.data_as(C.POINTER(C.c_uint64)) if sp.size else None, nparts, counts))
        return DeviceTable(self, h.value), [int(x) for x in counts]

    def sample_order_keys(self, t: DeviceTable, key_cols, desc, rows) -> np.ndarray:
        """uint64 order keys [len(rows)][len(key_cols)] of the given rows (ORDER BY's key mapping)."""
        kc = _i32arr(key_cols)
        d = _i32arr(desc if desc is not None else [0] * len(kc))
        r = np.ascontiguousarray(np.asarray(rows, dtype=np.int64).reshape(-1))
        out = np.empty((len(r), len(kc)), dtype=np.uint64)
        self._check(self.lib.hark_table_sample_order_keys(
            self.ctx, t.handle, _i32p(kc), _i32p(d), len(kc), r.ctypes.data_as(C.POINTER(C.c_int64)), len(r),
            out.ctypes.data_as(C.POINTER(C.c_uint64))))
        return out

    # ---- K8c: partition fused with the exchange over NVLink peer memory ----
    def peer_arena_create(self, nbytes: int) -> bytes:
        """Allocates this rank's receive arena; returns its 64-byte CUDA IPC handle (to be all-gathered)."""
        buf = C.create_string_buffer(64)
        self._check(self.lib.hark_peer_arena_create(self.ctx, int(nbytes), buf))
        return buf.raw

    def peer_arena_open(self, handles: Sequence[bytes], my_rank: int):
        blob = b"".join(handles)
        self._check(self.lib.hark_peer_arena_open(self.ctx, C.c_char_p(blob), len(handles), my_rank))

    def peer_arena_close(self):
        self._check(self.lib.hark_peer_arena_close(self.ctx))

    def peer_arena_bytes_needed(self, t: DeviceTable, rows: int) -> int:
        return int(self.lib.hark_peer_arena_bytes_needed(self.ctx, t.handle, int(rows)))

    def peer_scatter_count(self, t: DeviceTable, key_cols, desc, splitters: np.ndarray, world: int) -> List[int]:
        kc = _i32arr(key_cols)
        d = _i32arr(desc if desc is not None else [0] * len(kc))
        sp = np.ascontiguousarray(np.asarray(splitters, dtype=np.uint64).reshape(-1))
        counts = (C.c_int64 * world)()
        self._check(self.lib.hark_peer_scatter_count(
            self.ctx, t.handle, _i32p(kc), _i32p(d), len(kc),
            sp.ctypes.data_as(C.POINTER(C.c_uint64)) if sp.size else None, counts))
        return [int(x) for x in counts]

    def peer_scatter_run(self, t: DeviceTable, counts_matrix: np.ndarray):
        cm = np.ascontiguousarray(np.asarray(counts_matrix, dtype=np.int64))
        self._check(self.lib.hark_peer_scatter_run(self.ctx, t.handle, cm.ctypes.data_as(C.POINTER(C.c_int64))))

    def peer_scatter_result(self) -> DeviceTable:
        h = C.c_void_p()
        self._check(self.lib.hark_peer_scatter_result(self.ctx, C.byref(h)))
        return DeviceTable(self, h.value)

    def groupby_finalize(self, merged: DeviceTable, ops) -> DeviceTable:
        o = _i32arr(ops)
        h = C.c_void_p()
        self._check(self.lib.hark_entry_groupby_finalize(self.ctx, C.byref(h), merged.handle, _i32p(o), len(o)))
        return DeviceTable(self, h.value)

    def slice(self, t: DeviceTable, row0: int, nrows: int) -> DeviceTable:
        h = C.c_void_p()
        self._check(self.lib.hark_table_slice(self.ctx, C.byref(h), t.handle, row0, nrows))
        return DeviceTable(self, h.value)

    def concat(self, a: DeviceTable, b: DeviceTable) -> DeviceTable:
        h = C.c_void_p()
        self._check(self.lib.hark_table_concat(self.ctx, C.byref(h), a.handle, b.handle))
        return DeviceTable(self, h.value)
