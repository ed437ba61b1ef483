"""Shared helpers for the -m gpu parity tests (they call the product through the C-ABI/ctypes shim and
check it against the oracle)."""

import numpy as np
import pytest

from oracle import np_oracle as NO


def need_gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


_ENV = None


def get_env():
    """One libhark context for the whole test session."""
    global _ENV
    need_gpu()
    if _ENV is None:
        from harkdb_b200 import hark_ffi
        _ENV = hark_ffi.Futhark()
    return _ENV


def rand_table(rng, n, m, dtype, lo=0, hi=100, nan_frac=0.0):
    npdt = NO.NP_DTYPES[dtype]
    if dtype in (NO.F32, NO.F64):
        a = rng.random((n, m)).astype(npdt)
        if nan_frac > 0 and n * m:
            k = int(n * m * nan_frac)
            a[rng.integers(0, n, k), rng.integers(0, m, k)] = np.nan
        return a
    return rng.integers(lo, hi, (n, m)).astype(npdt)


def cols_of(a):
    return [np.ascontiguousarray(a[:, c]) for c in range(a.shape[1])]


def free_gb():
    import torch
    free, _ = torch.cuda.mem_get_info()
    return free / 2 ** 30


class _CAI:
    """Minimal __cuda_array_interface__ holder so tests can look at a device column with torch (checker only)."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


_TYPESTR = {0: "<i4", 1: "<u4", 2: "<i8", 3: "<f4", 4: "<f8"}


def as_torch(table, col):
    """Zero-copy torch view of a device column (the table must stay alive)."""
    import torch
    table._env.sync()    # libhark launches on its own stream; torch must not read the column before it is written
    n = table.shape[0]
    dt = table.dtypes[col]
    ts = _TYPESTR[dt]
    if dt == 1:          # torch has no uint32 arithmetic: view the bits as int32
        ts = "<i4"
    return torch.as_tensor(_CAI(table.column_ptr(col), n, ts), device="cuda")
