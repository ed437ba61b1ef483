"""Line-by-line Python simulation of the reference's Futhark programs — TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline leg may import this.

Why it exists: the Futhark compiler cannot be built here, so the reference
cannot be executed.  This module re-expresses every function on the hot path with the SOACs
of ``oracle/soac.py`` in the same order the sources apply them, so that the fast oracles
(``oracle/oracle.c``, ``oracle/np_oracle.py``) can be checked against something that is a
*reading* of the source rather than a re-derivation.

Pinning status:
  * segmented_* / replicated_iota / expand*  — pinned by the reference's own known-answer
    tests (futhark/lib/github.com/diku-dk/segmented/segmented_tests.fut:5-72); see
    tests/golden/segmented_kats.json and tests/test_oracle_kats.py.
  * query_sel / query_groupby / join          — PARITY UNPINNED by any reference test: the
    reference holds no expected outputs for them (test.py prints, asserts nothing).  They
    are pinned by source reading only.

All integer arithmetic is u32 wrap-around for groupby/join values and i32 for indices, as in
the sources.  Pure-Python loops: use for small tables only.
"""

from __future__ import annotations

from typing import Callable, List, Sequence, Tuple

from . import soac as S


# --------------------------------------------------------------------------------------
# futhark/lib/github.com/diku-dk/segmented/segmented.fut
# --------------------------------------------------------------------------------------

def segmented_scan(op: Callable, ne, flags: Sequence[bool], xs: Sequence) -> list:
    """segmented.fut:7-13 — scan of (flag, value) pairs with the flag-lifted operator."""
    def lifted(a, b):
        (x_flag, x), (y_flag, y) = a, b
        return (x_flag or y_flag, y if y_flag else op(x, y))
    pairs = S.scan(lifted, (False, ne), list(zip(flags, xs)))
    return [p[1] for p in pairs]


def segmented_reduce(op: Callable, ne, flags: Sequence[bool], xs: Sequence) -> list:
    """segmented.fut:20-37 — segmented scan, then scatter out each segment's last element."""
    n = len(xs)
    scanned = segmented_scan(op, ne, flags, xs)                        # :24
    segment_ends = S.rotate(1, list(flags))                            # :26
    segment_end_offsets = S.scan(lambda a, b: a + b, 0,
                                 S.fmap(lambda f: 1 if f else 0, segment_ends))   # :28
    num_segments = S.last(segment_end_offsets) if n > 0 else 0         # :29
    scratch = S.replicate(num_segments, ne)                            # :33
    index = lambda i, f: i - 1 if f else -1                            # :36
    return S.scatter(scratch, S.map2(index, segment_end_offsets, segment_ends), scanned)  # :37


def replicated_iota(reps: Sequence[int]) -> List[int]:
    """segmented.fut:44-50."""
    n = len(reps)
    s1 = S.scan(lambda a, b: a + b, 0, reps)                           # :45
    s2 = S.map2(lambda i, x: 0 if i == 0 else x, S.iota(n), S.rotate(-1, s1))   # :46-47
    total = S.reduce(lambda a, b: a + b, 0, reps)
    tmp = S.reduce_by_index(S.replicate(total, 0), max, 0, s2, S.iota(n))       # :48
    flags = S.fmap(lambda v: v > 0, tmp)                               # :49
    return segmented_scan(lambda a, b: a + b, 0, flags, tmp)           # :50


def segmented_iota(flags: Sequence[bool]) -> List[int]:
    """segmented.fut:58-60."""
    iotas = segmented_scan(lambda a, b: a + b, 0, flags, S.replicate(len(flags), 1))
    return S.fmap(lambda x: x - 1, iotas)


def expand(sz: Callable, get: Callable, arr: Sequence) -> list:
    """segmented.fut:70-74."""
    szs = S.fmap(sz, arr)
    idxs = replicated_iota(szs)
    iotas = segmented_iota(S.map2(lambda a, b: a != b, idxs, S.rotate(-1, idxs)))
    return S.map2(lambda i, j: get(arr[i], j), idxs, iotas)


def expand_reduce(sz: Callable, get: Callable, op: Callable, ne, arr: Sequence) -> list:
    """segmented.fut:84-91."""
    szs = S.fmap(sz, arr)
    idxs = replicated_iota(szs)
    flags = S.map2(lambda a, b: a != b, idxs, S.rotate(-1, idxs))
    iotas = segmented_iota(flags)
    vs = S.map2(lambda i, j: get(arr[i], j), idxs, iotas)
    return segmented_reduce(op, ne, flags, vs)


def expand_outer_reduce(sz: Callable, get: Callable, op: Callable, ne, arr: Sequence) -> list:
    """segmented.fut:97-103."""
    def sz2(x):
        s = sz(x)
        return 1 if s == 0 else s
    get2 = lambda x, i: ne if sz(x) == 0 else get(x, i)
    out = expand_reduce(sz2, get2, op, ne, arr)
    assert len(out) == len(arr), "size coercion :> [n]b would fail at run time"
    return out


# --------------------------------------------------------------------------------------
# futhark/select.fut
# --------------------------------------------------------------------------------------

def sel(cols: Sequence[int], row: Sequence[int]) -> List[int]:
    """select.fut:9-11 — gather the listed column indices out of one row."""
    def f(i):
        if not 0 <= i < len(row):
            raise IndexError(f"Index [{i}] out of bounds for array of shape [{len(row)}]")
        return row[i]
    return S.fmap(f, cols)


def sel_all(db: Sequence[Sequence[int]], cols: Sequence[int]) -> List[List[int]]:
    """select.fut:17-20 — the filter at :18 is a comment; only the projection runs."""
    return S.fmap(lambda row: sel(cols, row), db)


def query_sel(db, cols):
    """main.fut:7 -> select.fut:23.  i32 in, i32 out, row order preserved."""
    return sel_all(db, cols)


# --------------------------------------------------------------------------------------
# futhark/groupby.fut
# --------------------------------------------------------------------------------------

def _split_indices(keys: Sequence[int], bitn: int) -> List[int]:
    """groupby.fut:10-17 / join.fut:11-18 — destination index of a stable split on one key bit."""
    bits1 = S.fmap(lambda x: S.i32(S.u32(x) >> bitn) & 1, keys)        # :10
    bits0 = S.fmap(lambda b: 1 - b, bits1)                             # :11
    add = lambda a, b: S.i32(a + b)
    idxs0 = S.map2(lambda a, b: a * b, bits0, S.scan(add, 0, bits0))   # :12
    idxs1 = S.scan(add, 0, bits1)                                      # :13
    offs = S.reduce(add, 0, bits0)                                     # :14
    idxs1 = S.map2(lambda a, b: a * b, bits1, S.fmap(lambda x: x + offs, idxs1))   # :15
    idxs = S.map2(lambda a, b: a + b, idxs0, idxs1)                    # :16
    return S.fmap(lambda x: x - 1, idxs)                               # :17


def rsort_step_rows(xs: List[List[int]], bitn: int) -> List[List[int]]:
    """groupby.fut:8-18 — one 1-bit pass over whole rows keyed on column 0."""
    idxs = _split_indices([row[0] for row in xs], bitn)
    return S.scatter([list(r) for r in xs], idxs, xs)                  # :18


def rsort_rows(xs: List[List[int]]) -> List[List[int]]:
    """groupby.fut:21-22 — 32 passes, bit 0 first: ascending unsigned, stable."""
    for i in range(32):
        xs = rsort_step_rows(xs, i)
    return xs


def mk_flags(row_ids: Sequence[int]) -> List[int]:
    """groupby.fut:26-33."""
    return [1 if i == 0 else (1 if row_ids[i - 1] != row_ids[i] else 0)
            for i in range(len(row_ids))]


def type_func(typ: int, v1: int, v2: int) -> int:
    """groupby.fut:35-41 — 1 prod, 2 sum, 3 max, 4 min, anything else min; all u32."""
    if typ == 1:
        return S.u32(v1 * v2)
    if typ == 2:
        return S.u32(v1 + v2)
    if typ == 3:
        return max(v1, v2)
    return min(v1, v2)


def merge(s_cols_t: Sequence[int], a: Sequence[int], b: Sequence[int]) -> List[int]:
    """groupby.fut:45-48 — column 0 keeps the left operand, column i uses op t[i-1]."""
    return [a[i] if i == 0 else type_func(s_cols_t[i - 1], a[i], b[i]) for i in range(len(a))]


def groupby(db: Sequence[Sequence[int]], cols: Sequence[int], t_cols: Sequence[int]) -> List[List[int]]:
    """groupby.fut:51-58."""
    s = len(cols)
    if len(t_cols) < s - 1:
        raise IndexError("t_cols shorter than the aggregated column list (groupby.fut:47)")
    keep = [[S.u32(row[i]) for i in cols] for row in db]               # :52-53
    sorted_rows = rsort_rows(keep)                                     # :54
    idxs = mk_flags([r[0] for r in sorted_rows])                       # :55
    flag = [v == 1 for v in idxs]                                      # :56
    helper = lambda a, b: merge(t_cols, a, b)                          # :57
    return segmented_reduce(helper, S.replicate(s, 0), flag, sorted_rows)   # :58


def query_groupby(db, g_col: int, s_cols: Sequence[int], t_cols: Sequence[int]):
    """main.fut:9 -> groupby.fut:60-62."""
    cols = S.concat([g_col], list(s_cols))
    m = len(db[0]) if len(db) else None
    if m is not None:
        for c in cols:
            if not 0 <= c < m:
                raise IndexError(f"Index [{c}] out of bounds for array of shape [{m}]")
    return groupby(db, cols, t_cols)


# --------------------------------------------------------------------------------------
# futhark/join.fut
# --------------------------------------------------------------------------------------

Triple = Tuple[int, int, int]


def rsort_step_triples(xs: List[Triple], bitn: int) -> List[Triple]:
    """join.fut:9-19."""
    idxs = _split_indices([t[0] for t in xs], bitn)
    return S.scatter(list(xs), idxs, xs)


def rsort_triples(xs: List[Triple]) -> List[Triple]:
    """join.fut:22-23."""
    for i in range(32):
        xs = rsort_step_triples(xs, i)
    return xs


def mk_flags_col(row_ids: Sequence[Triple]) -> List[int]:
    """join.fut:27-34."""
    return [1 if i == 0 else (1 if row_ids[i - 1][0] != row_ids[i][0] else 0)
            for i in range(len(row_ids))]


def generate_pairs(arr: Sequence[Tuple[int, int]]) -> List[Tuple[int, int]]:
    """join.fut:37-41 — tag 1 rows crossed with the rest, left-major."""
    t_arr1, t_arr2 = S.partition(lambda x: x[0] == 1, arr)             # :38
    arr1 = [x[1] for x in t_arr1]
    arr2 = [x[1] for x in t_arr2]
    return expand(lambda _: len(arr2), lambda x, i: (x, arr2[i]), arr1)   # :41


def dim_helper(flags: Sequence[int]) -> List[int]:
    """join.fut:43 — segment lengths."""
    return segmented_reduce(lambda a, b: a + b, 0, [f == 1 for f in flags],
                            S.replicate(len(flags), 1))


def join(db1, db2, col1: int, col2: int, cols1: Sequence[int], cols2: Sequence[int]):
    """join.fut:52-75 (orphan entry: main.fut does not import it)."""
    n, s = len(db1), len(db2)
    l1 = [(S.u32(db1[i][col1]), 1, i) for i in range(n)]               # :55
    l2 = [(S.u32(db2[i][col2]), 2, i) for i in range(s)]               # :56
    to_sort = S.concat(l1, l2)                                         # :57
    srt = rsort_triples(to_sort)                                       # :58
    flags = mk_flags_col(srt)                                          # :59
    f_lens = dim_helper(flags)                                         # :60
    f_l = S.scan(lambda a, b: a + b, 0, f_lens)                        # :61
    p_lens = S.scatter(S.rotate(-1, list(f_l)), [0], [0])              # :63
    inds = list(zip(p_lens, f_l))                                      # :64
    sorted_copy = [(t[1], t[2]) for t in srt]                          # :65-66
    pairs: List[Tuple[int, int]] = []
    for (st, fn) in inds:                                              # :67-68
        pairs = S.concat(pairs, generate_pairs(sorted_copy[st:fn]))
    t1s = [sel(cols1, [S.u32(v) for v in db1[r1]]) for (r1, _) in pairs]   # :69,72
    t2s = [sel(cols2, [S.u32(v) for v in db2[r2]]) for (_, r2) in pairs]   # :70,73
    return [a + b for a, b in zip(t1s, t2s)]                           # :74-75
