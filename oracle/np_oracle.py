"""numpy twin of the CPU oracle — TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline leg may import this.
Nothing under ``harkdb_b200/`` imports it; the product path is libhark.so (CUDA) only.

Two groups of functions:

* reference-pinned semantics (vectorised restatements, checked against the line-by-line
  simulation in ``oracle/hark_ref.py`` and the C port in ``oracle/oracle.c``):
    query_sel      select.fut:9-23  / main.fut:7
    query_groupby  groupby.fut:51-62 / main.fut:9   (u32, ascending unsigned keys, ops 0-4)
    join           join.fut:52-75                   (key asc unsigned, then r1, then r2)
  PARITY UNPINNED by reference tests (the reference has no expected outputs for them).

* extensions the reference does not implement (WHERE, typed GROUP BY with COUNT/AVG/HAVING,
  ORDER BY, join+GROUP BY, the synthetic generator).  This module DEFINES their semantics
  (DESIGN.md §extensions); PARITY UNPINNED, oracle-defined.

Tables are lists of 1-D numpy column arrays (SoA), one dtype per column.
"""

from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

I32, U32, I64, F32, F64 = 0, 1, 2, 3, 4
NP_DTYPES = {I32: np.int32, U32: np.uint32, I64: np.int64, F32: np.float32, F64: np.float64}
DTYPE_CODES = {np.dtype(v): k for k, v in NP_DTYPES.items()}

GT, GE, LT, LE, EQ, NE = 0, 1, 2, 3, 4, 5
AGG_KEY, AGG_PROD, AGG_SUM, AGG_MAX, AGG_MIN, AGG_COUNT, AGG_AVG, AGG_SUMF64, AGG_SUM64 = 0, 1, 2, 3, 4, 5, 6, 7, 8
GEN_UNIFORM, GEN_AFFINE, GEN_CONST, GEN_LOGUNIFORM, GEN_AFFINE_UNIFORM = 0, 1, 2, 3, 4

Pred = Tuple[int, int, int, float]  # (col, op, ival, fval) — include/hark.h hark_pred


def dtype_code(a: np.ndarray) -> int:
    return DTYPE_CODES[np.dtype(a.dtype)]


# --------------------------------------------------------------------------------------------
# generator (include/hark.h hark_colspec; oracle.c oracle_synth_column is the C twin)
# --------------------------------------------------------------------------------------------

_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def mix64(seed: int, col: int, rows: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser over state (seed ^ (col+1)*C) + (row+1)*golden."""
    with np.errstate(over="ignore"):
        base = np.uint64((seed ^ (((col + 1) * 0xD6E8FEB86659FD93) & 0xFFFFFFFFFFFFFFFF)) & 0xFFFFFFFFFFFFFFFF)
        z = base + (rows.astype(np.uint64) + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def _mulhi64(a: np.ndarray, b: int) -> np.ndarray:
    a_lo = a & np.uint64(0xFFFFFFFF)
    a_hi = a >> np.uint64(32)
    b_lo = np.uint64(b & 0xFFFFFFFF)
    b_hi = np.uint64(b >> 32)
    with np.errstate(over="ignore"):
        ll = a_lo * b_lo
        lh = a_lo * b_hi
        hl = a_hi * b_lo
        hh = a_hi * b_hi
        mid = (ll >> np.uint64(32)) + (lh & np.uint64(0xFFFFFFFF)) + (hl & np.uint64(0xFFFFFFFF))
        return hh + (lh >> np.uint64(32)) + (hl >> np.uint64(32)) + (mid >> np.uint64(32))


def synth_column(dtype: int, spec: dict, seed: int, col: int, row0: int, n: int) -> np.ndarray:
    """spec keys: kind, lo, range, flo, fhi, a, b (missing keys default to 0 / [0,1))."""
    kind = spec.get("kind", GEN_UNIFORM)
    lo = int(spec.get("lo", 0))
    rng = int(spec.get("range", 0))
    flo = float(spec.get("flo", 0.0))
    fhi = float(spec.get("fhi", 1.0))
    a = int(spec.get("a", 1))
    b = int(spec.get("b", 0))
    rows = np.arange(row0, row0 + n, dtype=np.uint64)
    npdt = NP_DTYPES[dtype]
    with np.errstate(over="ignore"):
        if kind == GEN_UNIFORM:
            h = mix64(seed, col, rows)
            if dtype == F32:
                u = (h >> np.uint64(40)).astype(np.float32) * np.float32(2.0 ** -24)
                # fmaf(u, fhi-flo, flo): exact in float64 then one rounding == fused
                scale = np.float32(fhi - flo)
                return (u.astype(np.float64) * np.float64(scale) + np.float64(np.float32(flo))).astype(np.float32)
            if dtype == F64:
                u = (h >> np.uint64(11)).astype(np.float64) * (2.0 ** -53)
                if flo == 0.0 and fhi == 1.0:
                    return u
                return _fma64(u, fhi - flo, flo)
            v = np.uint64(lo & 0xFFFFFFFFFFFFFFFF) + (_mulhi64(h, rng) if rng else h)
        elif kind == GEN_AFFINE:
            v = np.uint64(a) * rows + np.uint64(b)
            if rng:
                v = v % np.uint64(rng)
            if dtype in (F32, F64):
                return v.astype(npdt)
        elif kind == GEN_AFFINE_UNIFORM:
            v = np.uint64(a) * _mulhi64(mix64(seed, col, rows), rng) + np.uint64(b)
            if dtype in (F32, F64):
                return v.astype(npdt)
        elif kind == GEN_LOGUNIFORM:
            r2 = max(rng, 2)
            nb = r2.bit_length() - 1
            e = _mulhi64(mix64(seed, col, rows), nb)
            h2 = mix64(seed ^ 0x5851F42D4C957F2D, col, rows)
            one = np.uint64(1)
            k = ((one << e) + (h2 & ((one << e) - one)) - one) % np.uint64(r2)
            v = np.uint64(lo & 0xFFFFFFFFFFFFFFFF) + k
            if dtype in (F32, F64):
                return v.view(np.int64).astype(npdt)
        else:
            if dtype in (F32, F64):
                return np.full(n, flo, dtype=npdt)
            v = np.full(n, lo & 0xFFFFFFFFFFFFFFFF, dtype=np.uint64)
    if dtype in (I32, U32):
        return (v & np.uint64(0xFFFFFFFF)).astype(np.uint32).view(npdt)
    return v.view(np.int64)


def _fma64(u: np.ndarray, s: float, c: float) -> np.ndarray:
    import math
    if hasattr(math, "fma"):
        return np.array([math.fma(float(x), s, c) for x in u], dtype=np.float64)
    from fractions import Fraction
    return np.array([float(Fraction(float(x)) * Fraction(s) + Fraction(c)) for x in u], dtype=np.float64)


# --------------------------------------------------------------------------------------------
# reference-pinned semantics
# --------------------------------------------------------------------------------------------

def query_sel(db: np.ndarray, cols: Sequence[int]) -> np.ndarray:
    """select.fut:9-23: out[r][j] = db[r][cols[j]]; IndexError stands for Futhark's bounds error."""
    db = np.asarray(db)
    n, m = db.shape
    cols = [int(c) for c in cols]
    if n > 0:
        for c in cols:
            if not 0 <= c < m:
                raise IndexError(f"Index [{c}] out of bounds for array of shape [{m}]")
    if len(cols) == 0 or n == 0:
        return np.zeros((n, len(cols)), dtype=db.dtype)
    return np.ascontiguousarray(db[:, cols])


def _fold_u32(op: int, vals: np.ndarray, starts: np.ndarray) -> np.ndarray:
    """Per-segment fold of u32 values with groupby.fut:35-41's operator for code `op`."""
    v = vals.astype(np.uint32)
    if op == AGG_PROD:
        # u32 product mod 2^32, folded exactly with Python ints per segment
        ends = np.append(starts[1:], len(v))
        out = np.empty(len(starts), dtype=np.uint32)
        for g, (s, e) in enumerate(zip(starts, ends)):
            acc = 1
            for x in v[s:e].tolist():
                acc = (acc * x) & 0xFFFFFFFF
            out[g] = acc
        return out
    if op == AGG_SUM:
        return np.add.reduceat(v.astype(np.uint64), starts).astype(np.uint64).astype(np.uint32) \
            if len(starts) else np.zeros(0, np.uint32)
    if op == AGG_MAX:
        return np.maximum.reduceat(v, starts) if len(starts) else np.zeros(0, np.uint32)
    return np.minimum.reduceat(v, starts) if len(starts) else np.zeros(0, np.uint32)


def query_groupby(db: np.ndarray, g_col: int, s_cols: Sequence[int], t_cols: Sequence[int]) -> np.ndarray:
    """groupby.fut:51-62 on a 2-D table viewed as u32: [key, agg_1..agg_c], keys ascending unsigned."""
    db = np.asarray(db)
    n, m = db.shape
    c = len(s_cols)
    if len(t_cols) < c:
        raise IndexError("t_cols shorter than s_cols (groupby.fut:47)")
    if n == 0:
        return np.zeros((0, c + 1), dtype=np.uint32)
    for col in [g_col, *s_cols]:
        if not 0 <= col < m:
            raise IndexError(f"Index [{col}] out of bounds for array of shape [{m}]")
    u = db.astype(np.int64).astype(np.uint32) if db.dtype != np.uint32 else db
    keys = u[:, g_col]
    order = np.argsort(keys, kind="stable")
    sk = keys[order]
    starts = np.flatnonzero(np.concatenate(([True], sk[1:] != sk[:-1])))
    out = np.empty((len(starts), c + 1), dtype=np.uint32)
    out[:, 0] = sk[starts]
    for j, (col, op) in enumerate(zip(s_cols, t_cols)):
        out[:, j + 1] = _fold_u32(int(op), u[order, col], starts)
    return out


def join(db1: np.ndarray, db2: np.ndarray, col1: int, col2: int, cols1: Sequence[int],
         cols2: Sequence[int]) -> np.ndarray:
    """join.fut:52-75: inner equi-join, rows ordered (key asc unsigned, r1 asc, r2 asc), u32 output."""
    db1 = np.asarray(db1)
    db2 = np.asarray(db2)
    l, k = len(cols1), len(cols2)
    u1 = db1.astype(np.int64).astype(np.uint32) if db1.dtype != np.uint32 else db1
    u2 = db2.astype(np.int64).astype(np.uint32) if db2.dtype != np.uint32 else db2
    if db1.shape[0] == 0 or db2.shape[0] == 0:
        return np.zeros((0, l + k), dtype=np.uint32)
    k1, k2 = u1[:, col1], u2[:, col2]
    o1 = np.argsort(k1, kind="stable")
    o2 = np.argsort(k2, kind="stable")
    s1, s2 = k1[o1], k2[o2]
    lb = np.searchsorted(s2, s1, side="left")
    ub = np.searchsorted(s2, s1, side="right")
    cnt = ub - lb
    total = int(cnt.sum())
    if total == 0:
        return np.zeros((0, l + k), dtype=np.uint32)
    left_pos = np.repeat(np.arange(len(s1)), cnt)
    offs = np.cumsum(cnt) - cnt
    within = np.arange(total) - np.repeat(offs, cnt)
    r1 = o1[left_pos]
    r2 = o2[np.repeat(lb, cnt) + within]
    out = np.empty((total, l + k), dtype=np.uint32)
    for j, c in enumerate(cols1):
        out[:, j] = u1[r1, c]
    for j, c in enumerate(cols2):
        out[:, l + j] = u2[r2, c]
    return out


# --------------------------------------------------------------------------------------------
# extensions (oracle-defined)
# --------------------------------------------------------------------------------------------

def _cmp(a: np.ndarray, op: int, ival: int, fval: float) -> np.ndarray:
    code = dtype_code(a)
    if code in (I32, U32, I64):
        x = a.astype(np.int64)
        c = np.int64(ival)
    elif code == F32:
        x = a
        c = np.float32(fval)
    else:
        x = a
        c = np.float64(fval)
    with np.errstate(invalid="ignore"):
        if op == GT:
            return x > c
        if op == GE:
            return x >= c
        if op == LT:
            return x < c
        if op == LE:
            return x <= c
        if op == EQ:
            return x == c
        return x != c


def query_filter(columns: Sequence[np.ndarray], cols: Sequence[int], preds: Sequence[Pred]) -> List[np.ndarray]:
    """SELECT cols WHERE <preds in conjunctive normal form>; input row order kept.  A predicate whose op carries
    PRED_OR is OR-ed with the next one; the row passes when every such clause holds; PRED_NOT negates the comparison
    (include/hark.h).  Without flags this is p1 AND p2 AND ..."""
    return [np.ascontiguousarray(columns[c][pred_mask(columns, preds)]) for c in cols]


PRED_OP_MASK, PRED_OR, PRED_NOT = 0xFF, 0x100, 0x200


def pred_mask(columns: Sequence[np.ndarray], preds: Sequence[Pred]) -> np.ndarray:
    n = len(columns[0]) if columns else 0
    mask = np.ones(n, dtype=bool)
    clause = np.zeros(n, dtype=bool)
    for (c, op, ival, fval) in preds:
        m = _cmp(columns[c], op & PRED_OP_MASK, ival, fval)
        if op & PRED_NOT:
            m = ~m
        clause |= m
        if op & PRED_OR:
            continue
        mask &= clause
        clause = np.zeros(n, dtype=bool)
    return mask


def _agg_typed(op: int, v: np.ndarray, starts: np.ndarray, counts: np.ndarray) -> np.ndarray:
    code = dtype_code(v)
    G = len(starts)
    if op == AGG_COUNT:
        return counts.astype(np.int64)
    if op == AGG_SUM64 and code in (I32, U32, I64):      # exact integer sum (wraps only at 2^64)
        with np.errstate(over="ignore"):
            return np.add.reduceat(v.astype(np.int64), starts) if G else np.zeros(0, np.int64)
    if op == AGG_AVG or op == AGG_SUMF64 or op == AGG_SUM64:
        s = np.add.reduceat(v.astype(np.float64), starts) if G else np.zeros(0, np.float64)
        return s / counts.astype(np.float64) if op == AGG_AVG else s
    if G == 0:
        return np.zeros(0, dtype=v.dtype)
    if op == AGG_SUM:
        if code in (F32, F64):
            return np.add.reduceat(v.astype(np.float64), starts).astype(v.dtype)
        with np.errstate(over="ignore"):
            if code in (I32, U32):
                return np.add.reduceat(v.astype(np.uint32).astype(np.uint64), starts).astype(np.uint32).view(v.dtype)
            return np.add.reduceat(v.view(np.uint64), starts).view(np.int64)
    if op == AGG_PROD:
        ends = np.append(starts[1:], len(v))
        out = np.empty(G, dtype=v.dtype)
        if code in (F32, F64):
            for g, (s, e) in enumerate(zip(starts, ends)):
                out[g] = np.prod(v[s:e].astype(np.float64))
            return out
        bits = 32 if code in (I32, U32) else 64
        mask = (1 << bits) - 1
        uo = out.view(np.uint32 if bits == 32 else np.uint64)
        uv = v.view(np.uint32 if bits == 32 else np.uint64)
        for g, (s, e) in enumerate(zip(starts, ends)):
            acc = 1
            for x in uv[s:e].tolist():
                acc = (acc * x) & mask
            uo[g] = acc
        return out
    if op == AGG_MAX:
        return np.maximum.reduceat(v, starts)
    return np.minimum.reduceat(v, starts)   # AGG_MIN, AGG_KEY (0) and unknown codes: groupby.fut:41


def agg_out_dtype(op: int, code: int) -> int:
    if op == AGG_COUNT:
        return I64
    if op == AGG_SUM64:
        return I64 if code in (I32, U32, I64) else F64
    if op == AGG_AVG or op == AGG_SUMF64:
        return F64
    return code


def query_groupby_ex(columns: Sequence[np.ndarray], g_col: int, s_cols: Sequence[int], ops: Sequence[int],
                     having: Sequence[Pred] = ()) -> List[np.ndarray]:
    """Typed GROUP BY: output [key, agg_1..agg_c]; keys ascending in the key dtype's own order."""
    key = columns[g_col]
    if dtype_code(key) not in (I32, U32, I64):
        raise ValueError("group key must be an integer column")
    order = np.argsort(key, kind="stable")
    sk = key[order]
    n = len(sk)
    starts = np.flatnonzero(np.concatenate(([True], sk[1:] != sk[:-1]))) if n else np.zeros(0, np.int64)
    counts = np.diff(np.append(starts, n)) if n else np.zeros(0, np.int64)
    out = [np.ascontiguousarray(sk[starts])]
    for col, op in zip(s_cols, ops):
        out.append(_agg_typed(int(op), columns[col][order], starts, counts))
    if having:
        mask = pred_mask(out, having)
        out = [np.ascontiguousarray(o[mask]) for o in out]
    return out


def query_groupby_multi(columns: Sequence[np.ndarray], g_cols: Sequence[int], s_cols: Sequence[int], ops: Sequence[int],
                        having: Sequence[Pred] = ()) -> List[np.ndarray]:
    """GROUP BY several integer key columns (the extension the reference wishes for at parse.py:64): output =
    [key_1..key_ng, agg_1..agg_c], rows ascending lexicographically by the keys, each in its own dtype's order;
    aggregates exactly as query_groupby_ex; HAVING indexes the output columns."""
    keys = [columns[g] for g in g_cols]
    for k in keys:
        if dtype_code(k) not in (I32, U32, I64):
            raise ValueError("group keys must be integer columns")
    n = len(keys[0]) if keys else 0
    order = np.lexsort(tuple(reversed([order_key(k) for k in keys]))) if n else np.zeros(0, np.int64)
    sk = [k[order] for k in keys]
    if n:
        head = np.zeros(n, dtype=bool)
        head[0] = True
        for k in sk:
            head[1:] |= k[1:] != k[:-1]
        starts = np.flatnonzero(head)
        counts = np.diff(np.append(starts, n))
    else:
        starts = counts = np.zeros(0, np.int64)
    out = [np.ascontiguousarray(k[starts]) for k in sk]
    for col, op in zip(s_cols, ops):
        out.append(_agg_typed(int(op), columns[col][order], starts, counts))
    if having:
        mask = pred_mask(out, having)
        out = [np.ascontiguousarray(o[mask]) for o in out]
    return out


def order_key(a: np.ndarray, desc: bool = False) -> np.ndarray:
    """Order-preserving map to unsigned ints: signed order for ints, IEEE order with NaN last for floats."""
    code = dtype_code(a)
    if code == I32:
        u = a.view(np.uint32) ^ np.uint32(0x80000000)
    elif code == U32:
        u = a.copy()
    elif code == I64:
        u = a.view(np.uint64) ^ np.uint64(0x8000000000000000)
    elif code == F32:
        b = a.view(np.uint32)
        u = np.where(b >> np.uint32(31), ~b, b | np.uint32(0x80000000))
        u = np.where(np.isnan(a), np.uint32(0xFFFFFFFF), u).astype(np.uint32)
    else:
        b = a.view(np.uint64)
        u = np.where(b >> np.uint64(63), ~b, b | np.uint64(0x8000000000000000))
        u = np.where(np.isnan(a), np.uint64(0xFFFFFFFFFFFFFFFF), u).astype(np.uint64)
    return ~u if desc else u


def query_orderby(columns: Sequence[np.ndarray], cols: Sequence[int], key_cols: Sequence[int],
                  desc: Optional[Sequence[int]] = None) -> List[np.ndarray]:
    """SELECT cols ORDER BY key_cols (lexicographic, per-key ASC/DESC), stable w.r.t. input order."""
    desc = list(desc) if desc is not None else [0] * len(key_cols)
    n = len(columns[0]) if columns else 0
    if len(key_cols) == 0:
        perm = np.arange(n)
    else:
        keys = [order_key(columns[c], bool(d)) for c, d in zip(key_cols, desc)]
        perm = np.lexsort(tuple(reversed(keys)))   # lexsort: last key is primary; it is stable
    return [np.ascontiguousarray(columns[c][perm]) for c in cols]


def join_ex(t1: Sequence[np.ndarray], t2: Sequence[np.ndarray], col1: int, col2: int, cols1: Sequence[int],
            cols2: Sequence[int]) -> List[np.ndarray]:
    """Typed join (extension of join.fut:52-75): key columns of one integer dtype compared in its own order, output
    columns keep their dtypes, rows ordered (key asc, r1 asc, r2 asc).  A multiset (order = 0) result is compared
    after sorting both sides' rows."""
    k1, k2 = np.asarray(t1[col1]), np.asarray(t2[col2])
    o1 = np.argsort(k1, kind="stable")
    o2 = np.argsort(k2, kind="stable")
    s1, s2 = k1[o1], k2[o2]
    lb = np.searchsorted(s2, s1, side="left")
    ub = np.searchsorted(s2, s1, side="right")
    cnt = ub - lb
    total = int(cnt.sum())
    left_pos = np.repeat(np.arange(len(s1)), cnt)
    offs = np.cumsum(cnt) - cnt
    within = np.arange(total) - np.repeat(offs, cnt)
    r1 = o1[left_pos]
    r2 = o2[np.repeat(lb, cnt) + within] if total else np.zeros(0, dtype=np.int64)
    return [np.asarray(t1[c])[r1] for c in cols1] + [np.asarray(t2[c])[r2] for c in cols2]


def join_groupby(fact: Sequence[np.ndarray], dim: Sequence[np.ndarray], fk_col: int, pk_col: int, g_col: int,
                 s_cols: Sequence[int], ops: Sequence[int]) -> List[np.ndarray]:
    """SELECT d.g, agg(f.s) FROM fact f JOIN dim d ON f.fk = d.pk GROUP BY d.g  (d.pk unique)."""
    pk = dim[pk_col]
    o = np.argsort(pk, kind="stable")
    spk = pk[o]
    if len(spk) > 1 and np.any(spk[1:] == spk[:-1]):
        raise ValueError("dim.pk must be unique")
    fk = fact[fk_col]
    pos = np.searchsorted(spk, fk)
    pos_c = np.minimum(pos, max(len(spk) - 1, 0))
    hit = (pos < len(spk)) & (spk[pos_c] == fk) if len(spk) else np.zeros(len(fk), bool)
    g = dim[g_col][o][pos_c[hit]] if len(spk) else dim[g_col][:0]
    cols = [g] + [fact[c][hit] for c in s_cols]
    return query_groupby_ex(cols, 0, list(range(1, len(cols))), ops)
