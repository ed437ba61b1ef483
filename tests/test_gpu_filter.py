"""K1 parity (-m gpu): tables, generator, projection and filter-compact through the C-ABI vs the oracle.
Bit-exact: everything here is byte/integer/index work or IEEE compares."""

import json
import os

import numpy as np
import pytest

from oracle import c_oracle as CO
from oracle import np_oracle as NO
from tests.gpu_util import cols_of, free_gb, get_env, rand_table

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
ALL_DT = [NO.I32, NO.U32, NO.I64, NO.F32, NO.F64]


@pytest.mark.parametrize("dtype", ALL_DT)
@pytest.mark.parametrize("shape", [(0, 3), (1, 1), (7, 8), (1000, 5), (3000, 70), (100003, 3), (5, 0)])
def test_upload_download_roundtrip(dtype, shape):
    env = get_env()
    rng = np.random.default_rng(1)
    a = rand_table(rng, shape[0], shape[1], dtype, lo=-50 if dtype != NO.U32 else 0, hi=2 ** 31 - 1)
    t = env.to_device(a)
    assert t.shape == shape and t.dtypes == [dtype] * shape[1]
    back = t.to_numpy()
    assert (back.dtype == a.dtype or shape[1] == 0) and np.array_equal(back, a)
    for c in range(min(shape[1], 3)):
        assert np.array_equal(t.column(c), a[:, c])
    t.free()


def test_upload_chunked_pipeline():
    env = get_env()
    env.set_option("upload.chunk_mb", 1)      # force many double-buffered chunks
    try:
        rng = np.random.default_rng(2)
        a = rng.integers(-2 ** 31, 2 ** 31, (700001, 8)).astype(np.int32)
        t = env.to_device(a)
        assert np.array_equal(t.to_numpy(), a)
        t.free()
    finally:
        env.set_option("upload.chunk_mb", 64)


def test_from_columns_mixed_dtypes_and_slice_concat():
    env = get_env()
    rng = np.random.default_rng(3)
    cols = [rng.integers(0, 9, 1001).astype(np.int32), rng.random(1001).astype(np.float64),
            rng.random(1001).astype(np.float32), rng.integers(-9, 9, 1001).astype(np.int64)]
    t = env.from_columns(cols)
    assert t.dtypes == [NO.I32, NO.F64, NO.F32, NO.I64]
    for c, col in enumerate(cols):
        assert np.array_equal(t.column(c), col)
    assert np.array_equal(t.column(1, 10, 50), cols[1][10:60])
    s = env.slice(t, 333, 501)      # odd offset: exercises the unaligned copy path
    for c, col in enumerate(cols):
        assert np.array_equal(s.column(c), col[333:834])
    cc = env.concat(s, t)
    for c, col in enumerate(cols):
        assert np.array_equal(cc.column(c), np.concatenate([col[333:834], col]))
    for x in (t, s, cc):
        x.free()


@pytest.mark.parametrize("dtype", ALL_DT)
def test_generator_matches_oracle(dtype):
    env = get_env()
    specs = [dict(kind=NO.GEN_UNIFORM, lo=-5, range=11), dict(kind=NO.GEN_UNIFORM, lo=0, range=0, flo=-2.0, fhi=3.0),
             dict(kind=NO.GEN_UNIFORM, lo=-(2 ** 19), range=2 ** 20), dict(kind=NO.GEN_AFFINE, a=7, b=3, range=1000),
             dict(kind=NO.GEN_CONST, lo=42, flo=4.25), dict(kind=NO.GEN_LOGUNIFORM, lo=-3, range=1 << 20),
             dict(kind=NO.GEN_LOGUNIFORM, lo=5, range=1000),
             dict(kind=NO.GEN_AFFINE_UNIFORM, a=2654435761, b=977, range=100003)]
    n, row0 = 70001, 10 ** 9 - 5
    t = env.synth(n, [dtype] * len(specs), specs, seed=42, row0=row0)
    for c, spec in enumerate(specs):
        exp = CO.synth_column(dtype, spec, 42, c, row0, n)
        assert np.array_equal(t.column(c), exp), (dtype, spec)
    t.free()


def test_query_sel_golden_and_errors():
    env = get_env()
    from harkdb_b200.hark_ffi import HarkError
    data = json.load(open(os.path.join(GOLDEN, "data_csv.json")))
    db = np.asarray(data["rows"], dtype=np.int64)
    res = env.from_futhark(env.query_sel(db, np.array([0, 2])))      # BASELINE config 1 / README.md:42
    assert res.dtype == np.int32
    assert res.tolist() == [[6, 6], [0, 0], [0, 0], [0, 0], [0, 0], [6, 6], [1, 3]]
    for case in json.load(open(os.path.join(GOLDEN, "harkdb_vectors.json")))["cases"]:
        if case["kind"] != "query_sel":
            continue
        rows = data["rows"] if case["db"] == "data_csv" else case["db"]
        a = np.asarray(rows, dtype=np.int64).astype(np.uint32).view(np.int32).reshape(len(rows), -1)
        if "error" in case:
            with pytest.raises(HarkError):
                env.query_sel(a, case["cols"])
            continue
        got = env.from_futhark(env.query_sel(a, case["cols"]))
        exp = np.asarray(case["output"], dtype=np.int64).astype(np.uint32).view(np.int32).reshape(a.shape[0], len(case["cols"]))
        assert np.array_equal(got, exp)


@pytest.mark.parametrize("dtype", ALL_DT)
@pytest.mark.parametrize("n", [0, 1, 1000, 1 << 20])
def test_query_sel_vs_oracle(dtype, n):
    env = get_env()
    rng = np.random.default_rng(n + dtype)
    a = rand_table(rng, n, 6, dtype)
    t = env.to_device(a)
    for cols in ([0], [5, 5, 1], [], [3, 2, 1, 0, 4, 5, 0, 1, 2, 3, 4, 5, 0, 1, 2, 3, 4, 5]):
        r = env.query_sel(t, cols)
        assert np.array_equal(r.to_numpy(), NO.query_sel(a, cols))
        r.free()
    t.free()


SIZES = [0, 1, 15, 16, 17, 511, 512, 513, 1023, 1024, 1025, 4095, 4096, 4097, 32767, 32768, 32769, 100000,
         (1 << 20) + 3]


@pytest.mark.parametrize("n", SIZES)
def test_filter_f32_config2_shape(n):
    """BASELINE config 2 at oracle-checkable sizes: 8 f32 columns, SELECT col1,col3 WHERE col2>t AND col5<u."""
    env = get_env()
    specs = [dict(kind=NO.GEN_UNIFORM)] * 8
    t = env.synth(n, [NO.F32] * 8, specs, seed=42)
    cols = [CO.synth_column(NO.F32, specs[c], 42, c, 0, n) for c in range(8)]
    for (tt, uu, impl) in [(0.5, 0.5, 0), (0.99, 0.01, 0), (-1.0, 2.0, 0), (2.0, 0.5, 0), (0.5, 0.5, 1), (0.5, 0.5, 3),
                           (0.999, 0.5, 3)]:
        env.set_option("filter.impl", impl)
        preds = [(1, NO.GT, 0, tt), (4, NO.LT, 0, uu)]
        r = env.query_filter(t, [0, 2], preds)
        exp = NO.query_filter(cols, [0, 2], preds)
        assert r.shape == (len(exp[0]), 2)
        assert np.array_equal(r.column(0), exp[0]) and np.array_equal(r.column(1), exp[1])
        st = env.stats()
        assert st["rows_in"] == n and st["rows_out"] == len(exp[0])
        if n:
            assert st["alg_bytes"] == 16 * n + 8 * len(exp[0])      # DESIGN.md roofline numerator
        r.free()
    env.set_option("filter.impl", 0)
    t.free()


@pytest.mark.parametrize("dtype", ALL_DT)
@pytest.mark.parametrize("op", [NO.GT, NO.GE, NO.LT, NO.LE, NO.EQ, NO.NE])
def test_filter_ops_all_dtypes(dtype, op):
    env = get_env()
    rng = np.random.default_rng(100 + dtype * 10 + op)
    n = 50021
    a = rand_table(rng, n, 4, dtype, lo=-20 if dtype not in (NO.U32,) else 0, hi=20, nan_frac=0.01)
    if dtype in (NO.F32, NO.F64):
        a[:, 1] = np.round(a[:, 1] * 8) / 8       # make equality hit
        preds = [(1, op, 0, 0.5)]
    else:
        preds = [(1, op, 3, 0.0)]
    t = env.to_device(a)
    for impl in (0, 1, 3):                        # v2 static counts, v1 per-tile kernel, v2 runtime counts
        env.set_option("filter.impl", impl)
        r = env.query_filter(t, [0, 1, 3], preds)
        exp = NO.query_filter(cols_of(a), [0, 1, 3], preds)
        for j in range(3):
            assert np.array_equal(r.column(j), exp[j], equal_nan=True), (dtype, op, impl, j)
        r.free()
    env.set_option("filter.impl", 0)
    t.free()


def test_filter_int_constant_outside_column_range():
    env = get_env()
    a = np.array([[-5, 1], [2 ** 31 - 1, 2], [0, 3]], dtype=np.int32)
    t = env.to_device(a)
    for preds in ([(0, NO.LT, 2 ** 40, 0.0)], [(0, NO.GT, -(2 ** 40), 0.0)], [(0, NO.GT, 2 ** 31 - 1, 0.0)]):
        r = env.query_filter(t, [1], preds)
        assert np.array_equal(r.column(0), NO.query_filter(cols_of(a), [1], preds)[0])
        r.free()
    u = np.array([[2 ** 32 - 1, 1], [5, 2]], dtype=np.uint32)
    tu = env.to_device(u)
    r = env.query_filter(tu, [1], [(0, NO.GT, 2 ** 31, 0.0)])       # unsigned column compares as a value, not bits
    assert r.column(0).tolist() == [1]
    r.free(); t.free(); tu.free()


@pytest.mark.parametrize("np_", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("ns", [0, 1, 2, 3, 4, 5, 17])
def test_filter_predicate_and_column_counts(np_, ns):
    env = get_env()
    rng = np.random.default_rng(np_ * 31 + ns)
    n = 30011
    a = rand_table(rng, n, 9, NO.I32, lo=0, hi=10)
    t = env.to_device(a)
    preds = [(int(rng.integers(0, 9)), int(rng.integers(0, 6)), int(rng.integers(1, 9)), 0.0) for _ in range(np_)]
    cols = [int(rng.integers(0, 9)) for _ in range(ns)]
    r = env.query_filter(t, cols, preds)
    exp = NO.query_filter(cols_of(a), cols, preds)
    mask_count = len(NO.query_filter(cols_of(a), [0], preds)[0])
    assert r.shape == (mask_count, ns)
    for j in range(ns):
        assert np.array_equal(r.column(j), exp[j])
    r.free(); t.free()


def test_filter_mixed_widths():
    env = get_env()
    rng = np.random.default_rng(9)
    n = 77777
    cols = [rng.integers(0, 50, n).astype(np.int32), rng.random(n).astype(np.float64),
            rng.random(n).astype(np.float32), rng.integers(-50, 50, n).astype(np.int64)]
    t = env.from_columns(cols)
    preds = [(0, NO.GE, 10, 0.0), (1, NO.LT, 0, 0.6), (3, NO.NE, 7, 0.0), (2, NO.GT, 0, 0.1)]
    r = env.query_filter(t, [3, 2, 1, 0], preds)
    exp = NO.query_filter(cols, [3, 2, 1, 0], preds)
    for j in range(4):
        assert np.array_equal(r.column(j), exp[j])
    r.free(); t.free()


def test_filter_errors():
    env = get_env()
    from harkdb_b200.hark_ffi import HarkError
    t = env.to_device(np.zeros((4, 2), dtype=np.int32))
    with pytest.raises(HarkError, match="out of bounds"):
        env.query_filter(t, [2], [(0, NO.GT, 0, 0.0)])
    with pytest.raises(HarkError, match="out of bounds"):
        env.query_filter(t, [0], [(5, NO.GT, 0, 0.0)])
    with pytest.raises(HarkError):
        env.query_filter(t, [0], [(0, 9, 0, 0.0)])
    with pytest.raises(HarkError):
        env.query_filter(t, [0], [(0, NO.GT, 0, 0.0)] * 17)
    t.free()


def test_filter_full_size_properties():
    """BASELINE config 2 at full size (1e9 x 8 f32 when memory allows): properties that do not need a
    full-size oracle — prefix/suffix slices against the regenerated rows, partition identity, idempotence."""
    env = get_env()
    n = 10 ** 9 if free_gb() > 60 else (1 << 27)
    specs = [dict(kind=NO.GEN_UNIFORM)] * 8
    t = env.synth(n, [NO.F32] * 8, specs, seed=42)
    tt, uu = 0.5, 0.5
    r = env.query_filter(t, [0, 2, 1, 4], [(1, NO.GT, 0, tt), (4, NO.LT, 0, uu)])
    cnt = r.shape[0]
    assert abs(cnt / n - 0.25) < 1e-3
    # prefix and suffix of the output == oracle filter of the regenerated head / tail rows
    w = 1 << 20
    for row0 in (0, n - w):
        cols = [CO.synth_column(NO.F32, specs[c], 42, c, row0, w) for c in range(8)]
        exp = NO.query_filter(cols, [0, 2, 1, 4], [(1, NO.GT, 0, tt), (4, NO.LT, 0, uu)])
        k = len(exp[0])
        for j in range(4):
            got = r.column(j, 0, k) if row0 == 0 else r.column(j, cnt - k, k)
            assert np.array_equal(got, exp[j])
    # idempotence: filtering the output by the same predicates (now columns 2 and 3) keeps every row
    r2 = env.query_filter(r, [0], [(2, NO.GT, 0, tt), (3, NO.LT, 0, uu)])
    assert r2.shape[0] == cnt
    r2.free()
    # partition identity: |c2>t ^ c5<u| + |c2>t ^ c5>=u| == |c2>t|
    a = env.query_filter(t, [], [(1, NO.GT, 0, tt), (4, NO.GE, 0, uu)])
    b = env.query_filter(t, [], [(1, NO.GT, 0, tt)])
    assert cnt + a.shape[0] == b.shape[0]
    for x in (a, b, r, t):
        x.free()


# ---- WHERE in conjunctive normal form: OR-clauses and NOT (hark.h HARK_PRED_OR / HARK_PRED_NOT) ----
@pytest.mark.parametrize("dtype", ALL_DT)
@pytest.mark.parametrize("n", [0, 1, 1023, 4097, 32769, 250007])
def test_filter_cnf_clauses(dtype, n):
    env = get_env()
    OR, NOT = NO.PRED_OR, NO.PRED_NOT
    rng = np.random.default_rng(7 * n + dtype)
    a = rand_table(rng, n, 5, dtype, lo=-8 if dtype != NO.U32 else 0, hi=9, nan_frac=0.02)
    if dtype in (NO.F32, NO.F64):
        a = np.round(a * 16) / 16
    c = 0.5 if dtype in (NO.F32, NO.F64) else 3
    iv, fv = (0, c) if dtype in (NO.F32, NO.F64) else (c, float(c))
    pred_lists = [
        [(0, NO.GT | OR, iv, fv), (1, NO.LT, iv, fv)],                                              # a OR b
        [(0, NO.GT | NOT, iv, fv)],                                                                 # NOT a (NaN rows pass)
        [(0, NO.GT | OR, iv, fv), (1, NO.LE | NOT | OR, iv, fv), (2, NO.EQ, iv, fv), (3, NO.NE | NOT, iv, fv)],
        [(4, NO.GE, iv, fv), (0, NO.EQ | OR, iv, fv), (0, NO.EQ | OR, iv + 1, fv + 0.25), (0, NO.EQ, iv - 2, fv - 0.25),
         (2, NO.LT | NOT | OR, iv, fv), (3, NO.GT | NOT, iv, fv)],                                  # c AND a IN (..) AND (..)
        [(p % 5, [NO.GT, NO.LT, NO.NE, NO.GE][p % 4] | (OR if p % 3 != 2 else 0) | (NOT if p % 5 == 1 else 0),
          iv + p % 3 - 1, fv + 0.125 * (p % 3 - 1)) for p in range(15)] + [(1, NO.NE, iv, fv)],     # 16 predicates
    ]
    t = env.to_device(a)
    for preds in pred_lists:
        r = env.query_filter(t, [0, 4, 2], preds)
        exp = NO.query_filter(cols_of(a), [0, 4, 2], preds)
        assert r.shape[0] == len(exp[0]), (dtype, n, preds)
        for j in range(3):
            assert np.array_equal(r.column(j), exp[j], equal_nan=True), (dtype, n, preds, j)
        r.free()
    t.free()


def test_filter_cnf_mixed_widths_and_bad_lists():
    env = get_env()
    OR, NOT = NO.PRED_OR, NO.PRED_NOT
    rng = np.random.default_rng(99)
    n = 70001
    cols = [rng.integers(-5, 6, n).astype(np.int32), rng.integers(-5, 6, n).astype(np.int64),
            rng.random(n).astype(np.float32), rng.random(n)]
    cols[3][rng.integers(0, n, 500)] = np.nan
    t = env.from_columns(cols)
    preds = [(0, NO.GT | OR, 2, 2.0), (3, NO.LT | NOT, 0, 0.5), (1, NO.EQ | OR, -1, -1.0), (2, NO.GE, 0, 0.75)]
    r = env.query_filter(t, [3, 1, 0], preds)
    exp = NO.query_filter(cols, [3, 1, 0], preds)
    for j in range(3):
        assert np.array_equal(r.column(j), exp[j], equal_nan=True)
    r.free()
    # constants outside an i32 column's range fold to always / never, also under NOT
    preds = [(0, NO.LT | NOT | OR, 1 << 40, 0.0), (1, NO.GT, 3, 3.0)]
    r = env.query_filter(t, [1], preds)
    assert np.array_equal(r.column(0), NO.query_filter(cols, [1], preds)[0])
    r.free()
    with pytest.raises(Exception, match="last predicate"):
        env.query_filter(t, [0], [(0, NO.GT | OR, 0, 0.0)])
    with pytest.raises(Exception, match="bad comparison"):
        env.query_filter(t, [0], [(0, NO.GT | 0x400, 0, 0.0)])
    with pytest.raises(Exception, match="at most 16"):
        env.query_filter(t, [0], [(0, NO.GT, 0, 0.0)] * 17)
    t.free()
