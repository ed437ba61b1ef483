"""Cross-checks between the three oracle layers (SOAC simulation, C port, numpy twin) on random
inputs, plus the extension semantics (filter, generator).  CPU only."""

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import c_oracle as CO
from oracle import hark_ref as R
from oracle import np_oracle as NO

u32s = st.integers(min_value=0, max_value=2 ** 32 - 1)
small_keys = st.sampled_from([0, 1, 2, 3, 2 ** 31 - 1, 2 ** 31, 2 ** 32 - 2, 2 ** 32 - 1])


@st.composite
def tables(draw, max_rows=24, max_cols=4, key_strategy=small_keys):
    n = draw(st.integers(0, max_rows))
    m = draw(st.integers(1, max_cols))
    rows = [[draw(key_strategy) if c == 0 else draw(u32s) for c in range(m)] for _ in range(n)]
    return rows, m


@settings(max_examples=60, deadline=None)
@given(tables(), st.data())
def test_groupby_three_ways(tbl, data):
    rows, m = tbl
    c = data.draw(st.integers(0, 3))
    s_cols = [data.draw(st.integers(0, m - 1)) for _ in range(c)]
    t_cols = [data.draw(st.integers(0, 5)) for _ in range(c)]
    g_col = data.draw(st.integers(0, m - 1))
    sim = R.query_groupby(rows, g_col, s_cols, t_cols)
    db = np.asarray(rows, dtype=np.int64).astype(np.uint32).reshape(len(rows), m)
    exp = np.asarray(sim, dtype=np.int64).astype(np.uint32).reshape(-1, c + 1)
    assert np.array_equal(CO.query_groupby(db, g_col, s_cols, t_cols), exp)
    assert np.array_equal(NO.query_groupby(db, g_col, s_cols, t_cols), exp)


@settings(max_examples=60, deadline=None)
@given(tables(max_rows=14), tables(max_rows=14), st.data())
def test_join_three_ways(t1, t2, data):
    (r1, m), (r2, t) = t1, t2
    col1 = data.draw(st.integers(0, m - 1))
    col2 = data.draw(st.integers(0, t - 1))
    cols1 = [data.draw(st.integers(0, m - 1)) for _ in range(data.draw(st.integers(0, 2)))]
    cols2 = [data.draw(st.integers(0, t - 1)) for _ in range(data.draw(st.integers(0, 2)))]
    sim = R.join(r1, r2, col1, col2, cols1, cols2)
    w = len(cols1) + len(cols2)
    a = np.asarray(r1, dtype=np.int64).astype(np.uint32).reshape(len(r1), m)
    b = np.asarray(r2, dtype=np.int64).astype(np.uint32).reshape(len(r2), t)
    got_c = CO.join(a, b, col1, col2, cols1, cols2)
    got_n = NO.join(a, b, col1, col2, cols1, cols2)
    if w == 0:   # zero-width rows: only the row count is observable
        assert got_c.shape[0] == len(sim) and got_n.shape[0] == len(sim)
    else:
        exp = np.asarray(sim, dtype=np.int64).astype(np.uint32).reshape(-1, w)
        assert np.array_equal(got_c, exp)
        assert np.array_equal(got_n, exp)


@settings(max_examples=40, deadline=None)
@given(tables(key_strategy=u32s), st.data())
def test_sel_three_ways(tbl, data):
    rows, m = tbl
    cols = [data.draw(st.integers(0, m - 1)) for _ in range(data.draw(st.integers(0, 5)))]
    sim = R.query_sel(rows, cols)
    db = np.asarray(rows, dtype=np.int64).astype(np.uint32).view(np.int32).reshape(len(rows), m)
    exp = np.asarray(sim, dtype=np.int64).astype(np.uint32).view(np.int32).reshape(len(rows), len(cols))
    assert np.array_equal(CO.query_sel(db, cols), exp)
    assert np.array_equal(NO.query_sel(db, cols), exp)


@pytest.mark.parametrize("dtype", [NO.I32, NO.U32, NO.I64, NO.F32, NO.F64])
@pytest.mark.parametrize("threads", [1, 3])
def test_filter_c_vs_numpy(dtype, threads):
    rng = np.random.default_rng(dtype * 7 + threads)
    n, m = 1000, 5
    npdt = NO.NP_DTYPES[dtype]
    if dtype in (NO.F32, NO.F64):
        db = rng.random((n, m)).astype(npdt)
        db[rng.integers(0, n, 20), rng.integers(0, m, 20)] = np.nan
        preds = [(1, NO.GT, 0, 0.5), (4, NO.LT, 0, 0.75)]
    else:
        db = rng.integers(0, 100, (n, m)).astype(npdt)
        preds = [(1, NO.GE, 50, 0.0), (4, NO.NE, 7, 0.0), (0, NO.LE, 90, 0.0)]
    cols = [0, 2, 2]
    exp = NO.query_filter([db[:, c].copy() for c in range(m)], cols, preds)
    got = CO.query_filter(db, cols, preds, threads=threads)
    assert got.shape == (len(exp[0]), 3)
    for j in range(3):
        assert np.array_equal(got[:, j], exp[j], equal_nan=True)
    # empty predicate list keeps every row; empty table works
    assert CO.query_filter(db, [1], [], threads=threads).shape == (n, 1)
    assert CO.query_filter(db[:0], [1], preds, threads=threads).shape == (0, 1)


@pytest.mark.parametrize("op", [NO.GT, NO.GE, NO.LT, NO.LE, NO.EQ, NO.NE])
def test_filter_ops_and_nan(op):
    col = np.array([0.0, 0.5, 1.0, np.nan, -0.0, 0.5], dtype=np.float32)
    db = np.stack([col, col], axis=1)
    exp = NO.query_filter([col, col], [0], [(1, op, 0, 0.5)])[0]
    got = CO.query_filter(db, [0], [(1, op, 0, 0.5)])[:, 0]
    assert np.array_equal(got, exp, equal_nan=True)
    ref = {NO.GT: [1.0], NO.GE: [0.5, 1.0, 0.5], NO.LT: [0.0, -0.0], NO.LE: [0.0, 0.5, -0.0, 0.5],
           NO.EQ: [0.5, 0.5]}
    if op in ref:
        assert got.tolist() == ref[op]
    else:
        assert len(got) == 4 and np.isnan(got[2])   # NaN != c is true (IEEE)


@pytest.mark.parametrize("dtype", [NO.I32, NO.U32, NO.I64, NO.F32, NO.F64])
def test_generator_c_vs_numpy(dtype):
    specs = [dict(kind=NO.GEN_UNIFORM, lo=-5, range=11, flo=0.0, fhi=1.0),
             dict(kind=NO.GEN_UNIFORM, lo=0, range=0, flo=-2.0, fhi=3.0),
             dict(kind=NO.GEN_UNIFORM, lo=-(2 ** 19), range=2 ** 20),
             dict(kind=NO.GEN_AFFINE, a=7, b=3, range=1000),
             dict(kind=NO.GEN_CONST, lo=42, flo=4.25), dict(kind=NO.GEN_LOGUNIFORM, lo=-3, range=1 << 20),
             dict(kind=NO.GEN_LOGUNIFORM, lo=5, range=1000),
             dict(kind=NO.GEN_AFFINE_UNIFORM, a=2654435761, b=977, range=100003)]
    for col, spec in enumerate(specs):
        a = CO.synth_column(dtype, spec, 42, col, 10 ** 9 - 5, 300)
        b = NO.synth_column(dtype, spec, 42, col, 10 ** 9 - 5, 300)
        assert a.dtype == b.dtype and np.array_equal(a, b), (dtype, spec)
    # a row range regenerates independently of where it starts
    full = CO.synth_column(dtype, specs[0], 1, 0, 0, 100)
    assert np.array_equal(full[40:60], CO.synth_column(dtype, specs[0], 1, 0, 40, 20))


def test_generator_distribution():
    u = CO.synth_column(NO.F32, dict(kind=NO.GEN_UNIFORM), 42, 1, 0, 1 << 18)
    assert 0.0 <= u.min() and u.max() < 1.0 and abs(u.mean() - 0.5) < 5e-3
    k = CO.synth_column(NO.I32, dict(kind=NO.GEN_UNIFORM, lo=0, range=1 << 10), 42, 0, 0, 1 << 18)
    assert k.min() == 0 and k.max() == 1023 and len(np.unique(k)) == 1024
    pk = CO.synth_column(NO.I32, dict(kind=NO.GEN_AFFINE, a=48271, b=11, range=100003), 42, 0, 0, 100003)
    assert len(np.unique(pk)) == 100003


def test_groupby_ex_and_orderby_semantics():
    key = np.array([3, -1, 3, 7, -1, 3], dtype=np.int32)
    val = np.array([1.5, 2.0, 0.25, 4.0, 8.0, 1.0], dtype=np.float32)
    out = NO.query_groupby_ex([key, val], 0, [1, 1, 1, 1, 1], [NO.AGG_SUM, NO.AGG_COUNT, NO.AGG_AVG, NO.AGG_MAX, NO.AGG_MIN])
    assert out[0].tolist() == [-1, 3, 7]                    # signed order for i32 keys
    assert out[1].tolist() == [10.0, 2.75, 4.0] and out[1].dtype == np.float32
    assert out[2].tolist() == [2, 3, 1] and out[2].dtype == np.int64
    assert np.allclose(out[3], [5.0, 2.75 / 3, 4.0]) and out[3].dtype == np.float64
    hv = NO.query_groupby_ex([key, val], 0, [1], [NO.AGG_COUNT], having=[(1, NO.GT, 1, 0.0)])
    assert hv[0].tolist() == [-1, 3] and hv[1].tolist() == [2, 3]
    a = np.array([2, 1, 2, 1, -5], dtype=np.int64)
    b = np.array([0.5, np.nan, -0.0, 0.0, 1.0], dtype=np.float64)
    o = NO.query_orderby([a, b], [0, 1], [0, 1], [0, 0])
    assert o[0].tolist() == [-5, 1, 1, 2, 2]
    assert o[1][1] == 0.0 and np.isnan(o[1][2])            # NaN last within the tie
    assert np.signbit(o[1][3]) and o[1][4] == 0.5          # -0.0 before 0.5
    d = NO.query_orderby([a, b], [0], [0], [1])
    assert d[0].tolist() == [2, 2, 1, 1, -5]


def test_join_ex_restates_the_pinned_join_on_u32_data():
    """np_oracle.join_ex (typed extension) == np_oracle.join (join.fut:52-75 restatement) == the SOAC simulation on u32."""
    rng = np.random.default_rng(3)
    a = rng.integers(0, 30, (200, 3)).astype(np.uint32)
    b = rng.integers(0, 30, (150, 2)).astype(np.uint32)
    ref = NO.join(a, b, 1, 0, [0, 2], [1])
    got = NO.join_ex([a[:, c].copy() for c in range(3)], [b[:, c].copy() for c in range(2)], 1, 0, [0, 2], [1])
    assert np.array_equal(np.stack(got, axis=1), ref)
