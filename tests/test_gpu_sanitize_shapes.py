"""Small-shape pass over every kernel family, checked against the oracle — the workload `make sanitize` runs under
`compute-sanitizer --tool memcheck` and `--tool racecheck` (look-back words, shared-memory rank atomics, the K8t bulk
copies, in-place tie repair, hash-table CAS).  Also a quick -m gpu test on its own.  K8c peer stores need two GPUs and
are not covered here (tests/test_gpu_sharded.py under torchrun is)."""

import numpy as np
import pytest

from oracle import np_oracle as NO
from tests.gpu_util import get_env

pytestmark = pytest.mark.gpu


def _eq(got, exp):
    assert len(got) == len(exp)
    for g, e in zip(got, exp):
        assert g.shape == e.shape and g.dtype == e.dtype
        if g.dtype.kind == "f":
            assert np.allclose(g, e, rtol=1e-5 if g.dtype == np.float32 else 1e-12, atol=0, equal_nan=True)
        else:
            assert np.array_equal(g, e)


def test_filter_shapes():
    env = get_env()
    rng = np.random.default_rng(1)
    n = 40003
    cols = [rng.random(n).astype(np.float32) for _ in range(4)] + [rng.integers(-9, 9, n).astype(np.int64)]
    t = env.from_columns(cols)
    for impl in (0, 1, 3):
        env.set_option("filter.impl", impl)
        preds = [(1, NO.GT, 0, 0.5), (3, NO.LT, 0, 0.5)]
        _eq(env.query_filter(t, [0, 2, 4], preds).columns(), NO.query_filter(cols, [0, 2, 4], preds))
    env.set_option("filter.impl", 0)
    preds = [(4, NO.GT | NO.PRED_OR, 3, 3.0), (0, NO.LT | NO.PRED_NOT, 0, 0.25), (2, NO.NE, 0, 0.5)]
    _eq(env.query_filter(t, [4, 1], preds).columns(), NO.query_filter(cols, [4, 1], preds))
    _eq(env.query_sel(t, [3, 0]).columns(), [cols[3], cols[0]])
    t.free()


@pytest.mark.parametrize("mode", ["dense", "tiles", "chunks", "sort"])
def test_groupby_shapes(mode):
    env = get_env()
    rng = np.random.default_rng(2)
    n = 30011
    cols = [rng.integers(-300, 2000, n).astype(np.int32), rng.integers(-1000, 1000, n).astype(np.int32),
            (rng.integers(0, 2 ** 20, n) * 2.0 ** -20).astype(np.float32), rng.random(n).astype(np.float32)]
    t = env.from_columns(cols)
    env.set_option("groupby.impl", 1 if mode == "sort" else 0)
    env.set_option("dense.log2_slots", 8 if mode in ("tiles", "chunks") else 0)
    env.set_option("dense.part_impl", 1 if mode == "chunks" else 0)
    try:
        ops, sc = [NO.AGG_SUM, NO.AGG_COUNT, NO.AGG_AVG, NO.AGG_MIN, NO.AGG_SUM, NO.AGG_SUM, NO.AGG_SUM64], [1, 1, 1, 1, 2, 3, 1]
        _eq(env.query_groupby_ex(t, 0, sc, ops, having=[(2, NO.GT, 10, 0.0)]).columns(),
            NO.query_groupby_ex(cols, 0, sc, ops, having=[(2, NO.GT, 10, 0.0)]))
        _eq(env.query_groupby_ex(t, 0, [1, 1], [NO.AGG_SUM, NO.AGG_COUNT]).columns(),            # the specialised programmes
            NO.query_groupby_ex(cols, 0, [1, 1], [NO.AGG_SUM, NO.AGG_COUNT]))
        _eq(env.query_groupby_ex(t, 0, [1, 1, 1], [NO.AGG_SUM, NO.AGG_COUNT, NO.AGG_AVG]).columns(),
            NO.query_groupby_ex(cols, 0, [1, 1, 1], [NO.AGG_SUM, NO.AGG_COUNT, NO.AGG_AVG]))
        _eq(env.query_groupby_ex(t, 0, [2, 2], [NO.AGG_SUM, NO.AGG_AVG]).columns(),
            NO.query_groupby_ex(cols, 0, [2, 2], [NO.AGG_SUM, NO.AGG_AVG]))
        db = np.stack([c.view(np.uint32) for c in cols[:2]], axis=1)
        assert np.array_equal(env.from_futhark(env.query_groupby(db, 0, [1, 1], [2, 3])), NO.query_groupby(db, 0, [1, 1], [2, 3]))
    finally:
        env.set_option("groupby.impl", 0)
        env.set_option("dense.log2_slots", 0)
        env.set_option("dense.part_impl", 0)
    t.free()


def test_orderby_shapes():
    env = get_env()
    rng = np.random.default_rng(3)
    n = 50021
    cols = [rng.integers(-40, 40, n).astype(np.int64), rng.integers(-2 ** 62, 2 ** 62, n).astype(np.int64),
            rng.random(n).astype(np.float32), np.arange(n, dtype=np.int32)]
    cols[2][::97] = np.nan
    t = env.from_columns(cols)
    for keys, desc in (([0, 1], [0, 0]), ([2, 0], [1, 0]), ([1], [1])):
        r = env.query_orderby(t, [3, 0, 1, 2], keys, desc)
        exp = NO.query_orderby(cols, [3, 0, 1, 2], keys, desc)
        for g, e in zip(r.columns(), exp):
            assert np.array_equal(g, e, equal_nan=True)
        r.free()
    env.set_option("sort.trunc", 0)
    r = env.query_orderby(t, [3], [0, 1], [0, 0])
    assert np.array_equal(r.column(0), NO.query_orderby(cols, [3], [0, 1], [0, 0])[0])
    env.set_option("sort.trunc", 1)
    # K3b (16-byte rows, TMA run stores): chunk protocol and look-back, digits straddling the two key columns or not
    env.set_option("sort.sweep16_min_rows", 1)
    try:
        exp = NO.query_orderby(cols, [0, 1], [0, 1], [0, 1])
        for mode, straddle in ((1, 1), (1, 0), (2, 1)):
            env.set_option("sort.sweep16", mode)
            env.set_option("sort.straddle", straddle)
            r = env.query_orderby(t, [0, 1], [0, 1], [0, 1])
            assert env.get_option("sort.last_sweep16") == 1
            _eq(r.columns(), exp)
            r.free()
    finally:
        env.set_option("sort.sweep16_min_rows", 1 << 16)
        env.set_option("sort.sweep16", 1)
        env.set_option("sort.straddle", 1)
    t.free()


def test_join_shapes():
    env = get_env()
    rng = np.random.default_rng(4)
    a = rng.integers(0, 500, (20011, 3), dtype=np.int64).astype(np.uint32)
    b = rng.integers(0, 600, (3001, 2), dtype=np.int64).astype(np.uint32)
    assert np.array_equal(env.from_futhark(env.join(a, b, 1, 0, [0, 1, 2], [1])), NO.join(a, b, 1, 0, [0, 1, 2], [1]))
    t1 = [rng.integers(-2 ** 40, 2 ** 40, 9001).astype(np.int64), rng.random(9001)]
    t2 = [np.concatenate([t1[0][:3000], rng.integers(-2 ** 40, 2 ** 40, 2000)]).astype(np.int64), np.arange(5000, dtype=np.int32)]
    d1, d2 = env.from_columns(t1), env.from_columns(t2)
    exp = NO.join_ex(t1, t2, 0, 0, [0, 1], [1])
    _eq(env.join_ex(d1, d2, 0, 0, [0, 1], [1], 1).columns(), exp)
    got = env.join_ex(d1, d2, 0, 0, [0, 1], [1], 0).columns()
    o1, o2 = np.lexsort((got[2], got[0])), np.lexsort((exp[2], exp[0]))
    _eq([g[o1] for g in got], [e[o2] for e in exp])
    d1.free(); d2.free()


@pytest.mark.parametrize("build,slices", [(1, 8192), (2, 8192), (2, 16 << 20), (1, 16 << 20)])
def test_join_groupby_shapes(build, slices):
    env = get_env()
    rng = np.random.default_rng(5)
    nd, nf = 20011, 120007
    pk = (rng.permutation(nd) - 3000).astype(np.int32)
    dim = [pk, rng.integers(-3, 60, nd).astype(np.int32)]
    fact = [rng.integers(-4000, nd, nf).astype(np.int32), rng.integers(0, 1000, nf).astype(np.int32)]
    d, f = env.from_columns(dim), env.from_columns(fact)
    env.set_option("join.build", build)
    env.set_option("join.lut_slice_bytes", slices)
    try:
        ops = [NO.AGG_SUM, NO.AGG_COUNT]
        _eq(env.join_groupby(f, d, 0, 0, 1, [1, 1], ops).columns(), NO.join_groupby(fact, dim, 0, 0, 1, [1, 1], ops))
        ops = [NO.AGG_SUM, NO.AGG_COUNT, NO.AGG_AVG, NO.AGG_MAX]
        _eq(env.join_groupby(f, d, 0, 0, 1, [1, 1, 1, 1], ops).columns(), NO.join_groupby(fact, dim, 0, 0, 1, [1, 1, 1, 1], ops))
    finally:
        env.set_option("join.build", 0)
        env.set_option("join.lut_slice_bytes", 16 << 20)
    d.free(); f.free()


def test_partition_building_blocks():
    env = get_env()
    rng = np.random.default_rng(6)
    n = 30011
    cols = [rng.integers(-1000, 1000, n).astype(np.int32), np.arange(n, dtype=np.int64)]
    t = env.from_columns(cols)
    p, counts = env.partition_by_hash(t, 0, 13)
    assert sum(counts) == n
    p.free()
    s = env.sort_by(t, 0)
    assert np.array_equal(s.column(1), NO.query_orderby(cols, [1], [0])[0])
    s.free(); t.free()
