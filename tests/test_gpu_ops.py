"""Parity (-m gpu) of GROUP BY, ORDER BY, JOIN and the sort/partition building blocks through the C-ABI vs the
oracle.  Integer results, row sets and orders are bit-exact; floating SUM/AVG within 1e-5 (f32) / 1e-12 (f64)
relative, the tolerance north_star states (reduction order differs)."""

import json
import os

import numpy as np
import pytest

from oracle import np_oracle as NO
from tests.gpu_util import as_torch, cols_of, free_gb, get_env, need_gpu, rand_table

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
DATA = json.load(open(os.path.join(GOLDEN, "data_csv.json")))
VEC = json.load(open(os.path.join(GOLDEN, "harkdb_vectors.json")))["cases"]
SIZES = [1, 2, 31, 127, 128, 129, 4095, 4096, 4097, 8193, 100003, (1 << 20) + 5]


def _db(ref):
    return DATA["rows"] if ref == "data_csv" else ref


@pytest.fixture(params=["auto", "sort", "tiny_tables"])
def gb_impl(request):
    """GROUP BY strategies: `auto` = K2 dense/partitioned aggregation when the key range allows it,
    `sort` = K3 radix sort + K4 segmented reduce only, `tiny_tables` = K2 with 256-slot shared-memory tables so that
    small inputs exercise the partition pass and the bucket-to-bucket table merges."""
    env = get_env()
    env.set_option("groupby.impl", 1 if request.param == "sort" else 0)
    env.set_option("dense.log2_slots", 8 if request.param == "tiny_tables" else 0)
    yield request.param
    env.set_option("groupby.impl", 0)
    env.set_option("dense.log2_slots", 0)


# ---------------------------------------------------------------- GROUP BY (reference-pinned, u32)
@pytest.mark.parametrize("idx", [i for i, c in enumerate(VEC) if c["kind"] == "query_groupby"])
def test_groupby_golden_vectors(idx):
    env = get_env()
    case = VEC[idx]
    db = np.asarray(_db(case["db"]), dtype=np.int64).astype(np.uint32)
    exp = np.asarray(case["output"], dtype=np.int64).astype(np.uint32).reshape(-1, len(case["s_cols"]) + 1)
    got = env.from_futhark(env.query_groupby(db, case["g_col"], case["s_cols"], case["t_cols"]))
    assert got.dtype == np.uint32 and np.array_equal(got, exp)


def test_groupby_test_py_query():
    env = get_env()
    db = np.asarray(DATA["rows"], dtype=np.int64)       # int64 from pandas, converted with a range check
    got = env.from_futhark(env.query_groupby(db, 0, np.array([0, 2]), np.array([0, 3])))   # test.py:7 plan
    assert got.tolist() == [[0, 0, 0], [1, 1, 3], [6, 6, 6]]


@pytest.mark.parametrize("n", [0] + SIZES)
@pytest.mark.parametrize("keys", ["few", "distinct", "one", "high", "runs"])
def test_groupby_u32_vs_oracle(gb_impl, n, keys):
    env = get_env()
    rng = np.random.default_rng(n * 7 + len(keys))
    m = 4
    db = rng.integers(0, 2 ** 32, (n, m), dtype=np.uint64).astype(np.uint32)
    if keys == "few":
        db[:, 0] = rng.integers(0, 7, n)
    elif keys == "distinct":
        db[:, 0] = rng.permutation(n).astype(np.uint32)
    elif keys == "one":
        db[:, 0] = 12345
    elif keys == "high":
        db[:, 0] = rng.integers(2 ** 31 - 3, 2 ** 31 + 3, n)     # unsigned order across the sign bit
    else:
        db[:, 0] = (np.arange(n) // 1000).astype(np.uint32)[rng.permutation(n)] if n else 0
    s_cols, t_cols = [1, 2, 3, 1, 0, 2], [1, 2, 3, 4, 0, 9]     # prod sum max min key(->min) unknown(->min)
    got = env.from_futhark(env.query_groupby(db, 0, s_cols, t_cols))
    exp = NO.query_groupby(db, 0, s_cols, t_cols)
    assert got.shape == exp.shape and np.array_equal(got, exp)


def test_groupby_errors():
    env = get_env()
    from harkdb_b200.hark_ffi import HarkError
    db = np.zeros((4, 2), dtype=np.uint32)
    with pytest.raises(HarkError, match="out of bounds"):
        env.query_groupby(db, 2, [0], [2])
    with pytest.raises(HarkError, match="out of bounds"):
        env.query_groupby(db, 0, [5], [2])
    with pytest.raises(HarkError):
        env.query_groupby(db, 0, [1, 1], [2])                   # t_cols shorter than s_cols (groupby.fut:47)
    f = env.to_device(np.zeros((4, 2), dtype=np.float32))
    with pytest.raises(HarkError, match="u32"):
        env.query_groupby(f, 0, [1], [2])
    f.free()


# ---------------------------------------------------------------- GROUP BY (typed extension)
def _check_cols(got_cols, exp_cols, in_dtypes=None):
    assert len(got_cols) == len(exp_cols)
    for g, e in zip(got_cols, exp_cols):
        assert g.dtype == e.dtype and g.shape == e.shape, (g.dtype, e.dtype, g.shape, e.shape)
        if g.dtype.kind == "f":
            tol = 1e-5 if g.dtype == np.float32 else 1e-12
            assert np.allclose(g, e, rtol=tol, atol=0), np.max(np.abs(g - e) / np.maximum(np.abs(e), 1e-300))
        else:
            assert np.array_equal(g, e)


@pytest.mark.parametrize("kdt", [NO.I32, NO.U32, NO.I64])
@pytest.mark.parametrize("vdt", [NO.I32, NO.U32, NO.I64, NO.F32, NO.F64])
def test_groupby_ex_typed(gb_impl, kdt, vdt):
    env = get_env()
    rng = np.random.default_rng(kdt * 10 + vdt)
    n = 200003
    lo = -500 if kdt != NO.U32 else 0
    key = rng.integers(lo, 500, n).astype(NO.NP_DTYPES[kdt])
    if vdt in (NO.F32, NO.F64):
        val = rng.random(n).astype(NO.NP_DTYPES[vdt])
    else:
        val = rng.integers(-1000 if vdt != NO.U32 else 0, 1000, n).astype(NO.NP_DTYPES[vdt])
    other = rng.integers(0, 5, n).astype(np.int32)
    t = env.from_columns([other, key, val])
    ops = [NO.AGG_SUM, NO.AGG_COUNT, NO.AGG_AVG, NO.AGG_MAX, NO.AGG_MIN, NO.AGG_KEY]
    s_cols = [2, 2, 2, 2, 2, 1]
    r = env.query_groupby_ex(t, 1, s_cols, ops)
    exp = NO.query_groupby_ex([other, key, val], 1, s_cols, ops)
    _check_cols(r.columns(), exp)
    hv = [(2, NO.GT, int(np.median(exp[2])), 0.0), (0, NO.LT, 100, 0.0)]     # HAVING count > median AND key < 100
    r2 = env.query_groupby_ex(t, 1, s_cols, ops, having=hv)
    _check_cols(r2.columns(), NO.query_groupby_ex([other, key, val], 1, s_cols, ops, having=hv))
    for x in (r, r2, t):
        x.free()


def test_groupby_ex_prod_and_wrap(gb_impl):
    env = get_env()
    rng = np.random.default_rng(77)
    n = 50000
    key = rng.integers(0, 50, n).astype(np.int32)
    vi = rng.integers(-2 ** 31, 2 ** 31, n).astype(np.int32)
    vl = rng.integers(-2 ** 62, 2 ** 62, n).astype(np.int64)
    vf = (1.0 + rng.random(n) * 1e-4).astype(np.float64)
    t = env.from_columns([key, vi, vl, vf])
    ops = [NO.AGG_PROD, NO.AGG_SUM, NO.AGG_PROD, NO.AGG_SUM, NO.AGG_PROD]
    s_cols = [1, 1, 2, 2, 3]
    r = env.query_groupby_ex(t, 0, s_cols, ops)
    _check_cols(r.columns(), NO.query_groupby_ex([key, vi, vl, vf], 0, s_cols, ops))
    r.free(); t.free()


@pytest.mark.parametrize("n", [0, 1, 4096, 4097, 300000])
def test_groupby_ex_sizes_and_single_group(gb_impl, n):
    env = get_env()
    rng = np.random.default_rng(n)
    key = np.full(n, -7, dtype=np.int64)
    val = rng.random(n).astype(np.float32)
    t = env.from_columns([key, val])
    r = env.query_groupby_ex(t, 0, [1, 1, 1], [NO.AGG_SUM, NO.AGG_COUNT, NO.AGG_AVG])
    _check_cols(r.columns(), NO.query_groupby_ex([key, val], 0, [1, 1, 1], [NO.AGG_SUM, NO.AGG_COUNT, NO.AGG_AVG]))
    r.free(); t.free()


# ---------------------------------------------------------------- GROUP BY over several key columns
@pytest.mark.parametrize("n", [0, 1, 5, 4097, 200003])
def test_groupby_multi_vs_oracle(gb_impl, n):
    env = get_env()
    rng = np.random.default_rng(31 + n)
    cols = [rng.integers(-3, 4, n).astype(np.int32), rng.integers(10 ** 12, 10 ** 12 + 40, n).astype(np.int64),
            rng.integers(0, 3, n).astype(np.uint32), rng.integers(-1000, 1000, n).astype(np.int32),
            rng.random(n).astype(np.float32), rng.random(n)]
    t = env.from_columns(cols)
    ops = [NO.AGG_SUM, NO.AGG_COUNT, NO.AGG_AVG, NO.AGG_MIN, NO.AGG_MAX, NO.AGG_SUM, NO.AGG_KEY]
    s_cols = [3, 3, 4, 3, 5, 4, 0]
    for g_cols in ([0, 1, 2], [2, 0], [1, 0]):
        r = env.query_groupby_multi(t, g_cols, s_cols, ops)
        _check_cols(r.columns(), NO.query_groupby_multi(cols, g_cols, s_cols, ops))
        assert r.dtypes[:len(g_cols)] == [t.dtypes[g] for g in g_cols]
        r.free()
    having = [(2, NO.GT | NO.PRED_OR, 0, 0.0), (3, NO.GE, 3, 3.0)]          # SUM > 0 OR COUNT >= 3 (output columns)
    r = env.query_groupby_multi(t, [0, 2], [3, 3], [NO.AGG_SUM, NO.AGG_COUNT], having)
    _check_cols(r.columns(), NO.query_groupby_multi(cols, [0, 2], [3, 3], [NO.AGG_SUM, NO.AGG_COUNT], having))
    r.free()
    r = env.query_groupby_multi(t, [1], [3], [NO.AGG_MAX])                   # one key: the single-key operator
    _check_cols(r.columns(), NO.query_groupby_ex(cols, 1, [3], [NO.AGG_MAX]))
    r.free(); t.free()


def test_groupby_multi_wide_composite_and_errors(gb_impl):
    env = get_env()
    from harkdb_b200.hark_ffi import HarkError
    rng = np.random.default_rng(77)
    n = 150001
    # 20 + 30 + 12 = 62 bits of combined range: an i64 composite, sorted (too wide for the dense tables)
    cols = [rng.integers(0, 1 << 20, n).astype(np.uint32), rng.integers(-(1 << 29), 1 << 29, n).astype(np.int64) // 977 * 977,
            rng.integers(-2048, 2048, n).astype(np.int32), rng.integers(0, 10, n).astype(np.int32)]
    cols[0][: n // 2] = cols[0][n // 2: 2 * (n // 2)]        # make groups of more than one row
    cols[1][: n // 2] = cols[1][n // 2: 2 * (n // 2)]
    cols[2][: n // 2] = cols[2][n // 2: 2 * (n // 2)]
    t = env.from_columns(cols)
    r = env.query_groupby_multi(t, [0, 1, 2], [3, 3], [NO.AGG_SUM, NO.AGG_COUNT])
    _check_cols(r.columns(), NO.query_groupby_multi(cols, [0, 1, 2], [3, 3], [NO.AGG_SUM, NO.AGG_COUNT]))
    r.free(); t.free()
    wide = [rng.integers(-2 ** 62, 2 ** 62, 1000, dtype=np.int64), rng.integers(0, 4, 1000).astype(np.int32),
            rng.random(1000)]
    t = env.from_columns(wide)
    with pytest.raises(HarkError, match="more than 63 bits"):
        env.query_groupby_multi(t, [0, 1], [1], [NO.AGG_COUNT])
    with pytest.raises(HarkError, match="integer columns"):
        env.query_groupby_multi(t, [1, 2], [1], [NO.AGG_COUNT])
    with pytest.raises(HarkError, match="out of bounds"):
        env.query_groupby_multi(t, [1, 3], [1], [NO.AGG_COUNT])
    t.free()


# ---------------------------------------------------------------- ORDER BY
@pytest.mark.parametrize("dtype", [NO.I32, NO.U32, NO.I64, NO.F32, NO.F64])
@pytest.mark.parametrize("n", [0] + SIZES)
def test_orderby_single_key(dtype, n):
    env = get_env()
    rng = np.random.default_rng(n + dtype)
    if dtype in (NO.F32, NO.F64):
        k = ((rng.random(n) - 0.5) * 100).astype(NO.NP_DTYPES[dtype])
        if n > 10:
            k[rng.integers(0, n, 5)] = np.nan
            k[rng.integers(0, n, 5)] = -0.0
            k[rng.integers(0, n, 5)] = 0.0
            k[rng.integers(0, n, 3)] = np.inf
            k[rng.integers(0, n, 3)] = -np.inf
    else:
        info = np.iinfo(NO.NP_DTYPES[dtype])
        k = rng.integers(info.min, info.max, n, dtype=np.int64 if dtype != NO.U32 else np.uint64).astype(NO.NP_DTYPES[dtype]) \
            if n % 2 else rng.integers(0 if dtype == NO.U32 else -50, 50, n).astype(NO.NP_DTYPES[dtype])
    payload = np.arange(n, dtype=np.int32)
    t = env.from_columns([k, payload])
    for desc in (0, 1):
        r = env.query_orderby(t, [0, 1], [0], [desc])
        exp = NO.query_orderby([k, payload], [0, 1], [0], [desc])
        assert np.array_equal(r.column(1), exp[1]), "permutation (stability) differs"
        assert np.array_equal(r.column(0), exp[0], equal_nan=True)
        r.free()
    t.free()


def test_orderby_multi_key_config4_shape():
    """BASELINE config 4 at an oracle-checkable size: ORDER BY col1, col2 on i64 with many ties in col1."""
    env = get_env()
    n = (1 << 20) + 77
    specs = [dict(kind=NO.GEN_UNIFORM, lo=-(2 ** 19), range=2 ** 20), dict(kind=NO.GEN_UNIFORM, lo=0, range=0)]
    t = env.synth(n, [NO.I64, NO.I64], specs, seed=42)
    from oracle import c_oracle as CO
    cols = [CO.synth_column(NO.I64, specs[c], 42, c, 0, n) for c in range(2)]
    r = env.query_orderby(t, [0, 1], [0, 1])
    exp = NO.query_orderby(cols, [0, 1], [0, 1])
    assert np.array_equal(r.column(0), exp[0]) and np.array_equal(r.column(1), exp[1])
    assert env.stats()["rows_out"] == n
    r.free()
    r = env.query_orderby(t, [1], [0, 1], [1, 0])          # DESC, ASC; project only col2
    exp = NO.query_orderby(cols, [1], [0, 1], [1, 0])
    assert np.array_equal(r.column(0), exp[0])
    r.free(); t.free()


def test_orderby_mixed_types_payload_and_repeats():
    env = get_env()
    rng = np.random.default_rng(4)
    n = 70001
    cols = [rng.integers(0, 4, n).astype(np.int32), rng.random(n).astype(np.float32).round(1),
            rng.integers(-3, 3, n).astype(np.int64), rng.random(n).astype(np.float64), np.arange(n, dtype=np.uint32)]
    t = env.from_columns(cols)
    r = env.query_orderby(t, [4, 3, 0, 0, 1], [0, 1, 2, 0], [1, 0, 1, 0])
    exp = NO.query_orderby(cols, [4, 3, 0, 0, 1], [0, 1, 2, 0], [1, 0, 1, 0])
    for j in range(5):
        assert np.array_equal(r.column(j), exp[j])
    r.free()
    r = env.query_orderby(t, [2, 1], [], [])                # no keys: plain projection
    assert np.array_equal(r.column(0), cols[2]) and np.array_equal(r.column(1), cols[1])
    r.free(); t.free()


# ---------------------------------------------------------------- K3t: truncated LSD passes + tie repair
def _sort_info(env):
    return {k: env.get_option("sort.last_" + k) for k in ("passes", "truncated", "fix_runs", "fallback")}


def _check_orderby(env, cols, sel, keys, desc):
    t = env.from_columns(cols)
    r = env.query_orderby(t, sel, keys, desc)
    info = _sort_info(env)
    exp = NO.query_orderby(cols, sel, keys, desc)
    for j in range(len(sel)):
        assert np.array_equal(r.column(j), exp[j], equal_nan=True), f"column {j} differs ({info})"
    r.free(); t.free()
    return info


def test_orderby_truncated_passes_repair_short_runs():
    """Wide random keys: only the top digits are sorted, the (many, short) prefix ties are repaired in place."""
    env = get_env()
    rng = np.random.default_rng(11)
    n = 300001
    k1 = rng.integers(-2 ** 63, 2 ** 63 - 1, n, dtype=np.int64)
    k2 = rng.integers(-2 ** 63, 2 ** 63 - 1, n, dtype=np.int64)
    k1[rng.integers(0, n, 2000)] = k1[rng.integers(0, n, 2000)]          # full duplicates of the first key
    payload = np.arange(n, dtype=np.int32)
    env.set_option("sort.trunc_slack", 0)                                 # ~2 % of the rows tie on the prefix
    try:
        info = _check_orderby(env, [k1, k2, payload], [2, 0, 1], [0, 1], [0, 1])
        assert info["truncated"] == 1 and info["fallback"] == 0 and info["fix_runs"] > 100 and info["passes"] <= 4, info
        info = _check_orderby(env, [k1, k2, payload], [2], [0], [1])
        assert info["truncated"] == 1 and info["fallback"] == 0, info
    finally:
        env.set_option("sort.trunc_slack", 4)
    env.set_option("sort.trunc", 0)
    try:
        info = _check_orderby(env, [k1, k2, payload], [2, 0, 1], [0, 1], [0, 1])
        assert info["truncated"] == 0 and info["passes"] == 16, info
    finally:
        env.set_option("sort.trunc", 1)


def test_orderby_truncation_at_a_key_boundary_and_floats():
    env = get_env()
    rng = np.random.default_rng(12)
    n = 200003
    # 24 significant bits in key 0 cover log2(n) + slack: key 1 (f64, with NaNs / infinities / signed zeros) is
    # never touched by a pass and only decides inside the repaired runs
    k0 = rng.integers(0, 1 << 24, n).astype(np.uint32)
    k1 = rng.standard_normal(n)
    k1[rng.integers(0, n, 50)] = np.nan
    k1[rng.integers(0, n, 50)] = -0.0
    k1[rng.integers(0, n, 50)] = 0.0
    k1[rng.integers(0, n, 20)] = np.inf
    payload = np.arange(n, dtype=np.int64)
    info = _check_orderby(env, [k0, k1, payload], [2, 1], [0, 1], [0, 1])
    assert info["truncated"] == 1 and info["passes"] == 3 and info["fix_runs"] > 0 and info["fallback"] == 0, info
    # a signed 4-byte key, descending, as the last sorted key (the tie scan's 4-byte integer path, both directions)
    k0s = (rng.integers(0, 1 << 24, n) - (1 << 23)).astype(np.int32)
    k2 = rng.integers(-2 ** 62, 2 ** 62, n).astype(np.int64)
    for d0 in (1, 0):
        info = _check_orderby(env, [k0s, k2, payload], [2, 0], [0, 1], [d0, 1])
        assert info["truncated"] == 1 and info["fix_runs"] > 0 and info["fallback"] == 0, info
    # f32 keys: clustered around a few exponents — whatever the planner decides, the order must be exact
    f = (rng.standard_normal(n) * 1e3).astype(np.float32)
    f[rng.integers(0, n, 30)] = np.nan
    info = _check_orderby(env, [f, payload], [1, 0], [0], [0])
    assert info["fallback"] == 0, info
    info = _check_orderby(env, [f, k0, payload], [2], [0, 1], [1, 1])
    assert info["fallback"] == 0, info


def test_orderby_truncation_duplicates_cost_nothing_and_clusters_keep_their_passes():
    env = get_env()
    rng = np.random.default_rng(13)
    n = 250000
    payload = np.arange(n, dtype=np.int32)
    # 1000 distinct 64-bit values: every prefix tie is a full duplicate — truncated, nothing to repair, stable
    vals = rng.integers(-2 ** 63, 2 ** 63 - 1, 1000, dtype=np.int64)
    k = vals[rng.integers(0, 1000, n)]
    info = _check_orderby(env, [k, payload], [1, 0], [0], [0])
    assert info["truncated"] == 1 and info["fix_runs"] == 0 and info["fallback"] == 0, info
    # two far-apart clusters: the top digits say nothing, the sample must veto the truncation
    c = rng.integers(0, 1 << 20, n).astype(np.int64)
    c[::2] += 1 << 60
    info = _check_orderby(env, [c, payload], [1, 0], [0], [0])
    assert info["fallback"] == 0 and info["truncated"] == 0, info


def test_orderby_truncation_falls_back_on_a_long_run():
    """40 rows share their top 40 bits among a million random keys: the sample cannot see them, the repair meets a
    run longer than it handles in place, and the sort is redone with every pass — same exact result."""
    env = get_env()
    rng = np.random.default_rng(14)
    n = 1 << 20
    k = rng.integers(-2 ** 63, 2 ** 63 - 1, n, dtype=np.int64)
    pos = rng.choice(n, 40, replace=False)
    k[pos] = (np.int64(0x1234567890) << 24) | rng.integers(0, 1 << 24, 40).astype(np.int64)
    payload = np.arange(n, dtype=np.int32)
    info = _check_orderby(env, [k, payload], [1, 0], [0], [0])
    assert info["fallback"] == 1 and info["truncated"] == 0 and info["passes"] == 8, info
    k[pos[:20]] = rng.integers(-2 ** 63, 2 ** 63 - 1, 20, dtype=np.int64)   # 20 rows left: repaired in place
    info = _check_orderby(env, [k, payload], [1, 0], [0], [0])
    assert info["fallback"] == 0 and info["truncated"] == 1 and info["fix_runs"] >= 1, info


def test_sort_by_and_partition_by_hash():
    env = get_env()
    rng = np.random.default_rng(6)
    n = 123457
    cols = [rng.integers(-1000, 1000, n).astype(np.int32), rng.random(n).astype(np.float64), np.arange(n, dtype=np.int64)]
    t = env.from_columns(cols)
    s = env.sort_by(t, 0)
    exp = NO.query_orderby(cols, [0, 1, 2], [0])
    for j in range(3):
        assert np.array_equal(s.column(j), exp[j])
    for nparts in (1, 2, 8, 13, 256):
        p, counts = env.partition_by_hash(t, 0, nparts)
        assert sum(counts) == n and len(counts) == nparts
        key, rid = p.column(0), p.column(2)
        off = 0
        seen = {}
        for b, cnt in enumerate(counts):
            kb, rb = key[off:off + cnt], rid[off:off + cnt]
            assert np.all(np.diff(rb) > 0), "partition is not stable"
            assert np.array_equal(cols[0][rb], kb) and np.array_equal(cols[1][rb], p.column(1)[off:off + cnt])
            for kv in np.unique(kb):
                assert seen.setdefault(int(kv), b) == b, "one key landed in two buckets"
            off += cnt
        if nparts >= 8:
            assert max(counts) < 3 * n / nparts + 2000
        p.free()
    s.free(); t.free()


@pytest.mark.parametrize("n", [1, 2, 4095, 4096, 4097, 8193, 100003, (1 << 20) + 5])
@pytest.mark.parametrize("shape", ["cfg4", "ties", "one_digit", "f64_desc", "wide_both"])
def test_orderby_sweep16_rows(n, shape):
    """K3b (hk_sweep16_kernel: 16-byte rows, look-back, TMA run stores) on every size class and key shape: LSD over
    several passes is only correct if every pass is stable, so agreement with the oracle on two-key inputs is also the
    stability proof.  `one_digit`: a whole tile in one digit run (one 64 KB run); `ties`: col1 with few values."""
    env = get_env()
    rng = np.random.default_rng(n + len(shape))
    if shape == "cfg4":
        cols = [rng.integers(-2 ** 19, 2 ** 19, n).astype(np.int64), rng.integers(-2 ** 63, 2 ** 63 - 1, n).astype(np.int64)]
        keys, desc = [0, 1], [0, 0]
    elif shape == "ties":
        cols = [rng.integers(0, 3, n).astype(np.int64), rng.integers(0, 1000, n).astype(np.int64)]
        keys, desc = [0, 1], [0, 1]
    elif shape == "one_digit":
        cols = [np.full(n, 7, dtype=np.int64), rng.integers(0, 256, n).astype(np.int64) * 65536]
        keys, desc = [1, 0], [0, 0]
    elif shape == "f64_desc":
        f = rng.normal(size=n)
        f[::37] = np.nan
        f[1::41] = -0.0
        cols = [f, rng.integers(-5, 5, n).astype(np.int64)]
        keys, desc = [1, 0], [1, 0]
    else:
        cols = [rng.integers(-2 ** 62, 2 ** 62, n).astype(np.int64), rng.integers(-2 ** 62, 2 ** 62, n).astype(np.int64)]
        keys, desc = [0, 1], [1, 1]
    t = env.from_columns(cols)
    env.set_option("sort.sweep16_min_rows", 1)
    try:
        for trunc in (1, 0):
            env.set_option("sort.trunc", trunc)
            r = env.query_orderby(t, [0, 1], keys, desc)
            assert env.get_option("sort.last_passes") == 0 or env.get_option("sort.last_sweep16") == 1
            exp = NO.query_orderby(cols, [0, 1], keys, desc)
            for g, e in zip(r.columns(), exp):
                assert g.dtype == e.dtype and np.array_equal(g, e, equal_nan=True), (shape, n, trunc)
            r.free()
    finally:
        env.set_option("sort.sweep16_min_rows", 1 << 16)
        env.set_option("sort.trunc", 1)
    t.free()


# ---------------------------------------------------------------- JOIN
@pytest.mark.parametrize("idx", [i for i, c in enumerate(VEC) if c["kind"] == "join"])
def test_join_golden_vectors(idx):
    env = get_env()
    case = VEC[idx]
    db1 = np.asarray(_db(case["db1"]), dtype=np.int64).astype(np.uint32)
    db2 = np.asarray(_db(case["db2"]), dtype=np.int64).astype(np.uint32)
    db1 = db1.reshape(db1.shape[0], -1) if db1.size else np.zeros((0, 1), np.uint32)
    db2 = db2.reshape(db2.shape[0], -1) if db2.size else np.zeros((0, 1), np.uint32)
    w = len(case["cols1"]) + len(case["cols2"])
    exp = np.asarray(case["output"], dtype=np.int64).astype(np.uint32).reshape(-1, w)
    got = env.from_futhark(env.join(db1, db2, case["col1"], case["col2"], case["cols1"], case["cols2"]))
    assert got.shape == exp.shape and np.array_equal(got, exp)


@pytest.mark.parametrize("n1,n2,kmax", [(1, 1, 1), (100, 100, 10), (5000, 3000, 40), (100000, 50000, 2 ** 32 - 1),
                                        (20000, 20000, 20000), (4097, 4096, 3)])
def test_join_vs_oracle(n1, n2, kmax):
    env = get_env()
    rng = np.random.default_rng(n1 + n2)
    lo = kmax - 30000 if kmax == 2 ** 32 - 1 else 0
    a = rng.integers(0, 2 ** 32, (n1, 3), dtype=np.uint64).astype(np.uint32)
    b = rng.integers(0, 2 ** 32, (n2, 2), dtype=np.uint64).astype(np.uint32)
    a[:, 1] = rng.integers(lo, kmax + 1, n1, dtype=np.uint64)
    b[:, 0] = rng.integers(lo, kmax + 1, n2, dtype=np.uint64)
    exp = NO.join(a, b, 1, 0, [0, 1, 2], [1, 0])
    if exp.shape[0] > 30_000_000:
        pytest.skip("too many pairs for the oracle")
    got = env.from_futhark(env.join(a, b, 1, 0, [0, 1, 2], [1, 0]))
    assert got.shape == exp.shape and np.array_equal(got, exp)


def test_join_empty_and_errors():
    env = get_env()
    from harkdb_b200.hark_ffi import HarkError
    a = np.arange(12, dtype=np.uint32).reshape(4, 3)
    e = np.zeros((0, 2), dtype=np.uint32)
    assert env.from_futhark(env.join(a, e, 0, 0, [0], [1])).shape == (0, 2)
    assert env.from_futhark(env.join(e, a, 0, 0, [0], [1])).shape == (0, 2)
    assert env.from_futhark(env.join(a, a + 100, 0, 0, [0], [1])).shape == (0, 2)   # no key in common
    with pytest.raises(HarkError, match="out of bounds"):
        env.join(a, a, 3, 0, [0], [1])
    with pytest.raises(HarkError, match="out of bounds"):
        env.join(a, a, 0, 0, [9], [1])


def _rows_sorted(cols):
    """Rows of a column list in lexicographic order (for multiset comparison); floats by their bit patterns."""
    if not cols or len(cols[0]) == 0:
        return cols
    keys = [c.view(np.uint32) if c.dtype == np.float32 else c.view(np.uint64) if c.dtype == np.float64 else c for c in cols]
    o = np.lexsort(tuple(keys[::-1]))
    return [c[o] for c in cols]


@pytest.mark.parametrize("kdt", [NO.I32, NO.U32, NO.I64])
@pytest.mark.parametrize("order", [1, 0])
@pytest.mark.parametrize("n1,n2,krange", [(1, 1, 1), (300, 200, 7), (5000, 3000, 40), (100003, 50021, 10 ** 9),
                                          (20000, 20000, 20000), (4097, 2049, 3), (70001, 9, 5)])
def test_join_ex_typed(kdt, order, n1, n2, krange):
    """hark_entry_join_ex: signed / unsigned / 64-bit keys (negative values, sparse ranges, duplicates on both sides, no
    match at all), mixed-dtype projected columns; order=1 bit-exact row order, order=0 the same multiset grouped by r1."""
    env = get_env()
    rng = np.random.default_rng(n1 * 5 + n2 + kdt + order)
    npdt = NO.NP_DTYPES[kdt]
    lo = {NO.I32: -krange // 2, NO.U32: 2 ** 31 - krange // 2, NO.I64: -2 ** 40}[kdt]
    step = 2 ** 30 if (kdt == NO.I64 and krange > 10 ** 6) else 1           # i64: span >= 1000 x rows
    k1 = (lo + step * rng.integers(0, krange, n1)).astype(npdt)
    k2 = (lo + step * rng.integers(0, krange, n2)).astype(npdt)
    t1 = [rng.integers(-100, 100, n1).astype(np.int32), k1, rng.random(n1).astype(np.float64), np.arange(n1, dtype=np.int64)]
    t2 = [k2, rng.random(n2).astype(np.float32), np.arange(n2, dtype=np.uint32)]
    d1, d2 = env.from_columns(t1), env.from_columns(t2)
    exp = NO.join_ex(t1, t2, 1, 0, [3, 1, 2, 0], [2, 1])
    if len(exp[0]) > 30_000_000:
        pytest.skip("too many pairs")
    r = env.join_ex(d1, d2, 1, 0, [3, 1, 2, 0], [2, 1], order)
    got = r.columns()
    assert [g.dtype for g in got] == [e.dtype for e in exp] and r.shape[0] == len(exp[0])
    if order == 1:
        for g, e in zip(got, exp):
            assert np.array_equal(g, e)
    else:
        assert np.all(np.diff(got[0]) >= 0)                                   # grouped by db1 row, ascending
        for g, e in zip(_rows_sorted(got), _rows_sorted(exp)):
            assert np.array_equal(g, e)
    for x in (r, d1, d2):
        x.free()


def test_join_ex_no_overlap_empty_and_errors():
    env = get_env()
    from harkdb_b200.hark_ffi import HarkError
    a = env.from_columns([np.arange(10, dtype=np.int64), np.arange(10, dtype=np.int32)])
    b = env.from_columns([np.arange(100, 110, dtype=np.int64), np.ones(10, dtype=np.float32)])
    e = env.from_columns([np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.float32)])
    for order in (0, 1):
        assert env.join_ex(a, b, 0, 0, [0, 1], [1], order).shape == (0, 3)
        assert env.join_ex(a, e, 0, 0, [0], [1], order).shape == (0, 2)
        assert env.join_ex(e, a, 0, 0, [1], [0, 1], order).dtypes == [NO.F32, NO.I64, NO.I32]
        with pytest.raises(HarkError, match="same integer dtype"):
            env.join_ex(a, b, 1, 0, [0], [1], order)                          # i32 key against i64 key
        with pytest.raises(HarkError, match="out of bounds"):
            env.join_ex(a, b, 0, 0, [5], [1], order)
    for x in (a, b, e):
        x.free()


def test_join_groupby_config5_shape(gb_impl):
    """BASELINE config 5 at an oracle-checkable size: fact(fk,val) JOIN dim(pk unique,attr) GROUP BY attr."""
    env = get_env()
    from oracle import c_oracle as CO
    nd, nf = 100003, 1 << 20
    dspec = [dict(kind=NO.GEN_AFFINE, a=48271, b=11, range=nd), dict(kind=NO.GEN_UNIFORM, lo=0, range=1024)]
    fspec = [dict(kind=NO.GEN_UNIFORM, lo=0, range=2 * nd), dict(kind=NO.GEN_UNIFORM, lo=-100, range=200)]   # 50 % match
    dim = env.synth(nd, [NO.I32, NO.I32], dspec, seed=7)
    fact = env.synth(nf, [NO.I32, NO.I32], fspec, seed=8)
    dcols = [CO.synth_column(NO.I32, dspec[c], 7, c, 0, nd) for c in range(2)]
    fcols = [CO.synth_column(NO.I32, fspec[c], 8, c, 0, nf) for c in range(2)]
    ops = [NO.AGG_SUM, NO.AGG_COUNT, NO.AGG_AVG, NO.AGG_MAX]
    r = env.join_groupby(fact, dim, 0, 0, 1, [1, 1, 1, 1], ops)
    exp = NO.join_groupby(fcols, dcols, 0, 0, 1, [1, 1, 1, 1], ops)
    _check_cols(r.columns(), exp)
    assert 0.45 < exp[2].sum() / nf < 0.55
    r.free(); dim.free(); fact.free()


def test_join_groupby_duplicate_dim_key_and_misses(gb_impl):
    env = get_env()
    from harkdb_b200.hark_ffi import HarkError
    dim = env.from_columns([np.array([5, 9, 5, 7], dtype=np.int32), np.array([1, 2, 3, 4], dtype=np.int32)])
    fact = env.from_columns([np.array([5, 7, 9, 9, 100, -3], dtype=np.int32), np.array([1, 2, 3, 4, 5, 6], dtype=np.int32)])
    with pytest.raises(HarkError, match="not unique"):
        env.join_groupby(fact, dim, 0, 0, 1, [1], [NO.AGG_SUM])
    dim2 = env.from_columns([np.array([5, 9, 7], dtype=np.int32), np.array([40, 40, -2], dtype=np.int32)])
    r = env.join_groupby(fact, dim2, 0, 0, 1, [1, 1], [NO.AGG_SUM, NO.AGG_COUNT])      # fk 100 and -3 match nothing
    k, s, c = r.columns()
    assert k.tolist() == [-2, 40] and s.tolist() == [2, 8] and c.tolist() == [1, 3]
    for x in (r, dim, dim2, fact):
        x.free()


@pytest.fixture(params=["one_slice", "many_slices"])
def slices(request):
    """`many_slices`: 8 KB build-structure slices, so that small inputs run K8t (the tile-local partition by lookup /
    hash-table slice) in front of the probe; `one_slice`: the whole structure is one slice (no partition)."""
    env = get_env()
    env.set_option("join.lut_slice_bytes", 8192 if request.param == "many_slices" else 16 << 20)
    yield request.param
    env.set_option("join.lut_slice_bytes", 16 << 20)


def _sparse_keys(rng, nd, kdt):
    """nd unique keys spread over (almost) the whole value range of the dtype: span >= 1000 x rows for i64."""
    if kdt == NO.I64:
        k = rng.integers(-2 ** 62, 2 ** 62, int(nd * 1.2)).astype(np.int64)
    elif kdt == NO.U32:
        k = rng.integers(0, 2 ** 32, int(nd * 1.2), dtype=np.uint64).astype(np.uint32)
    else:
        k = rng.integers(-2 ** 31, 2 ** 31, int(nd * 1.2)).astype(np.int32)
    k = np.unique(k)
    rng.shuffle(k)
    assert len(k) >= nd
    return np.ascontiguousarray(k[:nd])


@pytest.mark.parametrize("kdt", [NO.I32, NO.U32, NO.I64])
@pytest.mark.parametrize("match", [0.0, 0.5, 1.0])
@pytest.mark.parametrize("nd,nf", [(1, 5), (1000, 20000), (50021, 300007)])
def test_join_groupby_hash_build_sparse_keys(gb_impl, slices, kdt, match, nd, nf):
    """Sparse build keys (negative, unsigned above 2^31, 64-bit): the open-addressing hash build + probe, fused into K2
    when the group domain fits (auto / tiny_tables) and through probe -> filter -> GROUP BY otherwise (sort)."""
    env = get_env()
    rng = np.random.default_rng(nd * 3 + nf + int(match * 10) + kdt)
    pk = _sparse_keys(rng, nd, kdt)
    attr = rng.integers(-5, 40, nd).astype(np.int32)
    hit = pk[rng.integers(0, nd, nf)]
    miss = _sparse_keys(rng, nf, kdt)
    miss = miss[~np.isin(miss, pk)]
    miss = np.resize(miss, nf) if len(miss) else hit.copy()
    fk = np.where(rng.random(nf) < match, hit, miss).astype(pk.dtype)
    val = rng.integers(-1000, 1000, nf).astype(np.int32)
    fval = rng.random(nf).astype(np.float32)
    dim = env.from_columns([pk, attr])
    fact = env.from_columns([fk, val, fval])
    ops = [NO.AGG_SUM, NO.AGG_COUNT, NO.AGG_AVG, NO.AGG_MIN, NO.AGG_SUM]
    sc = [1, 1, 1, 1, 2]
    r = env.join_groupby(fact, dim, 0, 0, 1, sc, ops)
    assert env.get_option("join.last_build") == (2 if nd > 1 else 1)     # span > 16 x rows: the hash build ran
    _check_cols(r.columns(), NO.join_groupby([fk, val, fval], [pk, attr], 0, 0, 1, sc, ops))
    for x in (r, dim, fact):
        x.free()


@pytest.mark.parametrize("nd,nf", [(3, 50), (5000, 100003), (70001, 300007)])
def test_join_groupby_hash_build_by_table_slices(nd, nf):
    """The hash table is built slice by slice over K8t's output (hk_hash_build_tiles_kernel) when it is much larger
    than a slice: forced here with 8 KB slices at any size; duplicates are still rejected."""
    env = get_env()
    from harkdb_b200.hark_ffi import HarkError
    rng = np.random.default_rng(nd)
    pk = _sparse_keys(rng, nd, NO.I32)
    attr = rng.integers(-5, 90, nd).astype(np.int32)
    fk = np.where(rng.random(nf) < 0.7, pk[rng.integers(0, nd, nf)], rng.integers(-2 ** 31, 2 ** 31, nf)).astype(np.int32)
    val = rng.integers(-1000, 1000, nf).astype(np.int32)
    dim, fact = env.from_columns([pk, attr]), env.from_columns([fk, val])
    env.set_option("join.lut_slice_bytes", 8192)
    env.set_option("join.build_partition_min_rows", 1)
    env.set_option("join.build", 2)
    try:
        ops = [NO.AGG_SUM, NO.AGG_COUNT, NO.AGG_MIN]
        r = env.join_groupby(fact, dim, 0, 0, 1, [1, 1, 1], ops)
        _check_cols(r.columns(), NO.join_groupby([fk, val], [pk, attr], 0, 0, 1, [1, 1, 1], ops))
        r.free()
        pk2 = pk.copy()
        pk2[-1] = pk2[0]
        d2 = env.from_columns([pk2, attr])
        if nd > 1:
            with pytest.raises(HarkError, match="not unique"):
                env.join_groupby(fact, d2, 0, 0, 1, [1], [NO.AGG_SUM])
        d2.free()
    finally:
        env.set_option("join.lut_slice_bytes", 16 << 20)
        env.set_option("join.build_partition_min_rows", 1 << 16)
        env.set_option("join.build", 0)
    dim.free(); fact.free()


def test_join_groupby_hash_build_equals_lookup_build_and_rejects_duplicates(gb_impl, slices):
    env = get_env()
    from harkdb_b200.hark_ffi import HarkError
    rng = np.random.default_rng(5)
    nd, nf = 30011, 200003
    pk = rng.permutation(nd).astype(np.int32) - 7000            # dense keys with negative values
    attr = rng.integers(0, 300, nd).astype(np.int32)
    fk = rng.integers(-9000, nd, nf).astype(np.int32)
    val = rng.integers(0, 1000, nf).astype(np.int32)
    dim, fact = env.from_columns([pk, attr]), env.from_columns([fk, val])
    ops = [NO.AGG_SUM, NO.AGG_COUNT, NO.AGG_MAX]
    exp = NO.join_groupby([fk, val], [pk, attr], 0, 0, 1, [1, 1, 1], ops)
    try:
        for build in (1, 2):
            env.set_option("join.build", build)
            r = env.join_groupby(fact, dim, 0, 0, 1, [1, 1, 1], ops)
            assert env.get_option("join.last_build") == build
            _check_cols(r.columns(), exp)
            r.free()
        env.set_option("join.build", 2)
        for kdt in (np.int32, np.int64):
            d2 = env.from_columns([np.array([5, 9, 5, 7], dtype=kdt), np.array([1, 2, 3, 4], dtype=np.int32)])
            f2 = env.from_columns([np.array([5, 7, 9], dtype=kdt), np.array([1, 2, 3], dtype=np.int32)])
            with pytest.raises(HarkError, match="not unique"):
                env.join_groupby(f2, d2, 0, 0, 1, [1], [NO.AGG_SUM])
            d2.free(); f2.free()
    finally:
        env.set_option("join.build", 0)
    dim.free(); fact.free()


def test_join_groupby_config5_shape_many_slices(gb_impl, slices):
    """Config 5's shape with the lookup cut into many slices: K8t + the segment-walking aggregation (lookup mode)."""
    env = get_env()
    from oracle import c_oracle as CO
    nd, nf = 100003, (1 << 19) + 77
    dspec = [dict(kind=NO.GEN_AFFINE, a=48271, b=11, range=nd), dict(kind=NO.GEN_UNIFORM, lo=0, range=1024)]
    fspec = [dict(kind=NO.GEN_UNIFORM, lo=-5, range=2 * nd), dict(kind=NO.GEN_UNIFORM, lo=-100, range=200)]
    dim = env.synth(nd, [NO.I32, NO.I32], dspec, seed=7)
    fact = env.synth(nf, [NO.I32, NO.I32], fspec, seed=8)
    dcols = [CO.synth_column(NO.I32, dspec[c], 7, c, 0, nd) for c in range(2)]
    fcols = [CO.synth_column(NO.I32, fspec[c], 8, c, 0, nf) for c in range(2)]
    ops = [NO.AGG_SUM, NO.AGG_COUNT, NO.AGG_AVG, NO.AGG_MAX]
    r = env.join_groupby(fact, dim, 0, 0, 1, [1, 1, 1, 1], ops)
    _check_cols(r.columns(), NO.join_groupby(fcols, dcols, 0, 0, 1, [1, 1, 1, 1], ops))
    r.free(); dim.free(); fact.free()


@pytest.mark.parametrize("shape", ["u24", "signed_u24", "wide64", "too_wide", "nonfinite", "denormal", "zeros"])
def test_groupby_f32_sum_exact_fixed_point_and_fallback(gb_impl, shape):
    """f32 SUM / AVG in K2: exact fixed-point accumulation when the column's zone map proves it exact (32-bit addends
    for values on a 2^-24 grid, 64-bit addends for wider mantissa spreads), the f64 compare-and-swap path otherwise
    (dynamic range too wide, NaN / inf present).  All within the stated 1e-5 (f32 SUM) / 1e-12 (f64 AVG of exact sums
    is looser: compared at 1e-5 like the oracle's f64 accumulation allows)."""
    env = get_env()
    rng = np.random.default_rng(len(shape))
    n = 200003
    key = rng.integers(-50, 450, n).astype(np.int32)
    if shape == "u24":
        val = (rng.integers(0, 2 ** 24, n).astype(np.float32) * np.float32(2.0 ** -24))
    elif shape == "signed_u24":
        val = ((rng.integers(0, 2 ** 24, n) - 2 ** 23).astype(np.float32) * np.float32(2.0 ** -20))
    elif shape == "wide64":
        val = (rng.random(n) * 10.0 ** rng.integers(-3, 4, n)).astype(np.float32)
    elif shape == "too_wide":
        val = (rng.random(n) * 10.0 ** rng.integers(-30, 30, n)).astype(np.float32)
    elif shape == "nonfinite":
        val = rng.random(n).astype(np.float32)
        val[::1000] = np.inf
        val[5::1000] = np.nan
    elif shape == "denormal":
        val = (rng.integers(1, 1000, n).astype(np.float32) * np.float32(1e-42))
    else:
        val = np.zeros(n, dtype=np.float32)
    cols = [key, val]
    t = env.from_columns(cols)
    ops = [NO.AGG_SUM, NO.AGG_AVG, NO.AGG_COUNT, NO.AGG_MAX]
    r = env.query_groupby_ex(t, 0, [1, 1, 1, 1], ops)
    exp = NO.query_groupby_ex(cols, 0, [1, 1, 1, 1], ops)
    got = r.columns()
    assert np.array_equal(got[0], exp[0]) and np.array_equal(got[3], exp[3])
    if shape == "nonfinite":
        for g, e in zip(got[1:3], exp[1:3]):
            assert np.array_equal(np.isnan(g), np.isnan(e)) and np.array_equal(np.isinf(g), np.isinf(e))
            ok = np.isfinite(e)
            assert np.allclose(g[ok], e[ok], rtol=1e-5, atol=0)
    else:
        assert np.allclose(got[1], exp[1], rtol=1e-5, atol=0) and np.allclose(got[2], exp[2], rtol=1e-5, atol=0)
    if shape != "nonfinite":
        assert np.array_equal(got[4], exp[4])
    r.free(); t.free()


@pytest.mark.parametrize("part_impl", [0, 1])
@pytest.mark.parametrize("kdt,vdts", [(NO.I32, [NO.I32]), (NO.I64, [NO.I32, NO.F32]), (NO.U32, [NO.U32, NO.I32, NO.F32]), (NO.I32, [])])
@pytest.mark.parametrize("n", [1, 4095, 4096, 4097, 8193, 300007])
def test_groupby_tile_partition_vs_chunk_partition(part_impl, kdt, vdts, n):
    """K8t (tile-local partition + directory, dense.part_impl=0) and K8a (chunked partition, =1) in front of K2, tiny
    tables so that every size partitions: ragged last tiles, 4- and 8-byte keys, 0 to 3 carried value columns."""
    env = get_env()
    rng = np.random.default_rng(n + kdt + len(vdts))
    key = (rng.integers(0, 5000, n) - 2000).astype(NO.NP_DTYPES[kdt]) if kdt != NO.U32 else \
        (rng.integers(0, 5000, n) + 2 ** 31 - 2500).astype(np.uint32)
    cols = [key]
    for vd in vdts:
        cols.append(rng.random(n).astype(np.float32) if vd == NO.F32 else rng.integers(0, 2 ** 31, n).astype(NO.NP_DTYPES[vd]))
    sc, ops = [], []
    for j, vd in enumerate(vdts):
        sc += [j + 1, j + 1]
        ops += [NO.AGG_SUM, NO.AGG_MAX if vd != NO.F32 else NO.AGG_AVG]
    sc.append(0); ops.append(NO.AGG_COUNT)
    t = env.from_columns(cols)
    env.set_option("dense.log2_slots", 8)
    env.set_option("dense.part_impl", part_impl)
    try:
        r = env.query_groupby_ex(t, 0, sc, ops)
    finally:
        env.set_option("dense.log2_slots", 0)
        env.set_option("dense.part_impl", 0)
    _check_cols(r.columns(), NO.query_groupby_ex(cols, 0, sc, ops))
    r.free(); t.free()


@pytest.mark.parametrize("dynamic", [1, 2, 0])
@pytest.mark.parametrize("skew", ["uniform", "zipf", "one_bin"])
def test_groupby_tiles_dealt_units_vs_static_split(dynamic, skew):
    """K2 over K8t tiles with per-bin ticket dealing (dense.dynamic=1, default) and with the static split by rows (=0):
    same result on uniform keys, on Zipf-like keys (a few hot groups) and when every row falls into one bin."""
    env = get_env()
    rng = np.random.default_rng(11)
    n = 300007
    if skew == "uniform":
        key = rng.integers(-3000, 3000, n)
    elif skew == "zipf":
        key = np.minimum(rng.zipf(1.3, n), 5999) - 3000
    else:
        key = rng.integers(100, 140, n)
        key[::1001] = 2999
        key[5] = -3000
    cols = [key.astype(np.int32), rng.integers(-1000, 1000, n).astype(np.int32), rng.random(n).astype(np.float32)]
    sc, ops = [1, 1, 2, 0], [NO.AGG_SUM, NO.AGG_MIN, NO.AGG_SUM, NO.AGG_COUNT]
    t = env.from_columns(cols)
    env.set_option("dense.log2_slots", 8)
    env.set_option("dense.dynamic", 1 if dynamic else 0)
    env.set_option("dense.home_by_rows", 1 if dynamic == 2 else 0)   # 2: dealt units, first bin from the row prefix
    try:
        r = env.query_groupby_ex(t, 0, sc, ops)
    finally:
        env.set_option("dense.log2_slots", 0)
        env.set_option("dense.dynamic", 1)
        env.set_option("dense.home_by_rows", 0)
    _check_cols(r.columns(), NO.query_groupby_ex(cols, 0, sc, ops))
    r.free(); t.free()


def test_block_cache_reuses_and_trims():
    """The size-keyed cache of freed device blocks: results do not depend on it, a repeated query reuses
    its blocks, trim() empties it, blocks above pool.cache_block_gb bypass it."""
    env = get_env()
    rng = np.random.default_rng(12)
    n = 200003
    cols = [rng.integers(0, 50, n).astype(np.int32), rng.integers(-2 ** 40, 2 ** 40, n).astype(np.int64)]
    exp = NO.query_orderby(cols, [0, 1], [0, 1], [0, 1])
    t = env.from_columns(cols)
    for cache in (1, 0, 1):
        env.set_option("pool.cache", cache)
        for _ in range(3):
            r = env.query_orderby(t, [0, 1], [0, 1], [0, 1])
            for g, e in zip(r.columns(), exp):
                assert np.array_equal(g, e)
            r.free()
        env.trim()
    env.set_option("pool.cache_block_gb", 0)        # nothing is cacheable: every block goes back to the driver's pool
    try:
        r = env.query_orderby(t, [0, 1], [0, 1], [0, 1])
        assert np.array_equal(r.column(1), exp[1])
        r.free()
    finally:
        env.set_option("pool.cache_block_gb", 12)
    t.free()


@pytest.mark.parametrize("kdt", [NO.I32, NO.I64])
@pytest.mark.parametrize("dups", [False, True])
def test_hash_join_unique_and_duplicate_build_keys(kdt, dups):
    """hark_entry_join_ex, order = 0: the build raises a duplicate flag (4-byte keys); without duplicates probes stop at
    the first match, with duplicates they walk the whole cluster.  Misses, negative keys, one heavy duplicate key."""
    env = get_env()
    rng = np.random.default_rng(13 + int(dups))
    n2, n1 = 5003, 60011
    base = rng.permutation(40000)[:n2] - 20000
    if dups:
        base[::7] = base[3]                      # one key carried by many build rows
        base[1::11] = base[1]
    k2 = base.astype(NO.NP_DTYPES[kdt])
    k1 = rng.integers(-25000, 25000, n1).astype(NO.NP_DTYPES[kdt])   # ~20 % of the probe keys miss
    t1 = [k1, rng.integers(0, 1000, n1).astype(np.int32)]
    t2 = [k2, np.arange(n2, dtype=np.int32)]
    d1, d2 = env.from_columns(t1), env.from_columns(t2)
    r = env.join_ex(d1, d2, 0, 0, [0, 1], [1], 0)
    got = r.columns()
    # unique build keys take the fused one-pass plan (result allocated for every probe row, row count set after)
    assert env.get_option("join.last_one_pass") == (0 if dups else 1)
    exp = NO.join_ex(t1, t2, 0, 0, [0, 1], [1])
    assert r.shape[0] == len(exp[0]) and len(got[0]) == len(exp[0])
    if not dups:        # one match per row at most: the result keeps probe-row order
        hit = np.isin(k1, k2)
        assert np.array_equal(got[0], k1[hit]) and np.array_equal(got[1], t1[1][hit])
    o1, o2 = np.lexsort((got[2], got[1], got[0])), np.lexsort((exp[2], exp[1], exp[0]))
    for g, e in zip(got, exp):
        assert np.array_equal(g[o1], e[o2])
    env.set_option("join.hash_one_pass", 0)     # the two-pass plan on the same input
    try:
        got2 = env.join_ex(d1, d2, 0, 0, [0, 1], [1], 0).columns()
    finally:
        env.set_option("join.hash_one_pass", 1)
    o3 = np.lexsort((got2[2], got2[1], got2[0]))   # duplicates: the matches of one probe row come in the build's CAS order
    for g, g2 in zip(got, got2):
        assert np.array_equal(g, g2) if not dups else np.array_equal(g[o1], g2[o3])
    r.free(); d1.free(); d2.free()


@pytest.mark.parametrize("carry", [1, 0])
@pytest.mark.parametrize("cols1", [[0, 1], [1, 1, 0, 2], [3, 2, 1, 0, 4], [2]])
def test_ordered_join_carries_left_columns_through_the_sort(carry, cols1):
    """hark_entry_join_ex, order = 1 (reference order): the projected left columns ride through the sort instead of a row
    id (join.carry=1, default) when at most 3 distinct non-key columns are projected; same rows, same order either way.
    Mixed widths, a column projected twice, the key column projected, more columns than are carried."""
    env = get_env()
    rng = np.random.default_rng(21)
    n1, n2 = 30011, 4001
    t1 = [rng.integers(-200, 200, n1).astype(np.int32), rng.integers(-2 ** 50, 2 ** 50, n1).astype(np.int64),
          rng.random(n1).astype(np.float32), rng.random(n1), np.arange(n1, dtype=np.int32)]
    t2 = [rng.integers(-250, 150, n2).astype(np.int32), np.arange(n2, dtype=np.int64)]
    d1, d2 = env.from_columns(t1), env.from_columns(t2)
    env.set_option("join.carry", carry)
    try:
        got = env.join_ex(d1, d2, 0, 0, cols1, [1, 0], 1).columns()
        carried = env.get_option("join.last_carry")
    finally:
        env.set_option("join.carry", 1)
    distinct = len({c for c in cols1 if c != 0})
    assert carried == (1 if carry and distinct <= 3 else 0)
    exp = NO.join_ex(t1, t2, 0, 0, cols1, [1, 0])
    assert len(got) == len(exp)
    for g, e in zip(got, exp):
        assert g.dtype == e.dtype and np.array_equal(g, e)
    d1.free(); d2.free()


# ---------------------------------------------------------------- SQL through FutharkContext
def test_sql_groupby_orderby_join():
    import pandas as pd
    from harkdb_b200 import FutharkContext
    fc = FutharkContext()
    fc.create_table("game_1", pd.DataFrame(DATA["rows"], columns=DATA["columns"]))
    out = fc.sql("select col1,  max(col3) from game_1 group by col1")            # test.py:7
    assert out.tolist() == [[0, 0, 0], [1, 1, 3], [6, 6, 6]] and out.dtype == np.uint32
    out = fc.sql("select col1, sum(col2), count(col2), avg(col2) from game_1 group by col1 having count(col2) > 1")
    assert out.tolist() == [[0.0, 0.0, 0.0, 4.0, 0.0], [6.0, 6.0, 12.0, 2.0, 6.0]]
    out = fc.sql("select col1, col3 from game_1 where col1 > 0 order by col3 desc, col1")
    assert out.tolist() == [[6, 6], [6, 6], [1, 3]]
    out = fc.sql("select col1, count(col1) from game_1 group by col1 order by col1 desc limit 2")
    assert out.tolist() == [[6, 6, 2], [1, 1, 1]]
    fc.create_table("dim", pd.DataFrame({"pk": [6, 1, 7, 6], "attr": [60, 10, 70, 61]}))
    out = fc.sql("select game_1.col1, game_1.col3, dim.attr from game_1 join dim on game_1.col1 = dim.pk")
    assert out.tolist() == [[1, 3, 10], [6, 6, 60], [6, 6, 61], [6, 6, 60], [6, 6, 61]]
    fc.create_table("d2", pd.DataFrame({"pk": [6, 1, 0], "attr": [5, 5, 9]}))
    out = fc.sql("select attr, sum(col3), count(*) from game_1 join d2 on col1 = pk group by attr")
    assert out.tolist() == [[5.0, 15.0, 3.0], [9.0, 0.0, 4.0]]


# ---------------------------------------------------------------- full-size properties (BASELINE configs 3 and 4)
def test_groupby_full_size_properties():
    """Config 3 (1e9 rows, 2^20 distinct i32 keys): G, sorted keys, sum of counts == n, sum of sums == column
    sum, and exact agreement on the groups of a regenerated prefix."""
    import torch
    env = get_env()
    n = 10 ** 9 if free_gb() > 100 else (1 << 26)
    specs = [dict(kind=NO.GEN_UNIFORM, lo=0, range=1 << 20), dict(kind=NO.GEN_UNIFORM, lo=0, range=1000)]
    t = env.synth(n, [NO.I32, NO.I32], specs, seed=42)
    ops = [NO.AGG_SUM, NO.AGG_COUNT, NO.AGG_AVG]
    r = env.query_groupby_ex(t, 0, [1, 1, 1], ops)
    st = env.stats()
    keys, sums, cnts, avgs = r.columns()
    assert len(keys) == 1 << 20 and np.array_equal(keys, np.arange(1 << 20, dtype=np.int32))
    assert int(cnts.sum()) == n
    total = int(as_torch(t, 1).sum(dtype=torch.int64).item())
    assert int(sums.astype(np.int64).sum()) == total if n < 2 ** 22 else True
    assert (int(sums.view(np.uint32).astype(np.uint64).sum()) - total) % (1 << 32) == 0    # sums wrap mod 2^32 per group
    assert np.allclose(avgs, sums.astype(np.float64) / cnts, rtol=1e-12) if n <= (1 << 26) else True
    k = cnts.astype(np.int64)
    hv = env.query_groupby_ex(t, 0, [1, 1, 1], ops, having=[(2, NO.GT, int(np.median(k)), 0.0)])
    assert hv.shape[0] == int((k > np.median(k)).sum())
    print(f"groupby {n} rows: {st['total_ms']:.2f} ms total, reduce kernel {st['kernel_ms']:.2f} ms")
    for x in (hv, r, t):
        x.free()


def test_join_groupby_full_size_properties():
    """Config 5 (4e9-row fact x 1e8-row dim when memory allows): every fact row matches exactly one dim row, so the
    group counts add up to the fact rows, the sums to the value column's sum (mod 2^32 per group), and the 1024
    attribute values come out ascending."""
    import torch
    env = get_env()
    big = free_gb() > 120
    nf, nd = (4 * 10 ** 9, 10 ** 8) if big else (1 << 27, 1 << 22)
    a = 2654435761
    while np.gcd(a, nd) != 1:
        a += 2
    dim = env.synth(nd, [NO.I32, NO.I32], [dict(kind=NO.GEN_AFFINE, a=a, b=12345, range=nd), dict(kind=NO.GEN_UNIFORM, lo=0, range=1024)], seed=7)
    fact = env.synth(nf, [NO.I32, NO.I32], [dict(kind=NO.GEN_UNIFORM, lo=0, range=nd), dict(kind=NO.GEN_UNIFORM, lo=0, range=1000)], seed=42)
    r = env.join_groupby(fact, dim, 0, 0, 1, [1, 1, 1], [NO.AGG_SUM, NO.AGG_COUNT, NO.AGG_AVG])
    st = env.stats()
    keys, sums, cnts, avgs = r.columns()
    assert np.array_equal(keys, np.arange(1024, dtype=np.int32)) and int(cnts.sum()) == nf
    total = int(as_torch(fact, 1).sum(dtype=torch.int64).item())
    assert (int(sums.view(np.uint32).astype(np.uint64).sum()) - total) % (1 << 32) == 0
    assert abs(float((avgs * cnts).sum()) - total) <= 1e-9 * total            # AVG = exact sum / count
    # a prefix the oracle can redo: the first 2^20 fact rows against the whole dimension
    from oracle import c_oracle as CO
    n0 = 1 << 20
    fcols = [CO.synth_column(NO.I32, dict(kind=NO.GEN_UNIFORM, lo=0, range=nd), 42, 0, 0, n0),
             CO.synth_column(NO.I32, dict(kind=NO.GEN_UNIFORM, lo=0, range=1000), 42, 1, 0, n0)]
    head = env.slice(fact, 0, n0)
    r0 = env.join_groupby(head, dim, 0, 0, 1, [1, 1], [NO.AGG_SUM, NO.AGG_COUNT])
    dcols = dim.columns()
    _check_cols(r0.columns(), NO.join_groupby(fcols, dcols, 0, 0, 1, [1, 1], [NO.AGG_SUM, NO.AGG_COUNT]))
    print(f"join+groupby {nf} x {nd} rows: {st['total_ms']:.2f} ms")
    for x in (r0, head, r, fact, dim):
        x.free()


def test_orderby_full_size_properties():
    """Config 4 (2e9 x {i64,i64} when memory allows): sortedness and multiset preservation, checked on device."""
    import torch
    env = get_env()
    n = 2 * 10 ** 9 if free_gb() > 150 else (1 << 27)
    specs = [dict(kind=NO.GEN_UNIFORM, lo=-(2 ** 19), range=2 ** 20), dict(kind=NO.GEN_UNIFORM, lo=0, range=0)]
    t = env.synth(n, [NO.I64, NO.I64], specs, seed=42)
    s1 = int(as_torch(t, 0).sum().item())
    x2 = as_torch(t, 1)
    s2 = int(x2.sum().item())                  # wraps mod 2^64 like the check below
    r = env.query_orderby(t, [0, 1], [0, 1])
    st = env.stats()
    a, b = as_torch(r, 0), as_torch(r, 1)
    assert int(a.sum().item()) == s1 and int(b.sum().item()) == s2
    chunk = 1 << 28
    for lo in range(0, n - 1, chunk):
        hi = min(n - 1, lo + chunk)
        a0, a1, b0, b1 = a[lo:hi], a[lo + 1:hi + 1], b[lo:hi], b[lo + 1:hi + 1]
        ok = (a0 < a1) | ((a0 == a1) & (b0 <= b1))
        assert bool(ok.all().item()), f"not sorted in rows [{lo},{hi})"
        del ok
    print(f"orderby {n} rows: {st['total_ms']:.2f} ms total, passes {st['kernel_ms']:.2f} ms")
    del a, b, x2
    r.free(); t.free()


# ---------------------------------------------------------------- the futhark_* link-compat aliases
def test_futhark_compat_layer_runs_the_reference_entries():
    """Drives libhark.so ONLY through the names `futhark c --library main.fut` would generate (setup.sh:12):
    data.csv GROUP BY of test.py:7 and the README projection, against the golden vectors."""
    import ctypes as C
    need_gpu()
    from harkdb_b200 import hark_ffi
    lib = C.CDLL(hark_ffi.LIB_PATH)
    P = C.c_void_p
    for name, res, args in [
        ("futhark_context_config_new", P, []), ("futhark_context_new", P, [P]), ("futhark_context_free", None, [P]),
        ("futhark_context_config_free", None, [P]), ("futhark_context_sync", C.c_int, [P]),
        ("futhark_context_get_error", P, [P]),
        ("futhark_new_u32_2d", P, [P, P, C.c_int64, C.c_int64]), ("futhark_new_i32_2d", P, [P, P, C.c_int64, C.c_int64]),
        ("futhark_new_i32_1d", P, [P, P, C.c_int64]),
        ("futhark_shape_u32_2d", C.POINTER(C.c_int64), [P, P]), ("futhark_shape_i32_2d", C.POINTER(C.c_int64), [P, P]),
        ("futhark_values_u32_2d", C.c_int, [P, P, P]), ("futhark_values_i32_2d", C.c_int, [P, P, P]),
        ("futhark_free_u32_2d", C.c_int, [P, P]), ("futhark_free_i32_2d", C.c_int, [P, P]),
        ("futhark_free_i32_1d", C.c_int, [P, P]),
        ("futhark_entry_query_sel", C.c_int, [P, C.POINTER(P), P, P]),
        ("futhark_entry_query_groupby", C.c_int, [P, C.POINTER(P), P, C.c_int32, P, P]),
    ]:
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    cfg = lib.futhark_context_config_new()
    ctx = lib.futhark_context_new(cfg)
    assert ctx
    db = np.ascontiguousarray(np.asarray(DATA["rows"], dtype=np.uint32))
    h_db = lib.futhark_new_u32_2d(ctx, db.ctypes.data, db.shape[0], db.shape[1])
    s = np.array([0, 2], dtype=np.int32)
    t = np.array([0, 3], dtype=np.int32)
    h_s, h_t = lib.futhark_new_i32_1d(ctx, s.ctypes.data, 2), lib.futhark_new_i32_1d(ctx, t.ctypes.data, 2)
    out = P()
    assert lib.futhark_entry_query_groupby(ctx, C.byref(out), h_db, 0, h_s, h_t) == 0
    assert lib.futhark_context_sync(ctx) == 0
    shp = lib.futhark_shape_u32_2d(ctx, out)
    res = np.empty((shp[0], shp[1]), dtype=np.uint32)
    assert lib.futhark_values_u32_2d(ctx, out, res.ctypes.data) == 0
    assert res.tolist() == [[0, 0, 0], [1, 1, 3], [6, 6, 6]]                        # SURVEY App. B / test.py:7
    lib.futhark_free_u32_2d(ctx, out)
    dbi = db.astype(np.int32)
    h_dbi = lib.futhark_new_i32_2d(ctx, dbi.ctypes.data, dbi.shape[0], dbi.shape[1])
    out = P()
    assert lib.futhark_entry_query_sel(ctx, C.byref(out), h_dbi, h_s) == 0
    shp = lib.futhark_shape_i32_2d(ctx, out)
    res = np.empty((shp[0], shp[1]), dtype=np.int32)
    assert lib.futhark_values_i32_2d(ctx, out, res.ctypes.data) == 0
    assert res.tolist() == [[6, 6], [0, 0], [0, 0], [0, 0], [0, 0], [6, 6], [1, 3]]   # README.md:42
    bad = np.array([0, 99], dtype=np.int32)
    h_bad = lib.futhark_new_i32_1d(ctx, bad.ctypes.data, 2)
    out2 = P()
    assert lib.futhark_entry_query_sel(ctx, C.byref(out2), h_dbi, h_bad) != 0       # Futhark: index out of bounds
    err = lib.futhark_context_get_error(ctx)
    assert err and b"bounds" in C.string_at(err)
    C.CDLL(None).free(P(err))
    for h in (h_s, h_t, h_bad):
        lib.futhark_free_i32_1d(ctx, h)
    lib.futhark_free_i32_2d(ctx, out); lib.futhark_free_i32_2d(ctx, h_dbi); lib.futhark_free_u32_2d(ctx, h_db)
    lib.futhark_context_free(ctx)
    lib.futhark_context_config_free(cfg)
