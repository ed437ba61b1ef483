"""world_size-2 (and 3) `gloo` tests of the multi-GPU host logic (harkdb_b200/sharded.py) on CPU.

The local operators are played by the oracle (tests/oracle_engine.py); what is under test is everything the ranks do
TOGETHER: row-range sharding, splitter sampling, range repartition + all-to-all, partial-aggregate expansion and merge,
dimension all-gather, and that the concatenation of the ranks' results in rank order equals the single-process
oracle result bit for bit (integers) / within the stated tolerance (float SUM/AVG)."""

import os
import socket
import traceback

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import np_oracle as NO


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, scenario, q):
    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from harkdb_b200.sharded import ShardedEnv
        from tests.oracle_engine import OracleEngine
        senv = ShardedEnv(OracleEngine())
        globals()["_scn_" + scenario](senv)
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception:
        q.put((rank, traceback.format_exc()))


def _run(scenario, world=2):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, scenario, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    for rank, msg in res:
        assert msg == "ok", f"rank {rank}:\n{msg}"


def _check(got, exp, what=""):
    assert len(got) == len(exp), what
    for g, e in zip(got, exp):
        assert g.dtype == e.dtype and g.shape == e.shape, (what, g.dtype, e.dtype, g.shape, e.shape)
        if g.dtype.kind == "f":
            assert np.allclose(g, e, rtol=1e-5 if g.dtype == np.float32 else 1e-12, atol=0), what
        else:
            assert np.array_equal(g, e), what


def _table(seed, n, kinds):
    rng = np.random.default_rng(seed)          # every rank builds the same global table
    cols = []
    for k in kinds:
        if k == "key":
            cols.append(rng.integers(-50, 50, n).astype(np.int32))
        elif k == "skew":
            cols.append(np.minimum(rng.zipf(1.3, n), 1000).astype(np.int32))
        elif k == "i64":
            cols.append(rng.integers(-2 ** 40, 2 ** 40, n).astype(np.int64))
        elif k == "f64":
            cols.append(rng.random(n).astype(np.float64))
        elif k == "f32":
            cols.append(rng.random(n).astype(np.float32))
        else:
            cols.append(rng.integers(-1000, 1000, n).astype(np.int32))
    return cols


def _shard(senv, cols):
    n = len(cols[0])
    lo, hi = senv.rank * n // senv.world, (senv.rank + 1) * n // senv.world
    from harkdb_b200.sharded import ShardTable
    return ShardTable(senv, senv.engine.from_columns([c[lo:hi] for c in cols]))


# ------------------------------------------------------------------ scenarios (run on every rank)
def _scn_filter(senv):
    cols = _table(1, 10007, ["key", "val", "f64"])
    t = _shard(senv, cols)
    preds = [(0, NO.GT, 0, 0.0), (2, NO.LT, 0, 0.5)]
    r = senv.query_filter(t, [2, 0], preds)
    _check(senv.gather_columns(r), NO.query_filter(cols, [2, 0], preds), "filter")


def _scn_groupby(senv):
    for kinds, n in ((["key", "val", "f64", "i64"], 20011), (["skew", "val", "f32", "i64"], 5003), (["key", "val", "f64", "i64"], 3)):
        cols = _table(2, n, kinds)
        t = _shard(senv, cols)
        ops = [NO.AGG_SUM, NO.AGG_COUNT, NO.AGG_AVG, NO.AGG_MAX, NO.AGG_MIN, NO.AGG_AVG, NO.AGG_SUM, NO.AGG_PROD, 99]
        s_cols = [1, 1, 1, 1, 3, 2, 2, 1, 1]
        r = senv.query_groupby_ex(t, 0, s_cols, ops)
        _check(senv.gather_columns(r), NO.query_groupby_ex(cols, 0, s_cols, ops), f"groupby {kinds} {n}")
        hv = [(2, NO.GT, 3, 0.0)]
        r = senv.query_groupby_ex(t, 0, s_cols, ops, having=hv)
        _check(senv.gather_columns(r), NO.query_groupby_ex(cols, 0, s_cols, ops, having=hv), "groupby having")


def _scn_groupby_dense(senv):
    """Shapes the all-reduce merge takes (sums, counts, AVG, integer MIN / MAX over a small dense key domain): negative
    keys, u32 keys above 2^31, a rank without rows, HAVING, SUM64; then the same query with the reduce switched off
    must give the same answer through the repartition."""
    rng = np.random.default_rng(12)
    for kdt, n in ((np.int32, 20011), (np.uint32, 5003), (np.int64, 1), (np.int32, 0)):
        lo = {np.int32: -700, np.uint32: 2 ** 31 - 300, np.int64: -(2 ** 40)}[kdt]
        cols = [(lo + rng.integers(0, 900, n)).astype(kdt), rng.integers(-1000, 1000, n).astype(np.int32),
                rng.random(n).astype(np.float32), rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)]
        t = _shard(senv, cols)
        ops = [NO.AGG_SUM, NO.AGG_COUNT, NO.AGG_AVG, NO.AGG_MAX, NO.AGG_MIN, NO.AGG_SUM, NO.AGG_AVG, NO.AGG_MAX, NO.AGG_MIN, NO.AGG_SUM64]
        s_cols = [1, 1, 1, 1, 1, 2, 2, 3, 3, 3]
        for hv in ((), [(2, NO.GT, 20, 0.0)]):
            exp = NO.query_groupby_ex(cols, 0, s_cols, ops, having=list(hv))
            for dense in (True, False):
                senv.dense_merge = dense
                senv.pop_trace()
                senv.trace_on = True
                r = senv.query_groupby_ex(t, 0, s_cols, ops, hv)
                tr = senv.pop_trace()
                senv.trace_on = False
                assert ("merge_reduce" in tr) == (dense and senv.world > 1), tr
                assert ("exchange" in tr) == (not dense and senv.world > 1), tr
                _check(senv.gather_columns(r), exp, f"groupby dense={dense} {kdt} {n} {hv}")
    senv.dense_merge = True


def _scn_groupby_multi(senv):
    rng = np.random.default_rng(8)
    n = 15013
    cols = [rng.integers(-3, 4, n).astype(np.int32), rng.integers(10 ** 12, 10 ** 12 + 9, n).astype(np.int64),
            rng.integers(0, 3, n).astype(np.uint32), rng.integers(-100, 100, n).astype(np.int32), rng.random(n)]
    t = _shard(senv, cols)
    ops = [NO.AGG_SUM, NO.AGG_AVG, NO.AGG_COUNT, NO.AGG_MIN, NO.AGG_MAX, NO.AGG_KEY]
    s_cols = [3, 4, 3, 3, 4, 0]
    for g_cols in ([0, 1, 2], [2, 0], [1]):
        r = senv.query_groupby_multi(t, g_cols, s_cols, ops)
        exp = NO.query_groupby_multi(cols, g_cols, s_cols, ops) if len(g_cols) > 1 else NO.query_groupby_ex(cols, g_cols[0], s_cols, ops)
        _check(senv.gather_columns(r), exp, f"groupby_multi {g_cols}")
    hv = [(3, NO.GT | NO.PRED_OR, 40, 40.0), (5, NO.GE, 120, 120.0)]           # SUM > 40 OR COUNT >= 120
    r = senv.query_groupby_multi(t, [0, 2], s_cols, ops, having=hv)
    _check(senv.gather_columns(r), NO.query_groupby_multi(cols, [0, 2], s_cols, ops, having=hv), "groupby_multi having")
    import pandas as pd
    from harkdb_b200.sharded import ShardedFutharkContext
    fc = ShardedFutharkContext(engine=senv.engine)
    df = pd.DataFrame({"a": cols[0], "c": cols[2].astype(np.int64), "v": cols[3], "f": cols[4]})
    fc.create_table("t", df)
    out = fc.sql("select a, c, sum(v), avg(f) from t where v > -50 group by a, c order by a desc, c")
    g = df[df.v > -50].groupby(["a", "c"], sort=True).agg(s=("v", "sum"), m=("f", "mean")).reset_index()
    g = g.sort_values(["a", "c"], ascending=[False, True], kind="stable")
    assert np.array_equal(out[:, 0], g.a) and np.array_equal(out[:, 1], g.c) and np.array_equal(out[:, 4], g.s)
    assert np.allclose(out[:, 5], g.m, rtol=1e-12)


def _scn_groupby_pinned(senv):
    rng = np.random.default_rng(3)
    db = rng.integers(0, 2 ** 32, (6007, 4), dtype=np.uint64).astype(np.uint32)
    db[:, 0] = rng.integers(2 ** 31 - 20, 2 ** 31 + 20, 6007)           # unsigned key order across the sign bit
    out = senv.from_futhark(senv.query_groupby(db, 0, [1, 2, 3, 1, 0], [1, 2, 3, 4, 0]))
    assert np.array_equal(out, NO.query_groupby(db, 0, [1, 2, 3, 1, 0], [1, 2, 3, 4, 0]))
    empty = np.zeros((0, 4), dtype=np.uint32)
    assert senv.from_futhark(senv.query_groupby(empty, 0, [1], [2])).shape == (0, 2)


def _scn_orderby(senv):
    cols = _table(4, 30011, ["key", "i64", "f64", "val"])
    cols[2][::97] = np.nan
    t = _shard(senv, cols)
    for sel, keys, desc in (([0, 1], [0, 1], [0, 0]), ([3, 2, 0], [0, 2], [1, 0]), ([1], [2, 0, 3], [0, 1, 1]), ([0, 3], [0], [0])):
        r = senv.query_orderby(t, sel, keys, desc)
        got = senv.gather_columns(r)
        exp = NO.query_orderby(cols, sel, keys, desc)
        for g, e in zip(got, exp):
            assert g.dtype == e.dtype and np.array_equal(g, e, equal_nan=True), (sel, keys, desc)   # stable => bit-exact
    # heavy duplicates: every row has the same key -> everything lands on one rank, order = global row order
    cols2 = [np.full(5000, 7, dtype=np.int32), np.arange(5000, dtype=np.int32)]
    r = senv.query_orderby(_shard(senv, cols2), [0, 1], [0], [0])
    _check(senv.gather_columns(r), cols2, "orderby all-equal keys")


def _scn_join(senv):
    rng = np.random.default_rng(5)
    nd, nf = 2003, 40009
    pk = rng.permutation(nd * 2)[:nd].astype(np.int32)
    dim = [pk, rng.integers(0, 37, nd).astype(np.int32)]
    fact = [rng.integers(0, nd * 2, nf).astype(np.int32), rng.integers(-100, 100, nf).astype(np.int32)]
    ops = [NO.AGG_SUM, NO.AGG_COUNT, NO.AGG_AVG, NO.AGG_MIN]
    r = senv.join_groupby(_shard(senv, fact), _shard(senv, dim), 0, 0, 1, [1, 1, 1, 1], ops)
    _check(senv.gather_columns(r), NO.join_groupby(fact, dim, 0, 0, 1, [1, 1, 1, 1], ops), "join_groupby")
    # reference join (u32, ordered by key, left row, right row) with duplicates on both sides
    db1 = rng.integers(0, 50, (3001, 3), dtype=np.int64).astype(np.uint32)
    db2 = rng.integers(0, 60, (1009, 2), dtype=np.int64).astype(np.uint32)
    out = senv.from_futhark(senv.join(db1, db2, 0, 0, [0, 2], [1]))
    assert np.array_equal(out, NO.join(db1, db2, 0, 0, [0, 2], [1]))


def _scn_sql(senv):
    import pandas as pd
    from harkdb_b200.sharded import ShardedFutharkContext
    rows = [[6] * 8, [0] * 8, [0] * 8, [0] * 8, [0] * 8, [6] * 8, [1, 2, 3, 4, 5, 3, 2, 1]]       # data.csv
    fc = ShardedFutharkContext(engine=senv.engine)
    fc.create_table("game_1", pd.DataFrame(rows, columns=[f"col{i + 1}" for i in range(8)]))
    assert fc.sql("select col1, col3 from game_1").tolist() == [[6, 6], [0, 0], [0, 0], [0, 0], [0, 0], [6, 6], [1, 3]]
    assert fc.sql("select col1,  max(col3) from game_1 group by col1").tolist() == [[0, 0, 0], [1, 1, 3], [6, 6, 6]]
    out = fc.sql("select col1, sum(col2), count(col2), avg(col2) from game_1 group by col1 having count(col2) > 1")
    assert out.tolist() == [[0.0, 0.0, 0.0, 4.0, 0.0], [6.0, 6.0, 12.0, 2.0, 6.0]]
    assert fc.sql("select col1, col3 from game_1 where col1 > 0 order by col3 desc, col1").tolist() == [[6, 6], [6, 6], [1, 3]]
    # aggregates without GROUP BY (TPC-H Q6's shape): one row, the constant key is dropped
    out = fc.sql("select sum(col2), count(*), avg(col2), max(col3) from game_1 where col1 > 0")
    assert out.shape == (1, 4) and out[0].tolist() == [14.0, 3.0, 14.0 / 3.0, 6.0]
    assert fc.sql("select count(*) from game_1").tolist() == [[7]]
    assert fc.sql("select distinct col1 from game_1").tolist() == [[0], [1], [6]]
    assert fc.sql("select distinct col3, col1 from game_1 where col2 < 6 order by col1 desc").tolist() == [[3, 1], [0, 0]]
    out = fc.sql("select min(col1), count(*) from game_1 where col1 > 100")       # zero rows in: one row out (NULL = NaN)
    assert out.shape == (1, 2) and np.isnan(out[0, 0]) and out[0, 1] == 0


def _scn_sql_join(senv):
    """JOIN with WHERE pushed down to both sides, GROUP BY, HAVING, ORDER BY, LIMIT — against pandas."""
    import pandas as pd
    from harkdb_b200.sharded import ShardedFutharkContext
    rng = np.random.default_rng(11)
    nd, nf = 503, 20011
    dim = pd.DataFrame({"pk": rng.permutation(nd).astype(np.int64), "attr": rng.integers(0, 9, nd), "region": rng.integers(0, 4, nd)})
    fact = pd.DataFrame({"fk": rng.integers(0, nd + 40, nf), "val": rng.integers(-50, 50, nf), "qty": rng.integers(1, 20, nf)})
    fc = ShardedFutharkContext(engine=senv.engine)
    fc.create_table("fact", fact)
    fc.create_table("dim", dim)
    out = fc.sql("select d.attr, sum(f.val), count(*) as n from fact f join dim d on f.fk = d.pk "
                 "where f.qty > 5 and (d.region = 1 or d.region = 3) and not f.val between -10 and 10 "
                 "group by d.attr having n > 50 order by sum(f.val) desc, d.attr limit 5")
    j = fact[(fact.qty > 5) & ~fact.val.between(-10, 10)].merge(dim[dim.region.isin([1, 3])], left_on="fk", right_on="pk")
    g = j.groupby("attr", sort=True).agg(s=("val", "sum"), n=("val", "size")).reset_index()
    g = g[g.n > 50].sort_values(["s", "attr"], ascending=[False, True], kind="stable").head(5)
    assert out.tolist() == g[["attr", "s", "n"]].to_numpy().tolist(), (out.tolist(), g.to_numpy().tolist())
    # plain join, reference row order (key, left row, right row), then ORDER BY over its output columns
    out = fc.sql("select f.fk, f.val, d.attr from fact f join dim d on f.fk = d.pk where f.qty = 19 and d.attr < 2 "
                 "order by d.attr desc, f.fk")          # (non-negative sort keys: the reference join's output is u32)
    j = fact[fact.qty == 19].reset_index().merge(dim[dim.attr < 2].reset_index(), left_on="fk", right_on="pk")
    j = j.sort_values(["fk", "index_x", "index_y"], kind="stable")              # join.fut:55-75 order
    j = j.sort_values(["attr", "fk"], ascending=[False, True], kind="stable")
    assert out.astype(np.int32).tolist() == j[["fk", "val", "attr"]].to_numpy().tolist()
    with pytest.raises(Exception, match="both joined tables"):
        fc.sql("select f.val, d.attr from fact f join dim d on f.fk = d.pk where f.qty = 19 or d.attr < 2")


# ------------------------------------------------------------------ tests
@pytest.mark.parametrize("scenario", ["filter", "groupby", "groupby_dense", "groupby_multi", "groupby_pinned", "orderby", "join", "sql", "sql_join"])
def test_sharded_world2(scenario):
    _run(scenario, 2)


def test_sharded_world1_sql_join():
    _run("sql_join", 1)


def test_sharded_world3_uneven_shards():
    _run("groupby", 3)
    _run("groupby_multi", 3)
    _run("orderby", 3)


# ------------------------------------------------------------------ pure host logic (no process group)
def test_expand_and_merge_ops():
    from harkdb_b200 import sharded as S
    p_s, p_ops = S.expand_partial_ops([4, 5, 6], [S.AGG_AVG, S.AGG_COUNT, 42])
    assert p_s == [4, 4, 5, 6] and p_ops == [S.AGG_SUMF64, S.AGG_COUNT, S.AGG_COUNT, S.AGG_MIN]
    assert S.merge_ops_for(p_ops) == [S.AGG_SUM, S.AGG_SUM, S.AGG_SUM, S.AGG_MIN]
    assert S.expand_partial_ops([1, 2], [0, 9], pinned_u32=True) == ([1, 2], [S.AGG_MIN, S.AGG_MIN])


def test_pick_splitters_weighted_quantiles():
    from harkdb_b200 import sharded as S
    a = np.arange(0, 100, dtype=np.uint64).reshape(-1, 1)           # rank 0: 100 samples, 1 row each
    b = np.arange(100, 110, dtype=np.uint64).reshape(-1, 1)         # rank 1: 10 samples standing for 10 rows each
    sp = S.pick_splitters([a, b], [1.0, 10.0], 2, 1)
    assert sp.shape == (1, 1) and sp[0, 0] == 99                    # half of the 200 rows lie at or below key 99
    sp = S.pick_splitters([np.zeros((0, 2), np.uint64)], [1.0], 4, 2)
    assert sp.shape == (3, 2)
    pos = S.sample_positions(10, 64)
    assert pos.tolist() == list(range(10)) and S.sample_positions(0, 8).size == 0
