"""Host-side logic that needs no GPU: SQL front-end dict shapes, plan construction (parse.py parity with
the reference's keys and messages), table loaders, and that libhark.so loads and exports every symbol
include/hark.h declares."""

import ctypes
import os
import re

import numpy as np
import pandas as pd
import pytest

from harkdb_b200 import hark_ffi, sqlmini
from harkdb_b200.parse import finalize_pred, getIndex, sql_parse
from harkdb_b200.table import Table, entry_dtype, load_np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def _game_table():
    import json
    d = json.load(open(os.path.join(GOLDEN, "data_csv.json")))
    return Table("game_1", pd.DataFrame(d["rows"], columns=d["columns"]))


# ---- moz_sql_parser dict shapes parse.py relies on ----
def test_sql_shapes():
    p = sqlmini.parse
    assert p("select a, b from t") == {"select": [{"value": "a"}, {"value": "b"}], "from": "t"}
    assert p("select a from t") == {"select": {"value": "a"}, "from": "t"}
    assert p("SELECT * FROM t") == {"select": "*", "from": "t"}
    assert p("select max(c) from t")["select"] == {"value": {"max": "c"}}
    assert p("select a from t group by a")["groupby"] == {"value": "a"}
    assert p("select a from t group by a, b")["groupby"] == [{"value": "a"}, {"value": "b"}]
    assert p("select a from t where col2 > 0.5 and col5 < 2")["where"] == \
        {"and": [{"gt": ["col2", 0.5]}, {"lt": ["col5", 2]}]}
    assert p("select a from t where a >= -3")["where"] == {"gte": ["a", -3]}
    assert p("select a from t where a <> 3 or not b = 1")["where"] == \
        {"or": [{"neq": ["a", 3]}, {"not": {"eq": ["b", 1]}}]}
    assert p("select a, count(c) from t group by a having count(c) > 3")["having"] == {"gt": [{"count": "c"}, 3]}
    assert p("select a from t order by a, b desc")["orderby"] == [{"value": "a"}, {"value": "b", "sort": "desc"}]
    assert p("select a from t order by a")["orderby"] == {"value": "a"}
    assert p("select f.x from f inner join d on f.fk = d.pk")["from"] == \
        ["f", {"inner join": "d", "on": {"eq": ["f.fk", "d.pk"]}}]
    assert p("select a as x from t limit 10;") == {"select": {"value": "a", "name": "x"}, "from": "t", "limit": 10}
    with pytest.raises(sqlmini.SqlSyntaxError):
        p("select from t")
    with pytest.raises(sqlmini.SqlSyntaxError):
        p("select a from t where")


# ---- plans: exactly the reference's keys/values for the two forms it implements ----
def test_plan_select_matches_reference_shape():
    t = _game_table()
    plan = sql_parse({"game_1": t}, "select col1, col3 from game_1")      # README.md:42
    assert set(plan) == {"table", "select"} and plan["select"] == [0, 2]
    assert plan["table"] is t.get_data()            # not uploaded -> the host array, as parse.py:58
    # statements that crash the reference but are plain projections
    assert sql_parse({"game_1": t}, "select col8 from game_1")["select"] == [7]
    assert sql_parse({"game_1": t}, "select * from game_1")["select"] == list(range(8))


def test_plan_groupby_matches_reference_shape():
    t = _game_table()
    plan = sql_parse({"game_1": t}, "select col1,  max(col3) from game_1 group by col1")   # test.py:7
    assert set(plan) == {"select", "groupbys", "table", "g_col"}
    assert plan["select"] == [0, 2] and plan["groupbys"] == [0, 3] and plan["g_col"] == 0
    plan = sql_parse({"game_1": t}, "select prod(col2), sum(col3), max(col8), min(col5) from game_1 group by col2")
    assert plan["select"] == [1, 2, 7, 4] and plan["groupbys"] == [1, 2, 3, 4] and plan["g_col"] == 1


def test_plan_errors_are_the_references():
    t = _game_table()
    with pytest.raises(Exception, match="nope is not in tables"):                       # parse.py:33
        sql_parse({"game_1": t}, "select col1 from nope")
    with pytest.raises(Exception, match="colX is not in the schema of table game_1"):   # parse.py:54
        sql_parse({"game_1": t}, "select col1, colX from game_1")
    with pytest.raises(Exception, match="colX is not in the schema of table game_1"):   # parse.py:69
        sql_parse({"game_1": t}, "select col1 from game_1 group by colX")
    with pytest.raises(Exception, match="col2 is not an aggregation function"):         # parse.py:78
        sql_parse({"game_1": t}, "select col1, col2 from game_1 group by col1")
    with pytest.raises(Exception, match="colY is not in the schema of table game_1"):   # parse.py:87
        sql_parse({"game_1": t}, "select col1, max(colY) from game_1 group by col1")


def test_plan_extensions():
    t = _game_table()
    plan = sql_parse({"game_1": t}, "select col1, col3 from game_1 where col2 > 0.5 and 3 >= col5 order by col3 desc limit 4")
    assert plan["select"] == [0, 2]
    assert plan["where"] == [(1, 0, None, 0.5), (4, 3, 3, 3.0)]      # constant-on-the-left is flipped
    assert plan["orderby"] == [(2, 1)] and plan["limit"] == 4
    plan = sql_parse({"game_1": t}, "select col1, sum(col2), count(col2), avg(col2) from game_1 group by col1 "
                                    "having count(col2) > 1 order by col1 desc")
    assert plan["groupbys"] == [0, 2, 5, 6] and plan["select"] == [0, 1, 1, 1]
    assert plan["having"] == [(3, 0, 1, 1.0)]          # output column 3 = count (0 is the key)
    assert plan["orderby"] == [(0, 1)]
    plan = sql_parse({"game_1": t}, "select col1 from game_1 where col1 > 1 or col2 < 3")
    assert plan["where"] == [(0, 0 | 0x100, 1, 1.0), (1, 2, 3, 3.0)]     # one OR-clause (HARK_PRED_OR on the first)
    with pytest.raises(Exception, match="right-hand side must be a numeric constant"):
        sql_parse({"game_1": t}, "select col1 from game_1 where col1 > col2")


def test_plan_join():
    f = Table("fact", pd.DataFrame({"fk": [1, 2], "val": [5, 6]}))
    d = Table("dim", pd.DataFrame({"pk": [1, 2], "attr": [7, 8]}))
    tabs = {"fact": f, "dim": d}
    plan = sql_parse(tabs, "select f.val, d.attr from fact f join dim d on f.fk = d.pk")
    assert plan["join"] == (0, 0) and plan["select"] == [1] and plan["select2"] == [1]
    plan = sql_parse(tabs, "select attr, sum(val), count(*) from fact join dim on pk = fk group by attr")
    assert plan["join"] == (0, 0) and plan["g_col"] == 1 and plan["select"] == [1, 0] and plan["groupbys"] == [2, 5]


def test_plan_select_distinct():
    assert sqlmini.parse("select distinct a, b from t") == {"select_distinct": [{"value": "a"}, {"value": "b"}], "from": "t"}
    t = _game_table()
    plan = sql_parse({"game_1": t}, "select distinct col3, col1 from game_1 where col2 < 6 order by col1 desc limit 2")
    assert plan["g_cols"] == [2, 0] and plan["distinct"] is True and plan["groupbys"] == [5]
    assert plan["orderby"] == [(1, 1)] and plan["limit"] == 2 and plan["where"] == [(1, 2, 6, 6.0)]
    with pytest.raises(Exception, match="must appear in the select list"):
        sql_parse({"game_1": t}, "select distinct col1 from game_1 order by col2")


def test_plan_global_aggregates():
    t = _game_table()
    plan = sql_parse({"game_1": t}, "select sum(col2), count(*), avg(col3) from game_1 where col1 > 0 limit 1")
    assert plan["global"] is True and plan["select"] == [1, 0, 2] and plan["groupbys"] == [2, 5, 6]
    assert plan["where"] == [(0, 0, 0, 0.0)] and plan["limit"] == 1 and "g_col" not in plan
    with pytest.raises(Exception, match="needs a GROUP BY clause"):
        sql_parse({"game_1": t}, "select col1, sum(col2) from game_1")
    with pytest.raises(Exception, match="need a GROUP BY"):
        sql_parse({"game_1": t}, "select sum(col2) from game_1 having sum(col2) > 1")


def test_plan_join_where_pushdown_having_orderby():
    f = Table("fact", pd.DataFrame({"fk": [1, 2], "val": [5, 6], "qty": [1, 2]}))
    d = Table("dim", pd.DataFrame({"pk": [1, 2], "attr": [7, 8]}))
    tabs = {"fact": f, "dim": d}
    plan = sql_parse(tabs, "select d.attr, sum(f.val) as s, count(*) from fact f join dim d on f.fk = d.pk "
                           "where f.qty > 1 and (d.attr = 7 or d.attr = 8) and not f.val < 0 "
                           "group by d.attr having s > 3 order by count(*) desc, d.attr limit 3")
    assert plan["where"] == [(2, 0, 1, 1.0), (1, 2 | 0x200, 0, 0.0)]                 # fact-side clauses, fact indices
    assert plan["where2"] == [(1, 4 | 0x100, 7, 7.0), (1, 4, 8, 8.0)]               # one OR-clause on the dim side
    assert plan["having"] == [(1, 0, 3, 3.0)] and plan["orderby"] == [(2, 1), (0, 0)] and plan["limit"] == 3
    assert plan["select"] == [1, 0] and plan["groupbys"] == [2, 5] and plan["g_col"] == 1
    plan = sql_parse(tabs, "select f.val, d.attr from fact f join dim d on f.fk = d.pk order by d.attr desc, f.val")
    assert plan["orderby"] == [(1, 1), (0, 0)] and "where" not in plan
    with pytest.raises(Exception, match="both joined tables"):
        sql_parse(tabs, "select f.val, d.attr from fact f join dim d on f.fk = d.pk where f.qty = 1 or d.attr = 7")
    with pytest.raises(Exception, match="HAVING needs a GROUP BY"):
        sql_parse(tabs, "select f.val, d.attr from fact f join dim d on f.fk = d.pk having f.val > 1")


def test_finalize_pred_integer_columns():
    assert finalize_pred((0, 0, None, 2.5), True) == (0, 0, 2, 2.5)     # x > 2.5  <=> x > 2
    assert finalize_pred((0, 1, None, 2.5), True) == (0, 0, 2, 2.5)     # x >= 2.5 <=> x > 2
    assert finalize_pred((0, 2, None, 2.5), True) == (0, 2, 3, 2.5)     # x < 2.5  <=> x < 3
    assert finalize_pred((0, 3, None, -2.5), True) == (0, 2, -2, -2.5)  # x <= -2.5 <=> x < -2
    assert finalize_pred((0, 0, 7, 7.0), True) == (0, 0, 7, 7.0)
    assert finalize_pred((0, 0, None, 0.5), False) == (0, 0, 0, 0.5)


# ---- WHERE / HAVING boolean expressions -> conjunctive normal form (hark.h HARK_PRED_OR / HARK_PRED_NOT) ----
def test_sql_shapes_between_in():
    p = sqlmini.parse
    assert p("select a from t where a between 1 and 5 and b > 2")["where"] == \
        {"and": [{"between": ["a", 1, 5]}, {"gt": ["b", 2]}]}
    assert p("select a from t where a not between -1 and 5")["where"] == {"not_between": ["a", -1, 5]}
    assert p("select a from t where a in (1, 2, 3)")["where"] == {"in": ["a", [1, 2, 3]]}
    assert p("select a from t where a not in (7)")["where"] == {"nin": ["a", 7]}
    assert p("select a from t where (a > 1 or b < 2) and not (c = 3 or a = 4)")["where"] == \
        {"and": [{"or": [{"gt": ["a", 1]}, {"lt": ["b", 2]}]}, {"not": {"or": [{"eq": ["c", 3]}, {"eq": ["a", 4]}]}}]}


def _eval_tree(expr, cols, names):
    """Direct evaluation of a moz_sql_parser-shaped boolean expression with numpy (IEEE comparisons)."""
    (op, args), = expr.items()
    if op == "and":
        return np.logical_and.reduce([_eval_tree(e, cols, names) for e in args])
    if op == "or":
        return np.logical_or.reduce([_eval_tree(e, cols, names) for e in args])
    if op == "not":
        return ~_eval_tree(args, cols, names)
    if op in ("between", "not_between"):
        x = cols[names.index(args[0])]
        m = (x >= args[1]) & (x <= args[2])
        return ~m if op == "not_between" else m
    if op in ("in", "nin"):
        x = cols[names.index(args[0])]
        vals = args[1] if isinstance(args[1], list) else [args[1]]
        m = np.logical_or.reduce([x == v for v in vals])
        return ~m if op == "nin" else m
    x = cols[names.index(args[0])]
    with np.errstate(invalid="ignore"):
        return {"gt": x > args[1], "gte": x >= args[1], "lt": x < args[1], "lte": x <= args[1], "eq": x == args[1],
                "neq": x != args[1]}[op]


def test_where_boolean_expressions_become_cnf_with_the_same_truth_table():
    from harkdb_b200.parse import PRED_NOT, PRED_OR, _preds
    from oracle import np_oracle as NO
    rng = np.random.default_rng(5)
    n = 4000
    names = ["a", "b", "c", "d"]
    cols = [rng.integers(-5, 6, n).astype(np.int32), rng.integers(0, 10, n).astype(np.int64),
            rng.integers(-5, 6, n).astype(np.float32), rng.random(n)]
    cols[2][rng.integers(0, n, 200)] = np.nan
    cols[3][rng.integers(0, n, 200)] = np.nan
    is_int = [True, True, False, False]
    resolve = names.index
    stmts = [
        "a > 1 or b < 3",
        "not (a > 1 or b < 3)",
        "not c > 0",                                      # true for NaN rows, unlike c <= 0
        "(a > 1 and b < 3) or (c >= 0 and d < 0.5)",
        "a between -2 and 2 and not d between 0.25 and 0.75",
        "a in (1, 3, 5) or c not in (0, 2)",
        "not (a = 1 and (b <> 2 or not (c < 1.5))) and d >= 0.1",
        "a > 2.5 or not b <= 3.5 or a = 1.5 or not a <> 0.5",
        "(a > 0 or b > 5 or c > 1) and (a < 3 or d < 0.9) and b <> 4",
    ]
    for w in stmts:
        tree = sqlmini.parse("select a from t where " + w)["where"]
        preds = [finalize_pred(p, is_int[p[0]]) for p in _preds(tree, resolve)]
        assert not (preds[-1][1] & PRED_OR) and len(preds) <= 16
        got = NO.pred_mask(cols, preds)
        exp = _eval_tree(tree, cols, names)
        assert np.array_equal(got, exp), w
    # a plain conjunction is exactly the flag-free list the reference-shaped plan always had
    tree = sqlmini.parse("select a from t where a > 1 and b <= 2 and c <> 0.5")["where"]
    assert _preds(tree, resolve) == [(0, 0, 1, 1.0), (1, 3, 2, 2.0), (2, 5, None, 0.5)]
    assert _preds(sqlmini.parse("select a from t where not a > 1")["where"], resolve) == [(0, 0 | PRED_NOT, 1, 1.0)]
    with pytest.raises(Exception, match="conjunctive normal form"):
        big = " or ".join(f"(a > {i} and b < {i} and c > {i})" for i in range(4))
        _preds(sqlmini.parse("select a from t where " + big)["where"], resolve)


def test_cnf_conversion_property_random_trees():
    """hypothesis: any AND / OR / NOT / BETWEEN / IN tree over four columns either converts to a CNF predicate list
    with the tree's truth table (NaN rows included) or is rejected for its size — never silently wrong."""
    from hypothesis import given, settings, strategies as st
    from harkdb_b200.parse import _preds
    from oracle import np_oracle as NO
    rng = np.random.default_rng(17)
    n = 600
    names = ["a", "b", "c", "d"]
    cols = [rng.integers(-3, 4, n).astype(np.int32), rng.integers(0, 5, n).astype(np.int64),
            np.round(rng.random(n) * 4).astype(np.float32) / 4, np.round(rng.random(n) * 4) / 4]
    cols[2][rng.integers(0, n, 60)] = np.nan
    cols[3][rng.integers(0, n, 60)] = np.nan
    is_int = [True, True, False, False]
    consts = {"a": [-2, 0, 1, 2.5], "b": [0, 2, 3, 1.5], "c": [0.25, 0.5, 1], "d": [0, 0.5, 0.75]}
    leaf = st.one_of(
        st.builds(lambda nm, op, i: f"{nm} {op} {consts[nm][i % len(consts[nm])]}", st.sampled_from(names),
                  st.sampled_from(["<", "<=", ">", ">=", "=", "<>"]), st.integers(0, 3)),
        st.builds(lambda nm, i: f"{nm} between {consts[nm][0]} and {consts[nm][i % len(consts[nm])]}", st.sampled_from(names),
                  st.integers(0, 3)),
        st.builds(lambda nm: f"{nm} in ({consts[nm][0]}, {consts[nm][1]})", st.sampled_from(names)),
        st.builds(lambda nm: f"{nm} not in ({consts[nm][1]}, {consts[nm][2]})", st.sampled_from(names)))
    tree = st.recursive(leaf, lambda ch: st.one_of(
        st.builds(lambda x, y: f"({x} and {y})", ch, ch), st.builds(lambda x, y: f"({x} or {y})", ch, ch),
        st.builds(lambda x: f"not ({x})", ch)), max_leaves=5)

    @settings(max_examples=150, deadline=None)
    @given(tree)
    def check(w):
        t = sqlmini.parse("select a from t where " + w)["where"]
        try:
            preds = [finalize_pred(p, is_int[p[0]]) for p in _preds(t, names.index)]
        except Exception as e:
            assert "conjunctive normal form" in str(e), (w, e)
            return
        assert np.array_equal(NO.pred_mask(cols, preds), _eval_tree(t, cols, names)), w

    check()


def test_c_oracle_filter_cnf_matches_numpy_oracle():
    from harkdb_b200.parse import PRED_NOT, PRED_OR
    from oracle import c_oracle as CO
    from oracle import np_oracle as NO
    rng = np.random.default_rng(6)
    for dtype in (NO.I32, NO.F32, NO.F64, NO.I64):
        a = rng.integers(-4, 5, (3000, 3)).astype(NO.NP_DTYPES[dtype])
        if dtype in (NO.F32, NO.F64):
            a[rng.integers(0, 3000, 100), rng.integers(0, 3, 100)] = np.nan
        preds = [(0, NO.GT | PRED_OR, 1, 1.0), (1, NO.LT | PRED_NOT | PRED_OR, 0, 0.0), (2, NO.EQ, 2, 2.0),
                 (1, NO.NE | PRED_NOT, -1, -1.0)]
        got = CO.query_filter(a, [2, 0], preds)
        exp = NO.query_filter([np.ascontiguousarray(a[:, c]) for c in range(3)], [2, 0], preds)
        assert np.array_equal(got[:, 0], exp[0], equal_nan=True) and np.array_equal(got[:, 1], exp[1], equal_nan=True)


# ---- GROUP BY over several columns (parse.py:64's TODO), per-column dtypes, Arrow ----
def test_plan_groupby_multi():
    t = Table("li", pd.DataFrame({"flag": [1, 2], "status": [0, 1], "qty": [5, 6], "price": [1.5, 2.5]}))
    plan = sql_parse({"li": t}, "select flag, status, sum(qty), avg(price) as ap, count(*) from li where qty > 1 "
                                "group by flag, status having count(*) > 0 and ap < 10.5 order by status desc, sum(qty)")
    assert plan["g_cols"] == [0, 1] and "g_col" not in plan
    assert plan["select"] == [0, 1, 2, 3, 0] and plan["groupbys"] == [0, 0, 2, 6, 5]
    assert plan["where"] == [(2, 0, 1, 1.0)]
    assert plan["having"] == [(6, 0, 0, 0.0), (5, 2, None, 10.5)]      # outputs: 2 keys, then the 5 select items
    assert plan["orderby"] == [(1, 1), (4, 0)]
    with pytest.raises(Exception, match="not an aggregation function"):
        sql_parse({"li": t}, "select qty from li group by flag, status")
    with pytest.raises(Exception, match="twice"):
        sql_parse({"li": t}, "select flag from li group by flag, flag")


def test_table_keeps_one_dtype_per_column(tmp_path):
    import pyarrow as pa
    import pyarrow.parquet as pq
    df = pd.DataFrame({"k": [1, 2, 3], "big": [1, 2 ** 40, 3], "v": [0.5, 1.5, 2.5], "f": np.float32([1, 2, 3])})
    t = Table("x", df)
    assert t.get_data().dtype == np.float64 and t.get_data().shape == (3, 4)      # what the reference would hold
    assert t.get_column_dtypes() == [np.int32, np.int64, np.float64, np.float32]
    assert Table("h", pd.DataFrame({"a": [1, 2], "b": [3, 4]})).get_column_dtypes() == [np.int32, np.int32]
    from harkdb_b200.table import HostColumns
    h = t.get_handle()                              # not resident: the per-column arrays, narrowed
    assert isinstance(h, HostColumns) and [c.dtype for c in h] == [np.int32, np.int64, np.float64, np.float32]
    assert isinstance(Table("h", pd.DataFrame({"a": [1, 2], "b": [3, 4]})).get_handle(), np.ndarray)
    at = pa.table({"a": [1, 2, 3], "b": [1.0, 2.0, 3.0]})
    ta = Table("a", at)
    assert ta.get_schema() == ["a", "b"] and ta.get_column_dtypes() == [np.int32, np.float64]
    pq.write_table(at, tmp_path / "t.parquet")
    tp = Table("p", str(tmp_path / "t.parquet"))
    assert tp.get_schema() == ["a", "b"] and np.array_equal(tp.get_data(), ta.get_data())
    csv = tmp_path / "m.csv"
    csv.write_text("k,v\n1,0.5\n2,1.5\n")
    assert Table("c", str(csv)).get_column_dtypes() == [np.int32, np.float64]
    with pytest.raises(Exception, match="nulls"):
        Table("n", pa.table({"a": [1, None]}))


def test_np_oracle_groupby_multi_against_pandas():
    from oracle import np_oracle as NO
    rng = np.random.default_rng(21)
    n = 5000
    a = rng.integers(-3, 4, n).astype(np.int32)
    b = rng.integers(0, 5, n).astype(np.int64)
    c = rng.integers(0, 3, n).astype(np.uint32)
    v = rng.integers(-100, 100, n).astype(np.int32)
    f = rng.random(n)
    out = NO.query_groupby_multi([a, b, c, v, f], [0, 1, 2], [3, 3, 3, 4, 4, 3],
                                 [NO.AGG_SUM, NO.AGG_MIN, NO.AGG_MAX, NO.AGG_AVG, NO.AGG_COUNT, NO.AGG_KEY])
    g = pd.DataFrame({"a": a, "b": b, "c": c, "v": v, "f": f}).groupby(["a", "b", "c"], sort=True)
    exp = g.agg(s=("v", "sum"), mn=("v", "min"), mx=("v", "max"), av=("f", "mean"), n=("f", "size")).reset_index()
    assert [o.dtype for o in out[:3]] == [np.int32, np.int64, np.uint32]
    for j, nm in enumerate(["a", "b", "c", "s", "mn", "mx"]):
        assert np.array_equal(out[j].astype(np.int64), exp[nm].to_numpy().astype(np.int64)), nm
    assert np.allclose(out[6], exp["av"].to_numpy(), rtol=1e-12) and np.array_equal(out[7], exp["n"].to_numpy())
    assert np.array_equal(out[8], out[4])             # code 0 on a value column falls through to MIN (groupby.fut:41)
    hv = NO.query_groupby_multi([a, b, c, v, f], [1, 0], [3], [NO.AGG_COUNT], having=[(2, NO.GT, 150, 150.0)])
    cnt = pd.DataFrame({"a": a, "b": b}).groupby(["b", "a"], sort=True).size().reset_index(name="n")
    cnt = cnt[cnt["n"] > 150]
    assert np.array_equal(hv[0], cnt["b"].to_numpy()) and np.array_equal(hv[1], cnt["a"].to_numpy())
    assert np.array_equal(hv[2], cnt["n"].to_numpy())


# ---- K3t: the sort's pass-truncation decision is pure host code inside libhark.so (no device needed) ----
def _py_plan(norm_keys, n, slack=4):
    """Python restatement of plan_from_sample (csrc/sort.cu): (kstar, q, shift, passes) or None."""
    nk = len(norm_keys)
    bits = [int(int(k.max()) - 0).bit_length() for k in norm_keys]
    full, total = sum((b + 7) // 8 for b in bits), sum(bits)
    T = int(n - 1).bit_length() + slack
    if T + 8 > total:
        return None
    S = min(n, 32768)
    rows = (np.arange(S, dtype=object) * n // S).astype(np.int64)
    samp = [k[rows] for k in norm_keys]
    order = np.lexsort(samp[::-1])
    diff = []
    for s in range(1, S):
        a, b = order[s - 1], order[s]
        for k in range(nk):
            x, y = int(samp[k][a]), int(samp[k][b])
            if x != y:
                diff.append((k, (x ^ y).bit_length()))
                break
    thr = int(max(4.0, S * S / (8.0 * n)))
    while True:
        acc, kstar, q = 0, nk - 1, 0
        for k in range(nk):
            if acc >= T:
                kstar, q = k - 1, 0
                break
            qd = (T - acc + 7) // 8
            if 8 * qd >= bits[k]:
                acc += bits[k]
                continue
            kstar, q = k, qd
            break
        np_, shift = 0, 0
        for k in range(kstar, -1, -1):
            sh0 = bits[k] - 8 * q if (k == kstar and q > 0 and 8 * q < bits[k]) else 0
            if k == kstar:
                shift = sh0
            np_ += len(range(sh0, bits[k], 8))
        if np_ >= full:
            return None
        ties = sum(1 for d in diff if d[0] > kstar or (d[0] == kstar and d[1] <= shift))
        if ties <= thr:
            return (kstar, q, shift, np_)
        T += 8


def _lib_plan(lib, norm_keys, n, slack=4):
    import ctypes as C
    nk = len(norm_keys)
    bits = (C.c_int32 * nk)(*[int(int(k.max())).bit_length() for k in norm_keys])
    S = min(n, 32768)
    rows = (np.arange(S, dtype=object) * n // S).astype(np.int64)
    sample = np.ascontiguousarray(np.stack([k[rows] for k in norm_keys], axis=1).astype(np.uint64))
    out = [C.c_int32(0) for _ in range(4)]
    on = lib.hark_debug_plan_truncation(n, nk, bits, sample.ctypes.data_as(C.POINTER(C.c_uint64)), S, slack,
                                        *[C.byref(o) for o in out])
    return tuple(o.value for o in out) if on else None


def test_sort_truncation_planner_matches_its_python_restatement():
    lib = hark_ffi.load_library()
    rng = np.random.default_rng(23)

    def norm(a):                      # ordkey - min for non-negative integer test data
        a = a.astype(np.uint64)
        return a - a.min()

    n = 200003
    cases = {
        "config-4 shape": [norm(rng.integers(0, 1 << 20, n)), norm(rng.integers(0, 1 << 63, n, dtype=np.int64))],
        "one wide key": [norm(rng.integers(0, 1 << 63, n, dtype=np.int64))],
        "key boundary": [norm(rng.integers(0, 1 << 24, n)), norm(rng.integers(0, 1 << 62, n, dtype=np.int64))],
        "duplicates": [norm(rng.integers(0, 1 << 62, 1000, dtype=np.int64)[rng.integers(0, 1000, n)])],
        "two clusters": [norm(rng.integers(0, 1 << 20, n).astype(np.int64) + (np.arange(n) % 2) * (1 << 60))],
        "narrow": [norm(rng.integers(0, 1 << 16, n)), norm(rng.integers(0, 1 << 8, n))],
        "three keys": [norm(rng.integers(0, 8, n)), norm(rng.integers(0, 1 << 40, n, dtype=np.int64)),
                       norm(rng.integers(0, 1 << 50, n, dtype=np.int64))],
    }
    got = {}
    for name, keys in cases.items():
        for slack in (0, 4):
            exp = _py_plan(keys, n, slack)
            got[(name, slack)] = _lib_plan(lib, keys, n, slack)
            assert got[(name, slack)] == exp, (name, slack, got[(name, slack)], exp)
    assert got[("config-4 shape", 4)] == (1, 1, 55, 4)        # 20 bits of key 0 (3 passes) + the top digit of the 63-bit key 1
    assert got[("key boundary", 4)] == (0, 0, 0, 3)           # key 0 covers log2(n) + slack: key 1 is never sorted
    assert got[("duplicates", 4)] is not None and got[("two clusters", 4)] is None and got[("narrow", 4)] is None


def test_integration_c_example_compiles_and_links(tmp_path):
    """INTEGRATION.md §C is a real C11 program against include/hark.h: compile it and link it against libhark.so."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    code = re.search(r"```c\n(.*?)```", text, re.S).group(1)
    src = tmp_path / "client.c"
    src.write_text(code)
    libdir = os.path.join(ROOT, "harkdb_b200")
    lib = tmp_path / "libhark.so"            # -lhark wants that name on the link line
    os.symlink(os.path.join(libdir, "libhark.so"), lib)
    r = subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src),
                        "-L", str(tmp_path), "-lhark", "-Wl,-rpath," + libdir, "-Wl,--allow-shlib-undefined",
                        "-o", str(tmp_path / "client")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


# ---- table.py ----
def test_table_loaders(tmp_path):
    assert getIndex(["a", "b"], "b") == 1 and getIndex(["a"], "z") == -1
    arr = np.arange(6).reshape(2, 3)
    assert load_np(arr)[1] == ["col1", "col2", "col3"]      # by column count (reference bug table.py:14 fixed)
    t = Table("x", arr)
    assert t.get_name() == "x" and t.get_schema() == ["col1", "col2", "col3"] and t.get_data() is arr
    csv = tmp_path / "d.csv"
    csv.write_text("col1,col2\n6, 6\n1, 2\n")
    tc = Table("c", str(csv))
    assert tc.get_schema() == ["col1", "col2"] and tc.get_data().dtype == np.int64
    txt = tmp_path / "d.txt"
    txt.write_text("1 2 3\n4 5 6\n")
    tt = Table("t", str(txt))
    assert tt.get_schema() == ["c1", "c2", "c3"] and tt.get_data().dtype == np.float64
    with pytest.raises(Exception, match="We do not support loading this file type"):
        Table("b", "x.orc")
    with pytest.raises(Exception, match="Table is not in a file, numpy array or dataframe"):
        Table("b", 42)
    assert entry_dtype(np.array([[1, -5]])) == np.int32
    assert entry_dtype(np.array([[1, 2 ** 32 - 1]])) == np.uint32
    assert entry_dtype(np.array([[-1, 2 ** 32]])) == np.int64
    assert entry_dtype(np.array([[0.5]], dtype=np.float32)) == np.float32


def test_convert_for_entry_range_check():
    ok = hark_ffi.convert_for_entry(np.array([[6, 0]], dtype=np.int64), np.int32, "t")
    assert ok.dtype == np.int32
    with pytest.raises(hark_ffi.HarkError):
        hark_ffi.convert_for_entry(np.array([[2 ** 40]], dtype=np.int64), np.int32, "t")
    with pytest.raises(hark_ffi.HarkError):
        hark_ffi.convert_for_entry(np.array([[-1]], dtype=np.int64), np.uint32, "t")
    with pytest.raises(hark_ffi.HarkError):
        hark_ffi.convert_for_entry(np.array([[0.5]]), np.int32, "t")


# ---- the C-ABI library: loads, exports every declared symbol, fails loudly without a GPU ----
def _declared_symbols(header):
    text = open(header).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hark_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = hark_ffi.load_library()
    names = _declared_symbols(os.path.join(ROOT, "include", "hark.h"))
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/hark.h but not exported by libhark.so"
        assert n in hark_ffi.SIGNATURES, f"{n} has no ctypes signature"
    assert lib.hark_abi_version() == 1


def test_futhark_compat_symbols_exported():
    header = os.path.join(ROOT, "include", "hark_futhark_compat.h")
    text = re.sub(r"/\*.*?\*/", "", open(header).read(), flags=re.S)
    names = sorted(set(re.findall(r"\b(futhark_[a-z0-9_]+)\s*\(", text)))
    lib = ctypes.CDLL(hark_ffi.LIB_PATH)
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(hark_ffi.HarkError, match="no usable CUDA device|no CPU fallback"):
        hark_ffi.Futhark()
    from harkdb_b200 import FutharkContext
    with pytest.raises(hark_ffi.HarkError):
        FutharkContext()


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "harkdb_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                bad = re.search(r"(^|\n)\s*(import\s+oracle|from\s+\.*oracle|#include\s+\"[^\"]*oracle)|liboracle|c_oracle|np_oracle", src)
                assert not bad, f"{f} uses the oracle: {bad.group(0)!r}"
