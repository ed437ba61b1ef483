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
    with pytest.raises(Exception, match="only AND"):
        sql_parse({"game_1": t}, "select col1 from game_1 where col1 > 1 or col2 < 3")


def test_plan_join():
    f = Table("fact", pd.DataFrame({"fk": [1, 2], "val": [5, 6]}))
    d = Table("dim", pd.DataFrame({"pk": [1, 2], "attr": [7, 8]}))
    tabs = {"fact": f, "dim": d}
    plan = sql_parse(tabs, "select f.val, d.attr from fact f join dim d on f.fk = d.pk")
    assert plan["join"] == (0, 0) and plan["select"] == [1] and plan["select2"] == [1]
    plan = sql_parse(tabs, "select attr, sum(val), count(*) from fact join dim on pk = fk group by attr")
    assert plan["join"] == (0, 0) and plan["g_col"] == 1 and plan["select"] == [1, 0] and plan["groupbys"] == [2, 5]


def test_finalize_pred_integer_columns():
    assert finalize_pred((0, 0, None, 2.5), True) == (0, 0, 2, 2.5)     # x > 2.5  <=> x > 2
    assert finalize_pred((0, 1, None, 2.5), True) == (0, 0, 2, 2.5)     # x >= 2.5 <=> x > 2
    assert finalize_pred((0, 2, None, 2.5), True) == (0, 2, 3, 2.5)     # x < 2.5  <=> x < 3
    assert finalize_pred((0, 3, None, -2.5), True) == (0, 2, -2, -2.5)  # x <= -2.5 <=> x < -2
    assert finalize_pred((0, 0, 7, 7.0), True) == (0, 0, 7, 7.0)
    assert finalize_pred((0, 0, None, 0.5), False) == (0, 0, 0, 0.5)


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
        Table("b", "x.parquet")
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
