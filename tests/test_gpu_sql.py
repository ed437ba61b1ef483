"""End-to-end through HarkDB's public API (FutharkContext.create_table / sql) on the GPU (-m gpu)."""

import json
import os

import numpy as np
import pandas as pd
import pytest

from oracle import np_oracle as NO
from tests.gpu_util import need_gpu

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def fc():
    need_gpu()
    from harkdb_b200 import FutharkContext
    ctx = FutharkContext()
    d = json.load(open(os.path.join(GOLDEN, "data_csv.json")))
    ctx.create_table("game_1", pd.DataFrame(d["rows"], columns=d["columns"]))
    return ctx


def test_readme_select(fc):
    # README.md:42 / BASELINE config 1
    out = fc.sql("select col1, col3 from game_1")
    assert out.tolist() == [[6, 6], [0, 0], [0, 0], [0, 0], [0, 0], [6, 6], [1, 3]]


def test_statements_the_reference_crashes_on(fc):
    assert fc.sql("select col8 from game_1").tolist() == [[6], [0], [0], [0], [0], [6], [1]]
    assert fc.sql("select * from game_1").shape == (7, 8)


def test_csv_file_and_upload_per_query(tmp_path):
    need_gpu()
    from harkdb_b200 import FutharkContext
    d = json.load(open(os.path.join(GOLDEN, "data_csv.json")))
    p = tmp_path / "data.csv"
    pd.DataFrame(d["rows"], columns=d["columns"]).to_csv(p, index=False)
    ctx = FutharkContext(resident=False)       # the reference's behaviour: upload on every query
    ctx.create_table("game_1", str(p))
    assert ctx.sql("select col1, col3 from game_1").tolist() == [[6, 6], [0, 0], [0, 0], [0, 0], [0, 0], [6, 6], [1, 3]]
    ctx.drop_table("game_1")
    with pytest.raises(Exception, match="game_1 is not in tables"):
        ctx.sql("select col1 from game_1")


def test_where(fc):
    out = fc.sql("select col1, col3 from game_1 where col2 > 0 and col5 < 6")
    assert out.tolist() == [[1, 3]]
    out = fc.sql("select col1 from game_1 where col2 >= 0.5")
    assert out.tolist() == [[6], [6], [1]]
    assert fc.sql("select col1 from game_1 where col1 > 100").shape == (0, 1)


def test_where_or_not_between_in(fc):
    d = json.load(open(os.path.join(GOLDEN, "data_csv.json")))
    a = np.asarray(d["rows"])
    c = {f"col{i + 1}": a[:, i] for i in range(a.shape[1])}
    cases = [
        ("col1 > 5 or col3 = 3", (c["col1"] > 5) | (c["col3"] == 3)),
        ("not (col1 > 5 or col3 = 3)", ~((c["col1"] > 5) | (c["col3"] == 3))),
        ("col1 between 1 and 6 and col2 not in (0, 6)", (c["col1"] >= 1) & (c["col1"] <= 6) & ~np.isin(c["col2"], [0, 6])),
        ("col1 in (0, 1) or (col2 >= 1 and not col5 < 6)", np.isin(c["col1"], [0, 1]) | ((c["col2"] >= 1) & ~(c["col5"] < 6))),
        ("col1 > 0.5 or col2 < -0.5", (c["col1"] > 0.5) | (c["col2"] < -0.5)),
    ]
    for w, m in cases:
        out = fc.sql(f"select col1, col3 from game_1 where {w}")
        assert out.tolist() == a[m][:, [0, 2]].tolist(), w
    out = fc.sql("select col1, sum(col2), count(col2) from game_1 where col2 > 0 or col1 = 0 group by col1 "
                 "having count(col2) > 1 or sum(col2) >= 2 order by col1 desc")
    m = (c["col2"] > 0) | (c["col1"] == 0)
    exp = []
    for k in sorted(set(a[m][:, 0].tolist()), reverse=True):
        v = a[m][a[m][:, 0] == k][:, 1]
        if len(v) > 1 or v.sum() >= 2:
            exp.append([k, k, int(v.sum()), len(v)])      # reference shape: key column, then one column per select item
    assert out.tolist() == exp


def test_where_or_float_table_nan():
    need_gpu()
    from harkdb_b200 import FutharkContext
    rng = np.random.default_rng(8)
    a = rng.random((200001, 4)).astype(np.float32)
    a[rng.integers(0, len(a), 1000), 1] = np.nan
    ctx = FutharkContext()
    ctx.create_table("t", pd.DataFrame(a, columns=["a", "b", "c", "d"]))
    out = ctx.sql("select a, d from t where (a < 0.25 or not b > 0.5) and c between 0.1 and 0.9")
    with np.errstate(invalid="ignore"):
        m = ((a[:, 0] < np.float32(0.25)) | ~(a[:, 1] > np.float32(0.5))) & (a[:, 2] >= np.float32(0.1)) & (a[:, 2] <= np.float32(0.9))
    assert np.array_equal(out, a[m][:, [0, 3]])


def test_tpch_q1_shape_mixed_dtypes_group_by_two_columns():
    """TPC-H Q1's shape on a frame with integer keys next to float measures: per-column dtypes at create_table,
    WHERE, GROUP BY two columns, SUM / AVG / COUNT, HAVING, ORDER BY."""
    need_gpu()
    from harkdb_b200 import FutharkContext
    rng = np.random.default_rng(19)
    n = 300007
    df = pd.DataFrame({"flag": rng.integers(0, 3, n), "status": rng.integers(0, 2, n), "qty": rng.integers(1, 51, n),
                       "price": rng.random(n) * 1000, "shipdate": rng.integers(8000, 10600, n)})
    ctx = FutharkContext()
    ctx.create_table("lineitem", df)
    assert ctx.tables["lineitem"].get_handle().dtypes == [NO.I32, NO.I32, NO.I32, NO.F64, NO.I32]
    out = ctx.sql("select flag, status, sum(qty), avg(price), count(*) from lineitem where shipdate <= 10471 "
                  "group by flag, status having count(*) > 10 order by flag desc, status")
    g = df[df.shipdate <= 10471].groupby(["flag", "status"], sort=True).agg(q=("qty", "sum"), p=("price", "mean"),
                                                                            c=("qty", "size")).reset_index()
    g = g[g.c > 10].sort_values(["flag", "status"], ascending=[False, True], kind="stable")
    assert out.shape == (len(g), 7)                       # 2 key columns + the 5 select items
    assert np.array_equal(out[:, 0], g.flag) and np.array_equal(out[:, 1], g.status)
    assert np.array_equal(out[:, 2], g.flag) and np.array_equal(out[:, 3], g.status)
    assert np.array_equal(out[:, 4], g.q) and np.allclose(out[:, 5], g.p, rtol=1e-12) and np.array_equal(out[:, 6], g.c)
    # TPC-H Q6's shape: filter + aggregates without GROUP BY
    out = ctx.sql("select sum(price), count(*), avg(qty), min(shipdate) from lineitem "
                  "where shipdate between 9000 and 9364 and qty < 24 and not flag = 1")
    m = df.shipdate.between(9000, 9364) & (df.qty < 24) & (df.flag != 1)
    assert out.shape == (1, 4)
    assert np.isclose(out[0, 0], df.price[m].sum(), rtol=1e-12) and out[0, 1] == m.sum()
    assert np.isclose(out[0, 2], df.qty[m].mean(), rtol=1e-12) and out[0, 3] == df.shipdate[m].min()
    assert ctx.sql("select count(*) from lineitem where qty > 1000").tolist() == [[0.0]]     # one row, like SQL
    # single key on the same frame: integer key, float measure
    out = ctx.sql("select flag, max(price) from lineitem group by flag")
    gm = df.groupby("flag", sort=True).price.max()
    assert np.array_equal(out[:, 0], gm.index) and np.array_equal(out[:, 2], gm.to_numpy())


def test_where_float_table():
    need_gpu()
    from harkdb_b200 import FutharkContext
    rng = np.random.default_rng(5)
    a = rng.random((100000, 8)).astype(np.float32)
    ctx = FutharkContext()
    ctx.create_table("t", pd.DataFrame(a, columns=[f"col{i}" for i in range(1, 9)]))
    out = ctx.sql("SELECT col1,col3 FROM t WHERE col2 > 0.5 AND col5 < 0.25")
    m = (a[:, 1] > np.float32(0.5)) & (a[:, 4] < np.float32(0.25))
    assert out.dtype == np.float32 and np.array_equal(out, a[m][:, [0, 2]])


def test_sql_device_result_stays_resident(fc):
    """sql(..., device=True): the result is a DeviceTable that further entries consume without a download."""
    from harkdb_b200.hark_ffi import DeviceTable
    r = fc.sql("select col1, col3 from game_1 where col1 > 0", device=True)
    assert isinstance(r, DeviceTable) and r.shape == (3, 2)
    assert r.to_numpy().tolist() == [[6, 6], [6, 6], [1, 3]]
    g = fc.FutEnv.query_groupby_ex(r, 0, [1], [NO.AGG_COUNT])          # chained on the device
    assert [c.tolist() for c in g.columns()] == [[1, 6], [1, 2]]
    g.free(); r.free()
    r = fc.sql("select col1, count(col1) from game_1 group by col1 order by col1 desc limit 2", device=True)
    assert isinstance(r, DeviceTable) and r.shape[0] == 2
    r.free()
    assert isinstance(fc.sql("select col1 from game_1"), np.ndarray)    # default unchanged
