"""The extension semantics, pinned to sqlite3 (tests/golden/sql_ext_vectors.json, made by make_sql_ext_vectors.py):
every statement goes through HarkDB's public API — on CPU with np_oracle as the local operator engine (so the ORACLE is
held to sqlite), and with -m gpu through libhark.so (so the CUDA path is held to the same vectors)."""

import json
import math
import os

import numpy as np
import pandas as pd
import pytest

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "sql_ext_vectors.json")))


def _frames():
    out = {}
    for name, tb in GOLDEN["tables"].items():
        out[name] = pd.DataFrame(tb["rows"], columns=tb["columns"])
    return out


def _check_case(fc, case):
    got = np.asarray(fc.sql(case["sql"]))
    exp = [[(math.nan if v == "NaN" else v) for v in r] for r in case["rows"]]
    ncol = case["ncol"]
    assert got.ndim == 2, case["sql"]
    got = got[:, got.shape[1] - ncol:]                      # grouped results carry the key column(s) in front
    assert got.shape == (len(exp), ncol), (case["sql"], got.shape, len(exp))
    e = np.asarray(exp, dtype=np.float64).reshape(len(exp), ncol)
    g = got.astype(np.float64)
    assert np.allclose(g, e, rtol=1e-12, atol=0, equal_nan=True), (case["sql"], g[:5].tolist(), e[:5].tolist())


@pytest.mark.parametrize("idx", range(len(GOLDEN["cases"])))
def test_oracle_matches_sqlite(idx):
    from harkdb_b200.sharded import ShardedFutharkContext
    from tests.oracle_engine import OracleEngine
    fc = ShardedFutharkContext(engine=OracleEngine())       # world 1, no process group: np_oracle runs every operator
    for name, df in _frames().items():
        fc.create_table(name, df)
    _check_case(fc, GOLDEN["cases"][idx])


def test_oracle_matches_sqlite_non_resident_column_tables():
    """Tables given as dicts of columns and NOT resident: every statement uploads only the columns it names
    (FutharkContext._pruned) and must still answer like sqlite."""
    from harkdb_b200.sharded import ShardedFutharkContext
    from tests.oracle_engine import OracleEngine
    fc = ShardedFutharkContext(engine=OracleEngine())
    fc.resident = False
    for name, df in _frames().items():
        fc.create_table(name, {c: df[c].to_numpy() for c in df.columns})
    for case in GOLDEN["cases"]:
        _check_case(fc, case)


@pytest.mark.gpu
def test_cuda_path_matches_sqlite():
    from tests.gpu_util import need_gpu
    need_gpu()
    from harkdb_b200 import FutharkContext
    fc = FutharkContext()
    for name, df in _frames().items():
        fc.create_table(name, df)
    for case in GOLDEN["cases"]:
        _check_case(fc, case)
    # and with the tables NOT resident (upload per query, per-column dtypes): same answers
    fc2 = FutharkContext(resident=False)
    for name, df in _frames().items():
        fc2.create_table(name, df)
    for case in GOLDEN["cases"]:
        _check_case(fc2, case)
    # dict-of-columns tables, not resident: only the columns a statement names cross PCIe
    fc3 = FutharkContext(resident=False)
    for name, df in _frames().items():
        fc3.create_table(name, {c: df[c].to_numpy() for c in df.columns})
    for case in GOLDEN["cases"]:
        _check_case(fc3, case)


def test_vectors_are_current():
    """The committed vectors are what the generator produces today (sqlite3 is in the stdlib)."""
    import sqlite3
    con = sqlite3.connect(":memory:")
    for name, tb in GOLDEN["tables"].items():
        con.execute(f"create table {name} ({', '.join(tb['columns'])})")
        con.executemany(f"insert into {name} values ({', '.join('?' * len(tb['columns']))})", tb["rows"])
    for case in GOLDEN["cases"]:
        rows = [[("NaN" if v is None else v) for v in r] for r in con.execute(case["sqlite_sql"]).fetchall()]
        assert json.loads(json.dumps(rows)) == case["rows"], case["sql"]
