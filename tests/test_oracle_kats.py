"""Oracle vs the reference's own known-answer tests (segmented_tests.fut:5-72) and vs the committed
vectors of the SOAC-level simulation.  CPU only."""

import json
import os

import numpy as np
import pytest

from oracle import c_oracle as CO
from oracle import hark_ref as R
from oracle import np_oracle as NO

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
KATS = json.load(open(os.path.join(GOLDEN, "segmented_kats.json")))["kats"]
DATA = json.load(open(os.path.join(GOLDEN, "data_csv.json")))
VEC = json.load(open(os.path.join(GOLDEN, "harkdb_vectors.json")))["cases"]

add = lambda a, b: a + b
ident = lambda x: x
mul = lambda x, i: x * i

SIM = {
    "test_segmented_scan": lambda f, a: R.segmented_scan(add, 0, f, a),
    "test_segmented_reduce": lambda f, a: R.segmented_reduce(add, 0, f, a),
    "test_replicated_iota": lambda r: R.replicated_iota(r),
    "test_segmented_iota": lambda f: R.segmented_iota(f),
    "test_expand": lambda a: R.expand(ident, mul, a),
    "test_expand_reduce": lambda a: R.expand_reduce(ident, mul, add, 0, a),
    "test_expand_outer_reduce": lambda a: R.expand_outer_reduce(ident, mul, add, 0, a),
}
CPORT = {
    "test_segmented_scan": CO.segmented_scan_add,
    "test_segmented_reduce": CO.segmented_reduce_add,
    "test_replicated_iota": CO.replicated_iota,
    "test_segmented_iota": CO.segmented_iota,
    "test_expand": CO.expand_mul,
    "test_expand_reduce": CO.expand_reduce_mul_add,
    "test_expand_outer_reduce": CO.expand_outer_reduce_mul_add,
}

ALL_KATS = [(e, i) for e, cases in KATS.items() for i in range(len(cases))]


def test_kat_inventory():
    # 3 + 2 + 6 + 4 + 1 + 1 + 1 cases, every entry of segmented_tests.fut
    assert sorted(KATS) == sorted(SIM)
    assert len(ALL_KATS) == 18


@pytest.mark.parametrize("entry,i", ALL_KATS)
def test_soac_simulation_matches_reference_kats(entry, i):
    case = KATS[entry][i]
    assert list(SIM[entry](*case["input"])) == case["output"]


@pytest.mark.parametrize("entry,i", ALL_KATS)
def test_c_port_matches_reference_kats(entry, i):
    case = KATS[entry][i]
    assert CPORT[entry](*case["input"]).tolist() == case["output"]


def _db(ref):
    return DATA["rows"] if ref == "data_csv" else ref


def test_data_csv_fixture_shape():
    assert DATA["columns"] == [f"col{i}" for i in range(1, 9)]
    assert np.asarray(DATA["rows"]).shape == (7, 8)
    assert DATA["dtype"] == "int64"      # what pandas hands the reference (table.py:28)


def test_readme_example_vector():
    # README.md:42 / BASELINE config 1: select col1, col3 from game_1
    out = NO.query_sel(np.asarray(DATA["rows"], dtype=np.int32), [0, 2])
    assert out.tolist() == [[6, 6], [0, 0], [0, 0], [0, 0], [0, 0], [6, 6], [1, 3]]


@pytest.mark.parametrize("idx", range(len(VEC)))
def test_fast_oracles_match_simulation_vectors(idx):
    case = VEC[idx]
    kind = case["kind"]
    if kind == "query_sel":
        db = np.asarray(_db(case["db"]), dtype=np.int64).astype(np.uint32).view(np.int32).reshape(len(_db(case["db"])), -1)
        if "error" in case:
            with pytest.raises(IndexError):
                NO.query_sel(db, case["cols"])
            with pytest.raises(IndexError):
                CO.query_sel(db, case["cols"])
            return
        exp = np.asarray(case["output"], dtype=np.int64).astype(np.uint32).view(np.int32).reshape(db.shape[0], len(case["cols"]))
        assert np.array_equal(NO.query_sel(db, case["cols"]), exp)
        assert np.array_equal(CO.query_sel(db, case["cols"]), exp)
    elif kind == "query_groupby":
        db = np.asarray(_db(case["db"]), dtype=np.int64).astype(np.uint32)
        exp = np.asarray(case["output"], dtype=np.int64).astype(np.uint32).reshape(-1, len(case["s_cols"]) + 1)
        assert np.array_equal(NO.query_groupby(db, case["g_col"], case["s_cols"], case["t_cols"]), exp)
        assert np.array_equal(CO.query_groupby(db, case["g_col"], case["s_cols"], case["t_cols"]), exp)
    else:
        db1 = np.asarray(_db(case["db1"]), dtype=np.int64).astype(np.uint32)
        db2 = np.asarray(_db(case["db2"]), dtype=np.int64).astype(np.uint32)
        w = len(case["cols1"]) + len(case["cols2"])
        db1 = db1.reshape(db1.shape[0], -1) if db1.size else np.zeros((0, 1), np.uint32)
        db2 = db2.reshape(db2.shape[0], -1) if db2.size else np.zeros((0, 1), np.uint32)
        exp = np.asarray(case["output"], dtype=np.int64).astype(np.uint32).reshape(-1, w)
        args = (case["col1"], case["col2"], case["cols1"], case["cols2"])
        assert np.array_equal(NO.join(db1, db2, *args), exp)
        assert np.array_equal(CO.join(db1, db2, *args), exp)
