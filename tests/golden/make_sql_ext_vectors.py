#!/usr/bin/env python
"""Pins the EXTENSION semantics (WHERE / HAVING / ORDER BY / COUNT / AVG / JOIN / DISTINCT / LIMIT — clauses the reference
parses but never executes) to something that is not ours: every statement below is run through sqlite3 (Python's
stdlib) on small seeded tables, and the rows sqlite returns are committed as tests/golden/sql_ext_vectors.json.
tests/test_sql_ext_golden.py then holds np_oracle (CPU, through the oracle-backed FutharkContext) and the CUDA path (-m gpu)
to these vectors.

    python tests/golden/make_sql_ext_vectors.py        # rewrites tests/golden/sql_ext_vectors.json

Where this framework deliberately differs from sqlite, the case carries the sqlite statement that expresses OUR semantics
(`sqlite_sql`), and the difference is listed in DEVIATIONS (also written into the JSON):
"""
import json
import math
import os
import sqlite3

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

DEVIATIONS = [
    "Row order without ORDER BY is the input row order (what `filter` at select.fut:18 keeps); GROUP BY output is ascending in "
    "the key(s).  sqlite promises neither, so the sqlite side of those cases carries an explicit ORDER BY rowid / key.",
    "ORDER BY is stable with respect to the input row order; the sqlite side appends rowid as the last sort key.",
    "sql() returns the GROUP BY key column(s) in front of the select items (the reference's result shape, "
    "FutharkContext.py:70-71); the vectors hold the select items only.",
    "There are no NULLs: an aggregate without GROUP BY over zero rows returns one row with COUNT = 0 and NaN where SQL says NULL.",
    "NaN handling (not exercised here, sqlite stores NaN as NULL): comparisons with NaN are false except <>, NOT negates the "
    "comparison's result, ORDER BY puts NaN last.",
    "Integer division / arithmetic in predicates is not supported; AVG is a float64 quotient like sqlite's.",
]


def tables():
    rng = np.random.default_rng(20261017)
    n = 2000
    t = {"columns": ["u", "k1", "k2", "v", "w"],
         "rows": np.stack([np.arange(n), rng.integers(-5, 6, n), rng.integers(0, 4, n), rng.integers(-1000, 1001, n),
                           rng.integers(0, 3_000_000, n)], axis=1).tolist()}
    m = 500
    f = {"columns": ["id", "grp", "x", "y"],
         "rows": [[int(i), int(g), float(x), float(y)] for i, g, x, y in
                  zip(range(m), rng.integers(0, 7, m), np.round(rng.random(m) * 100, 6), np.round(rng.normal(size=m), 6))]}
    nd = 40
    d = {"columns": ["pk", "attr", "region"],
         "rows": np.stack([rng.permutation(60)[:nd] - 5, rng.integers(0, 5, nd), rng.integers(0, 3, nd)], axis=1).tolist()}
    return {"t": t, "f": f, "d": d}


# (our statement, sqlite statement or None when identical, number of select items, ordered?)
CASES = [
    ("select k1, v from t where v > 100 and k2 < 2", "select k1, v from t where v > 100 and k2 < 2 order by rowid", 2, True),
    ("select u, v from t where v > 900 or v < -900", "select u, v from t where v > 900 or v < -900 order by rowid", 2, True),
    ("select u from t where not (k1 > 0 or k2 = 3)", "select u from t where not (k1 > 0 or k2 = 3) order by rowid", 1, True),
    ("select u, k1 from t where k1 between -1 and 2 and k2 not in (0, 3)",
     "select u, k1 from t where k1 between -1 and 2 and k2 not in (0, 3) order by rowid", 2, True),
    ("select u from t where k1 in (-5, 5) or (v >= 990 and not k2 < 2)",
     "select u from t where k1 in (-5, 5) or (v >= 990 and not k2 < 2) order by rowid", 1, True),
    ("select u from t where v <> 0 and k1 = -3 and w <= 1500000", "select u from t where v <> 0 and k1 = -3 and w <= 1500000 order by rowid", 1, True),
    # GROUP BY over a column with NEGATIVE values, no WHERE: must be signed (the u32 entry is not taken here)
    ("select k1, min(v), max(v) from t group by k1", "select k1, min(v), max(v) from t group by k1 order by k1", 3, True),
    ("select k1, sum(v), count(v), avg(v) from t group by k1", "select k1, sum(v), count(v), avg(v) from t group by k1 order by k1", 4, True),
    # sums beyond 32 bits do not wrap
    ("select k2, sum(w), count(*) from t group by k2", "select k2, sum(w), count(*) from t group by k2 order by k2", 3, True),
    ("select k1, sum(v), count(v) from t where v > 0 or k2 = 0 group by k1 having count(v) > 100 or sum(v) >= 40000",
     "select k1, sum(v), count(v) from t where v > 0 or k2 = 0 group by k1 having count(v) > 100 or sum(v) >= 40000 order by k1", 3, True),
    ("select k1, k2, sum(v), avg(w), count(*) from t where w < 2500000 group by k1, k2 having count(*) > 40 order by k1 desc, k2",
     None, 5, True),
    ("select k1, count(k1) from t group by k1 order by k1 desc limit 3", None, 2, True),
    ("select u, k1, v from t where v > 800 order by k1 desc, v", "select u, k1, v from t where v > 800 order by k1 desc, v, rowid", 3, True),
    ("select k2, u from t where u < 50 order by k2", "select k2, u from t where u < 50 order by k2, rowid", 2, True),
    ("select sum(v), count(*), avg(w), max(k1), min(v) from t where k2 = 1", None, 5, True),
    ("select count(*) from t", None, 1, True),
    ("select count(*), sum(v) from t where v > 5000", None, 2, True),
    ("select distinct k1 from t", "select distinct k1 from t order by k1", 1, True),
    ("select distinct k2, k1 from t where v < -950 order by k1 desc", "select distinct k2, k1 from t where v < -950 order by k1 desc, k2", 2, True),
    ("select u, v from t where v > 990 limit 4", "select u, v from t where v > 990 order by rowid limit 4", 2, True),
    # float table
    ("select id, x from f where x > 50.5 and y < 0", "select id, x from f where x > 50.5 and y < 0 order by rowid", 2, True),
    ("select grp, sum(x), avg(y), count(*), max(y) from f group by grp", "select grp, sum(x), avg(y), count(*), max(y) from f group by grp order by grp", 5, True),
    ("select id, y from f where grp = 3 order by y desc", "select id, y from f where grp = 3 order by y desc, rowid", 2, True),
    # joins (dim.pk unique, some fact keys without a match, negative keys)
    ("select attr, sum(v), count(*) from t join d on k1 = pk group by attr", "select attr, sum(v), count(*) from t join d on k1 = pk group by attr order by attr", 3, True),
    ("select d.attr, sum(t.w), count(*) as n from t join d on t.k1 = d.pk where t.k2 > 0 and (d.region = 1 or d.region = 2) "
     "group by d.attr having n > 10 order by sum(t.w) desc, d.attr limit 3", None, 3, True),
    ("select t.u, t.v, d.attr from t join d on t.k1 = d.pk where t.v > 950 order by d.attr desc, t.u",
     None, 3, True),
]


def main():
    tabs = tables()
    con = sqlite3.connect(":memory:")
    for name, tb in tabs.items():
        con.execute(f"create table {name} ({', '.join(tb['columns'])})")
        con.executemany(f"insert into {name} values ({', '.join('?' * len(tb['columns']))})", tb["rows"])
    out = []
    for ours, lite, ncol, ordered in CASES:
        rows = con.execute(lite or ours).fetchall()
        rows = [[(float("nan") if v is None else v) for v in r] for r in rows]
        assert all(len(r) == ncol for r in rows), (ours, rows[:2])
        out.append({"sql": ours, "sqlite_sql": lite or ours, "ncol": ncol, "ordered": ordered,
                    "rows": [[("NaN" if isinstance(v, float) and math.isnan(v) else v) for v in r] for r in rows]})
    path = os.path.join(HERE, "sql_ext_vectors.json")
    json.dump({"generator": "tests/golden/make_sql_ext_vectors.py", "sqlite_version": sqlite3.sqlite_version,
               "deviations": DEVIATIONS, "tables": tabs, "cases": out}, open(path, "w"), indent=0)
    print(f"{len(out)} cases -> {path}; rows per case: {[len(c['rows']) for c in out]}")


if __name__ == "__main__":
    main()
