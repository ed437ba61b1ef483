#!/usr/bin/env python
"""Regenerate tests/golden/*.json.  Runs ONLY in the build container (needs /root/reference).

  python tests/golden/make_golden.py

1. segmented_kats.json — the reference's own known-answer tests, parsed out of
   futhark/lib/github.com/diku-dk/segmented/segmented_tests.fut (`-- input {..}` / `-- output {..}`
   comment blocks, :5-72).  These are the only golden vectors the reference holds for this path.
2. data_csv.json — the reference fixture data.csv (:1-8) as loaded by pandas (what table.py:26-28 does).
3. harkdb_vectors.json — query_sel / query_groupby / join outputs on data.csv and on seeded random
   tables, computed by oracle/hark_ref.py, the line-by-line SOAC simulation of select.fut / groupby.fut /
   join.fut.  The Futhark compiler is not available, so these are OUR reading of the source, not
   outputs of the reference ("parity unpinned by reference tests").
"""

import json
import os
import random
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def parse_value(tok: str):
    tok = tok.strip()
    m = re.fullmatch(r"empty\(\[0\](\w+)\)", tok)
    if m:
        return []
    body = tok.strip()[1:-1].strip()
    if not body:
        return []
    out = []
    for x in body.split(","):
        x = x.strip()
        out.append(True if x == "true" else False if x == "false" else int(x))
    return out


def split_values(s: str):
    """Split '{ [..] [..] }' or '{ empty([0]bool) empty([0]i32) }' into value tokens."""
    s = s.strip()
    assert s[0] == "{" and s[-1] == "}", s
    s = s[1:-1].strip()
    return re.findall(r"empty\(\[0\]\w+\)|\[[^\]]*\]", s)


def parse_kats(path: str):
    text = open(path).read()
    kats = {}
    entry = None
    pending_input = None
    # join comment lines, then walk tokens "entry:", "input {...}", "output {...}"
    comment = "\n".join(l[2:].strip() for l in text.splitlines() if l.startswith("--"))
    for m in re.finditer(r"entry:\s*(\w+)|input\s*(\{[^}]*\})|output\s*(\{[^}]*\})", comment):
        if m.group(1):
            entry = m.group(1)
            kats.setdefault(entry, [])
        elif m.group(2):
            pending_input = [parse_value(v) for v in split_values(m.group(2))]
        elif m.group(3):
            outv = [parse_value(v) for v in split_values(m.group(3))]
            kats[entry].append({"input": pending_input, "output": outv[0]})
            pending_input = None
    return kats


def main():
    import pandas as pd
    from oracle import hark_ref as R

    kats = parse_kats(os.path.join(REF, "futhark/lib/github.com/diku-dk/segmented/segmented_tests.fut"))
    n_cases = sum(len(v) for v in kats.values())
    assert n_cases == 18, n_cases   # 3+2+6+4+1+1+1
    json.dump({"source": "segmented_tests.fut:5-72 (reference's own KATs)", "kats": kats},
              open(os.path.join(HERE, "segmented_kats.json"), "w"), indent=1)

    df = pd.read_csv(os.path.join(REF, "data.csv"))
    data = {"source": "data.csv:1-8 via pandas.read_csv (table.py:26-28)", "columns": df.columns.tolist(),
            "dtype": str(df.values.dtype), "rows": df.values.tolist()}
    json.dump(data, open(os.path.join(HERE, "data_csv.json"), "w"), indent=1)
    db = data["rows"]

    vec = {"source": "oracle/hark_ref.py (SOAC-level simulation of select.fut/groupby.fut/join.fut); NOT reference output",
           "cases": []}

    def add(kind, args, fn):
        try:
            out = fn()
            vec["cases"].append({"kind": kind, **args, "output": out})
        except IndexError as e:
            vec["cases"].append({"kind": kind, **args, "error": "index"})

    # BASELINE config 1 / README example, and the test.py query through parse.py's plan
    add("query_sel", {"db": "data_csv", "cols": [0, 2]}, lambda: R.query_sel(db, [0, 2]))
    add("query_sel", {"db": "data_csv", "cols": [7, 7, 0, 3]}, lambda: R.query_sel(db, [7, 7, 0, 3]))
    add("query_sel", {"db": "data_csv", "cols": []}, lambda: R.query_sel(db, []))
    add("query_sel", {"db": "data_csv", "cols": [8]}, lambda: R.query_sel(db, [8]))
    add("query_groupby", {"db": "data_csv", "g_col": 0, "s_cols": [0, 2], "t_cols": [0, 3]},
        lambda: R.query_groupby(db, 0, [0, 2], [0, 3]))
    for t in (0, 1, 2, 3, 4, 9):
        add("query_groupby", {"db": "data_csv", "g_col": 0, "s_cols": [1], "t_cols": [t]},
            lambda t=t: R.query_groupby(db, 0, [1], [t]))
    add("query_groupby", {"db": "data_csv", "g_col": 0, "s_cols": [2, 7, 4], "t_cols": [2, 4, 1]},
        lambda: R.query_groupby(db, 0, [2, 7, 4], [2, 4, 1]))
    add("query_groupby", {"db": "data_csv", "g_col": 5, "s_cols": [], "t_cols": []},
        lambda: R.query_groupby(db, 5, [], []))
    db2 = [[6, 60], [1, 10], [7, 70], [6, 61]]
    add("join", {"db1": "data_csv", "db2": db2, "col1": 0, "col2": 0, "cols1": [0, 2], "cols2": [1]},
        lambda: R.join(db, db2, 0, 0, [0, 2], [1]))
    add("join", {"db1": db2, "db2": "data_csv", "col1": 0, "col2": 7, "cols1": [1], "cols2": [0, 4]},
        lambda: R.join(db2, db, 0, 7, [1], [0, 4]))

    # seeded random tables incl. keys >= 2^31 (unsigned order) and u32 wrap-around in sum/prod
    rng = random.Random(20261017)
    for case in range(6):
        n = [1, 5, 17, 33, 40, 64][case]
        m = rng.randint(2, 5)
        hi = [3, 7, 2 ** 32 - 1, 2 ** 32 - 1, 5, 2 ** 31 + 3][case]
        lo = [0, 0, 2 ** 32 - 4, 0, 0, 2 ** 31 - 3][case]
        tbl = [[rng.randint(lo, hi) if c == 0 else rng.randint(0, 2 ** 32 - 1) for c in range(m)] for _ in range(n)]
        s_cols = [rng.randrange(m) for _ in range(rng.randint(1, 4))]
        t_cols = [rng.choice([0, 1, 2, 3, 4]) for _ in s_cols]
        vec["cases"].append({"kind": "query_groupby", "db": tbl, "g_col": 0, "s_cols": s_cols, "t_cols": t_cols,
                             "output": R.query_groupby(tbl, 0, s_cols, t_cols)})
        cols = [rng.randrange(m) for _ in range(rng.randint(0, 4))]
        vec["cases"].append({"kind": "query_sel", "db": tbl, "cols": cols, "output": R.query_sel(tbl, cols)})
    for case in range(5):
        n, s = [(0, 4), (6, 0), (9, 7), (20, 20), (30, 12)][case]
        m, t = rng.randint(1, 4), rng.randint(1, 4)
        kmax = [3, 3, 4, 2 ** 32 - 1, 6][case]
        klo = [0, 0, 0, 2 ** 32 - 5, 0][case]
        a = [[rng.randint(klo, kmax) for _ in range(m)] for _ in range(n)]
        b = [[rng.randint(klo, kmax) for _ in range(t)] for _ in range(s)]
        c1, c2 = rng.randrange(m), rng.randrange(t)
        cols1 = [rng.randrange(m) for _ in range(rng.randint(0, 3))]
        cols2 = [rng.randrange(t) for _ in range(rng.randint(1, 3))]
        vec["cases"].append({"kind": "join", "db1": a, "db2": b, "col1": c1, "col2": c2, "cols1": cols1,
                             "cols2": cols2, "output": R.join(a, b, c1, c2, cols1, cols2)})
    json.dump(vec, open(os.path.join(HERE, "harkdb_vectors.json"), "w"), indent=None, separators=(",", ":"))
    print("wrote", n_cases, "KATs and", len(vec["cases"]), "harkdb vectors")


if __name__ == "__main__":
    main()
