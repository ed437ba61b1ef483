#!/usr/bin/env python
"""Bisects a large-n ORDER BY failure: which n / key layout loses elements (on-device multiset checks)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tools.ops_bench import as_torch
from harkdb_b200 import hark_ffi
env = hark_ffi.Futhark()
I64 = 2
def check(n, keycols, specs, tag):
    t = env.synth(n, [I64, I64], specs, seed=42)
    a0, b0 = as_torch(t, 0), as_torch(t, 1)
    s1, s2 = int(a0.sum().item()), int(b0.sum().item())
    r = env.query_orderby(t, [0, 1], keycols)
    a, b = as_torch(r, 0), as_torch(r, 1)
    ok1, ok2 = int(a.sum().item()) == s1, int(b.sum().item()) == s2
    msg = {"tag": tag, "n": n, "keys": keycols, "sum_a_ok": ok1, "sum_b_ok": ok2}
    if not (ok1 and ok2):
        # where do the sorted columns differ from torch's own sort of the key column?
        col = 1 if not ok2 else 0
        ref, _ = torch.sort(as_torch(t, col))
        got, _ = torch.sort(as_torch(r, col))
        diff = (ref != got).nonzero()
        msg["n_diff_sorted_multiset"] = int(diff.numel())
        if diff.numel():
            i = int(diff[0].item()); msg["first_diff"] = [i, int(ref[i].item()), int(got[i].item())]
        del ref, got, diff
    print(json.dumps(msg), flush=True)
    del a, b, a0, b0
    r.free(); t.free()
full = dict(kind=0, lo=0, range=0); small = dict(kind=0, lo=-(2 ** 19), range=2 ** 20)
for n in [1 << 27, 1 << 29, 1 << 30, (1 << 30) + (1 << 29)]:
    check(n, [0, 1], [small, full], "cfg4")
check(1 << 30, [1], [small, full], "only-col2")
check(1 << 30, [0], [small, full], "only-col1")
check(1 << 28, [1, 0], [small, full], "col2-then-col1")
