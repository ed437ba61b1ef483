"""The CPU arms bench.py times (oracle/cpu_arms.c) return the same answers as the oracle: a fast wrong baseline would
make every speed-up beside it meaningless."""

import numpy as np
import pytest

from oracle import c_oracle as CO
from oracle import np_oracle as NO


@pytest.mark.parametrize("threads", [1, 3])
def test_filter_arms_match_oracle(threads):
    n = 100003
    cols = [CO.synth_column(NO.F32, dict(kind=0), 42, c, 0, n) for c in range(8)]
    exp = NO.query_filter(cols, [0, 2], [(1, NO.GT, 0, 0.5), (4, NO.LT, 0, 0.5)])
    o0, o1, st, cnt, total = CO.arm_filter_best_f32(cols, 1, 0.5, 4, 0.5, 0, 2, threads)
    assert total == len(exp[0])
    assert np.array_equal(CO.chunks_concat(o0, st, cnt), exp[0]) and np.array_equal(CO.chunks_concat(o1, st, cnt), exp[1])
    db = np.ascontiguousarray(np.stack(cols, axis=1))
    o, st, cnt, total = CO.arm_filter_rowmajor_f32(db, 1, 0.5, 4, 0.5, 0, 2, threads)
    got = CO.chunks_concat(o, st, cnt)
    assert total == len(exp[0]) and np.array_equal(got[:, 0], exp[0]) and np.array_equal(got[:, 1], exp[1])


@pytest.mark.parametrize("threads", [1, 4])
def test_groupby_arms_match_oracle(threads):
    n = 50021
    key = CO.synth_column(NO.I32, dict(kind=0, lo=-7, range=300), 42, 0, 0, n)
    val = CO.synth_column(NO.I32, dict(kind=0, lo=-50, range=1000), 42, 1, 0, n)
    fval = CO.synth_column(NO.F32, dict(kind=0, flo=0.0, fhi=1.0), 42, 1, 0, n)
    exp = NO.query_groupby_ex([key, val], 0, [1, 1], [NO.AGG_SUM, NO.AGG_COUNT])
    cnt, s = CO.arm_groupby_best(key, val, -7, 300, threads)
    keep = cnt > 0
    assert np.array_equal(np.arange(-7, 293)[keep], exp[0]) and np.array_equal(cnt[keep], exp[2])
    assert np.array_equal(s[keep].astype(np.int32), exp[1])
    cntf, sf = CO.arm_groupby_best(key, fval, -7, 300, threads)
    expf = NO.query_groupby_ex([key, fval], 0, [1], [NO.AGG_SUM])
    assert np.allclose(sf[cntf > 0], expf[1].astype(np.float64), rtol=1e-5)
    rows = np.stack([key.view(np.uint32), val.view(np.uint32)], axis=1)
    got = CO.arm_groupby_ref_mt(rows, [2], threads)
    assert np.array_equal(got, CO.query_groupby(rows, 0, [1], [2]))


@pytest.mark.parametrize("bits,threads", [(8, 1), (8, 3), (1, 2)])
def test_orderby_arm_matches_oracle(bits, threads):
    n = 30011
    a = CO.synth_column(NO.I64, dict(kind=0, lo=-50, range=100), 42, 0, 0, n)
    b = CO.synth_column(NO.I64, dict(kind=0, lo=0, range=0), 42, 1, 0, n)
    ga, gb = CO.arm_orderby_i64x2(a, b, bits, threads)
    exp = NO.query_orderby([a, b], [0, 1], [0, 1])
    assert np.array_equal(ga, exp[0]) and np.array_equal(gb, exp[1])


@pytest.mark.parametrize("threads", [1, 4])
def test_join_groupby_arm_matches_oracle(threads):
    nd, nf = 1009, 40009
    pk = CO.synth_column(NO.I32, dict(kind=1, a=48271, b=11, range=nd), 7, 0, 0, nd)
    attr = CO.synth_column(NO.I32, dict(kind=0, lo=-3, range=17), 7, 1, 0, nd)
    fk = CO.synth_column(NO.I32, dict(kind=0, lo=0, range=2 * nd), 42, 0, 0, nf)
    val = CO.synth_column(NO.I32, dict(kind=0, lo=0, range=1000), 42, 1, 0, nf)
    exp = NO.join_groupby([fk, val], [pk, attr], 0, 0, 1, [1, 1], [NO.AGG_SUM, NO.AGG_COUNT])
    g, s, c = CO.arm_join_groupby_best(fk, val, pk, attr, threads)
    assert np.array_equal(g, exp[0]) and np.array_equal(s.astype(np.int32), exp[1]) and np.array_equal(c, exp[2])
