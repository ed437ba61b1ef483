"""-m gpu parity of the device half of the multi-GPU layer (K8b partition by splitters, order-key sampling, partial
aggregate finalize, the SUMF64 partial) against the oracle stand-in engine, plus ShardedEnv end to end on one rank
(world 1) and, when the test session runs under torchrun with several GPUs, on all of them."""

import os

import numpy as np
import pytest

from oracle import np_oracle as NO
from tests.gpu_util import get_env, need_gpu
from tests.oracle_engine import OracleEngine, OTable

pytestmark = pytest.mark.gpu


def _cols(seed, n):
    rng = np.random.default_rng(seed)
    f = rng.random(n).astype(np.float32)
    if n > 10:
        f[::53] = np.nan
    return [rng.integers(-40, 40, n).astype(np.int32), rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32),
            rng.integers(-2 ** 50, 2 ** 50, n).astype(np.int64), f, rng.random(n).astype(np.float64)]


@pytest.mark.parametrize("n", [0, 1, 5, 4097, 200003])
@pytest.mark.parametrize("keys,desc", [([0], [0]), ([1], [1]), ([0, 2], [0, 1]), ([3, 0], [0, 0]), ([4, 3, 1, 0], [1, 1, 0, 0])])
def test_partition_by_splitters_and_sampling(n, keys, desc):
    env = get_env()
    oe = OracleEngine()
    cols = _cols(n + len(keys), n)
    t = env.from_columns(cols)
    ot = OTable(cols)
    pos = np.unique(np.random.default_rng(1).integers(0, max(n, 1), 64)) if n else np.zeros(0, np.int64)
    got_s = env.sample_order_keys(t, keys, desc, pos)
    exp_s = oe.sample_order_keys(ot, keys, desc, pos) if n else np.zeros((0, len(keys)), np.uint64)
    assert got_s.dtype == np.uint64 and np.array_equal(got_s, exp_s)
    for nparts in (1, 2, 7):
        from harkdb_b200.sharded import pick_splitters
        sp = pick_splitters([exp_s], [1.0], nparts, len(keys))
        r, counts = env.partition_by_splitters(t, keys, desc, sp, nparts)
        er, ecounts = oe.partition_by_splitters(ot, keys, desc, sp, nparts)
        assert counts == ecounts and sum(counts) == n
        for g, e in zip(r.columns(), er.cols):
            assert g.dtype == e.dtype and np.array_equal(g, e, equal_nan=True)      # stable: bit-exact row order
        r.free()
    t.free()


def test_partition_by_splitters_wide_table_uses_rowid_path():
    env = get_env()
    rng = np.random.default_rng(9)
    cols = [rng.integers(0, 1000, 50021).astype(np.int32) for _ in range(12)]
    t = env.from_columns(cols)
    sp = np.array([[np.uint64(250 ^ 0x80000000)], [np.uint64(700 ^ 0x80000000)]], dtype=np.uint64)
    r, counts = env.partition_by_splitters(t, [3], [0], sp, 3)
    er, ecounts = OracleEngine().partition_by_splitters(OTable(cols), [3], [0], sp, 3)
    assert counts == ecounts
    for g, e in zip(r.columns(), er.cols):
        assert np.array_equal(g, e)
    r.free(); t.free()


def test_sumf64_partial_and_finalize():
    env = get_env()
    cols = _cols(5, 100003)
    cols[3] = np.nan_to_num(cols[3])
    t = env.from_columns(cols)
    for impl in (0, 1):
        env.set_option("groupby.impl", impl)
        ops = [NO.AGG_SUMF64, NO.AGG_COUNT, NO.AGG_SUMF64, NO.AGG_SUM, NO.AGG_SUMF64]
        s_cols = [1, 1, 3, 1, 4]
        r = env.query_groupby_ex(t, 0, s_cols, ops)
        exp = NO.query_groupby_ex(cols, 0, s_cols, ops)
        for g, e in zip(r.columns(), exp):
            assert g.dtype == e.dtype
            assert np.allclose(g, e, rtol=1e-12, atol=0) if g.dtype.kind == "f" else np.array_equal(g, e)
        fin = env.groupby_finalize(r, [NO.AGG_AVG, NO.AGG_SUMF64, NO.AGG_SUM, NO.AGG_SUMF64])
        efin = OracleEngine().groupby_finalize(OTable(exp), [NO.AGG_AVG, NO.AGG_SUMF64, NO.AGG_SUM, NO.AGG_SUMF64])
        for g, e in zip(fin.columns(), efin.cols):
            assert g.dtype == e.dtype and (np.allclose(g, e, rtol=1e-12, atol=0) if g.dtype.kind == "f" else np.array_equal(g, e))
        fin.free(); r.free()
    env.set_option("groupby.impl", 0)
    from harkdb_b200.hark_ffi import HarkError
    with pytest.raises(HarkError, match="layout"):
        env.groupby_finalize(t, [NO.AGG_AVG])
    t.free()


def _sharded_env():
    need_gpu()
    import torch
    import torch.distributed as dist
    from harkdb_b200.sharded import HarkEngine, ShardedEnv
    if "RANK" in os.environ and not dist.is_initialized():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
    return ShardedEnv(HarkEngine(int(os.environ.get("LOCAL_RANK", "0"))))


def test_sharded_env_end_to_end_on_the_available_ranks():
    """world 1 under plain pytest; all GPUs under `torchrun -m pytest tests/test_gpu_sharded.py -m gpu`."""
    from tests import test_sharded_gloo as G
    senv = _sharded_env()
    for name in ("filter", "groupby", "groupby_dense", "groupby_multi", "groupby_pinned", "orderby", "join", "sql", "sql_join"):
        getattr(G, "_scn_" + name)(senv)


@pytest.mark.parametrize("peer", ["1", "0"])
def test_two_ranks_when_two_gpus_are_visible(peer):
    """Multi-rank parity where the driver can see it: with >= 2 devices and no torchrun around this session, spawn
    `torch.distributed.run --nproc-per-node 2` over the end-to-end scenario test above — once with the K8c peer-memory
    exchange, once with HARK_PEER=0 (K8b partition + NCCL all-to-all)."""
    need_gpu()
    import subprocess
    import sys
    import torch
    if "RANK" in os.environ:
        pytest.skip("already running under torchrun")
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, HARK_PEER=peer, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29541" if peer == "1" else "29542", "-m", "pytest", "tests/test_gpu_sharded.py", "-m", "gpu", "-q", "-x",
                        "-p", "no:cacheprovider", "-k", "end_to_end"], cwd=root, env=env, capture_output=True, text=True, timeout=900)
    if r.returncode != 0:
        pytest.fail("2-rank run failed:\n" + r.stdout[-6000:] + "\n---- stderr ----\n" + r.stderr[-3000:], pytrace=False)
