"""World-size-2 CPU (gloo) test of the multi-GPU host logic: contiguous sharding of the
thrown-event index range, the one all-reduce of the acceptance sums, and the finalize
arithmetic (host entry point of the C ABI).  No GPU compute."""
import ctypes
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nuspacesim_b200 import _lib
    from nuspacesim_b200.compute import shard_range
    from nuspacesim_b200.config import NssConfig
    from nuspacesim_b200.params import params_from_config

    lo, cnt = shard_range(n_total, rank, world)
    # synthetic per-event factors, a pure function of the global thrown index
    idx = np.arange(lo, lo + cnt, dtype=np.float64)
    valid = (idx % 11) != 0
    f = np.where(valid & ((idx % 7) == 0), 1e-3 * (1 + (idx % 13)), 0.0)
    geo = np.where(valid, 0.5 + (idx % 5) * 0.1, 0.0)
    sums = torch.tensor([geo.sum(), f.sum(), (f * f).sum(), np.count_nonzero(f), valid.sum(),
                         cnt, 0.0, 0.0], dtype=torch.float64)
    dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    p = params_from_config(NssConfig())
    s = sums.numpy().copy()
    r = np.zeros(4)
    rc = _lib.lib().nss_acceptance_finalize(ctypes.byref(p), s.ctypes.data_as(_lib.c_f64p),
                                            r.ctypes.data_as(_lib.c_f64p))
    q.put((rank, lo, cnt, s.tolist(), r.tolist(), rc, p.mcnorm))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_sharded_acceptance_sums():
    world, n_total = 2, 100_003
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=150) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (r0, lo0, c0, s0, f0, rc0, mcnorm), (r1, lo1, c1, s1, f1, rc1, _) = res
    assert rc0 == rc1 == 0
    assert lo0 == 0 and lo0 + c0 == lo1 and lo1 + c1 == n_total
    assert s0 == s1 and f0 == f1  # every rank holds the same all-reduced result
    idx = np.arange(n_total, dtype=np.float64)
    valid = (idx % 11) != 0
    f = np.where(valid & ((idx % 7) == 0), 1e-3 * (1 + (idx % 13)), 0.0)
    np.testing.assert_allclose(f0[0], f.sum() * mcnorm / n_total, rtol=1e-12)
    fv = f[valid]
    np.testing.assert_allclose(f0[3], np.sqrt(np.var(fv, ddof=1) / n_total) * mcnorm, rtol=1e-9)
    assert f0[2] == np.count_nonzero(f)


def _merge_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import importlib

    C = importlib.import_module("nuspacesim_b200.compute")
    rng = np.random.default_rng(100 + rank)
    # two "slices" per rank of per-event factors f (valid events only)
    rows = []
    for _ in range(2):
        f = rng.random(1000 + 37 * rank) * (rng.random(1000 + 37 * rank) < 0.2)
        rows.append([2 * f.sum(), f.sum(), ((f - f.mean()) ** 2).sum(), np.count_nonzero(f), f.size])
    mine = C._merge_sums(np.array(rows))
    tot = C._merge_sums(C._gather_rank_sums(mine))
    q.put((rank, tot.tolist()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_column_mode_sum_merge():
    """compute(shard=True)'s acceptance merge: per-slice sums -> per-rank -> global, with the
    squared deviations combined exactly (the GPU test covers the columns)."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_merge_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=150) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert res[0][1] == res[1][1]
    fs = []
    for rank in range(world):
        rng = np.random.default_rng(100 + rank)
        for _ in range(2):
            fs.append(rng.random(1000 + 37 * rank) * (rng.random(1000 + 37 * rank) < 0.2))
    f = np.concatenate(fs)
    tot = np.array(res[0][1])
    np.testing.assert_allclose(tot[1], f.sum(), rtol=1e-13)
    np.testing.assert_allclose(tot[2], ((f - f.mean()) ** 2).sum(), rtol=1e-11)
    assert tot[3] == np.count_nonzero(f) and tot[4] == f.size
