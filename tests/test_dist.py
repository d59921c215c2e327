"""Host-side logic of the column-sharded multi-GPU path, on CPU with gloo (world_size 2 and 3):
the column deal is a partition and the final all-gather reassembles W exactly."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import slimtest as st
from slim_b200.dist import all_gather_columns, shard_columns


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_model(ncols, seed=0):
    rng = np.random.default_rng(seed)
    cnt = rng.integers(0, 9, size=ncols)
    cnt[rng.integers(0, ncols, size=3)] = 0  # a few empty columns
    colptr = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    colind = np.concatenate([np.sort(rng.choice(ncols, c, replace=False)) for c in cnt] +
                            [np.zeros(0, np.int64)]).astype(np.int32)
    colval = rng.random(int(colptr[-1])).astype(np.float32)
    return colptr, colind, colval


def _worker(rank, world, port, ncols, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        colptr, colind, colval = _fake_model(ncols)
        colcnt = np.random.default_rng(5).integers(1, 1000, size=ncols)
        cols = np.arange(ncols, dtype=np.int32)
        mine = shard_columns(cols, colcnt, rank, world)
        counts = torch.tensor(np.diff(colptr)[mine], dtype=torch.int32)
        ind = torch.tensor(np.concatenate([colind[colptr[j]:colptr[j + 1]] for j in mine] +
                                          [np.zeros(0, np.int32)]), dtype=torch.int32)
        val = torch.tensor(np.concatenate([colval[colptr[j]:colptr[j + 1]] for j in mine] +
                                          [np.zeros(0, np.float32)]), dtype=torch.float32)
        cp, ci, cv = all_gather_columns(mine, counts, ind, val, ncols)
        ok = (np.array_equal(cp, colptr) and np.array_equal(ci, colind) and
              np.array_equal(cv.view(np.uint32), colval.view(np.uint32)))
        q.put((rank, bool(ok), len(mine)))
    except Exception as e:  # surface the failure instead of letting the parent time out
        q.put((rank, False, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,ncols", [(2, 101), (3, 40), (2, 1)])
def test_all_gather_reassembles_model(world, ncols):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ncols, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res
    assert sum(n for _, _, n in res) == ncols


def test_shard_columns_is_a_balanced_partition():
    rng = np.random.default_rng(1)
    colcnt = (rng.pareto(1.1, size=5000) * 50).astype(np.int64) + 1
    cols = np.sort(rng.choice(5000, 1000, replace=False)).astype(np.int32)
    for world in (1, 2, 4, 8):
        parts = [shard_columns(cols, colcnt, r, world) for r in range(world)]
        allidx = np.sort(np.concatenate(parts))
        assert np.array_equal(allidx, np.arange(len(cols)))
        loads = np.array([colcnt[cols[p]].sum() for p in parts], dtype=np.float64)
        # dealing in descending-nnz order keeps the shards within a head column of each other
        assert loads.max() - loads.min() <= colcnt[cols].max()


def test_stratified_sample_mix():
    from slim_b200.synth import stratified_columns

    cnt = np.sort((np.random.default_rng(2).pareto(1.0, 10000) * 20).astype(np.int64))[::-1].copy()
    a = stratified_columns(cnt, 100, 0)
    b = stratified_columns(cnt, 100, 1)
    assert len(set(a) & set(b)) == 0 and len(a) == 100
    # mid-stratum comb: the sample mean tracks the population mean
    assert abs(cnt[a].mean() - cnt.mean()) < 0.5 * cnt.mean()
