"""Multi-GPU path inside libslim.so (gather.cuh): NCCL communicators, the variable-length all-gather of W with
device-side reassembly, the GPU-built CSR index of the model, and SLIM_Learn sharding over the visible devices.
World size 1 runs on any GPU box (the collective, the placement kernels and the sharded code path of SLIM_Learn are
all exercised with one rank); the 2-rank tests need two GPUs and are skipped otherwise."""
import ctypes as C
import os
import socket
from pathlib import Path

import numpy as np
import pytest

import slimtest as st

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    from slim_b200 import _lib

    L = _lib.load()
    if L.SLIMB200_DeviceCount() < 1:
        pytest.skip("needs a CUDA device")
    return L


@pytest.fixture(scope="module")
def ours():
    return st.SlimLib(ROOT / "slim_b200" / "lib" / "libslim.so")


def _solve_all(rp, ri, rv, **kw):
    from slim_b200 import Staged, learn_columns

    with Staged(rp, ri, rv) as s:
        res = learn_columns(s, dict(l1r=1.0, l2r=1.0, optTol=1e-7, niters=50, **kw))
        w = res.to_host()
        res.close()
    return w


def test_result_to_model_builds_both_views_on_the_gpu(lib, oracle):
    from slim_b200 import Staged, learn_columns

    rp, ri, rv = st.synth_zipf(3000, 400, 25, seed=11)
    with Staged(rp, ri, rv) as s:
        res = learn_columns(s, dict(l1r=1.0, l2r=1.0, optTol=1e-7, niters=50))
        w = res.to_host()
        stt = C.c_int32(0)
        h = lib.SLIMB200_ResultToModel(res.handle, C.byref(stt))
        assert h and stt.value == st.SLIM_OK
        mv = st.model_views(h)
        res.close()
    assert np.array_equal(mv["colptr"], w["colptr"]) and np.array_equal(mv["colind"], w["colind"])
    assert np.array_equal(mv["colval"].view(np.uint32), w["colval"].view(np.uint32))
    tp, ti, tv = oracle.transpose(400, w["colptr"], w["colind"], w["colval"])  # estimate.c:590 / csr.c:1546-1584
    assert np.array_equal(mv["rowptr"], tp) and np.array_equal(mv["rowind"], ti)
    assert np.array_equal(mv["rowval"].view(np.uint32), tv.view(np.uint32))
    hp = C.c_void_p(h)
    lib.SLIM_FreeModel(C.byref(hp))


def test_allgather_world1_reassembles_shuffled_columns(lib):
    # one rank: the collective degenerates to a copy, the header / scan / placement kernels do all the work
    from slim_b200 import Staged, learn_columns
    from slim_b200.core import ColumnResult

    rp, ri, rv = st.synth_zipf(2500, 300, 20, seed=3)
    cols = np.random.default_rng(0).permutation(300).astype(np.int32)  # solved in a scrambled order
    uid = np.zeros(128, np.uint8)
    assert lib.SLIMB200_CommUniqueId(uid.ctypes.data_as(C.c_void_p)) == st.SLIM_OK
    stt = C.c_int32(0)
    comm = lib.SLIMB200_CommInitRank(0, 1, 0, uid.ctypes.data_as(C.c_void_p), C.byref(stt))
    assert comm, lib.SLIMB200_LastError()
    with Staged(rp, ri, rv) as s:
        res = learn_columns(s, dict(l1r=1.0, l2r=1.0, optTol=1e-7, niters=50), cols=cols)
        w = res.to_host()
        h = lib.SLIMB200_AllGatherColumns(comm, res.handle, cols.ctypes.data_as(C.POINTER(C.c_int32)), 300, C.byref(stt))
        assert h, lib.SLIMB200_LastError()
        full = ColumnResult(lib, h)
        got = full.to_host()
        # a position owned twice / out of range is rejected
        bad = cols.copy()
        bad[0] = bad[1]
        assert not lib.SLIMB200_AllGatherColumns(comm, res.handle, bad.ctypes.data_as(C.POINTER(C.c_int32)), 300,
                                                 C.byref(stt))
        assert stt.value != st.SLIM_OK
        full.close()
        res.close()
    ref = _solve_all(rp, ri, rv)  # all columns in natural order
    for k, j in enumerate(cols):  # the scrambled solve itself ...
        assert np.array_equal(w["colind"][w["colptr"][k]:w["colptr"][k + 1]], ref["colind"][ref["colptr"][j]:ref["colptr"][j + 1]])
    assert np.array_equal(got["colptr"], ref["colptr"])  # ... and its reassembly into column order
    assert np.array_equal(got["colind"], ref["colind"])
    assert np.array_equal(got["colval"].view(np.uint32), ref["colval"].view(np.uint32))
    cp = C.c_void_p(comm)
    lib.SLIMB200_CommFree(C.byref(cp))


def _learn_env(ours, rp, ri, rv, env):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        io, do = st.options(l1r=1.0, l2r=1.0, opttol=1e-7, niters=50)
        h, status = ours.learn(rp, ri, rv, io, do)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    assert h and status == st.SLIM_OK
    mv = st.model_views(h)
    ours.free(h)
    return mv


def _same_model(a, b):
    for k in ("colptr", "colind", "rowptr", "rowind"):
        assert np.array_equal(a[k], b[k]), k
    for k in ("colval", "rowval"):
        assert np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)), k


def test_slim_learn_sharded_path_single_device(lib, ours):
    # SLIM_Learn through stage_multi / learn_multi / NCCL with ONE device == the plain path, bit for bit
    rp, ri, rv = st.synth_zipf(4000, 500, 25, seed=21)
    plain = _learn_env(ours, rp, ri, rv, {"SLIMB200_GPUS": "1"})
    multi = _learn_env(ours, rp, ri, rv, {"SLIMB200_GPUS": "1", "SLIMB200_FORCE_MULTI": "1"})
    _same_model(plain, multi)


def test_slim_learn_on_two_gpus_matches_one(lib, ours):
    if lib.SLIMB200_DeviceCount() < 2:
        pytest.skip("needs two GPUs")
    rp, ri, rv = st.synth_zipf(20000, 2000, 50, seed=4)
    one = _learn_env(ours, rp, ri, rv, {"SLIMB200_GPUS": "1"})
    two = _learn_env(ours, rp, ri, rv, {"SLIMB200_GPUS": "2"})
    _same_model(one, two)
    every = _learn_env(ours, rp, ri, rv, {"SLIMB200_DEVICES": ",".join(str(d) for d in range(lib.SLIMB200_DeviceCount()))})
    _same_model(one, every)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_main(rank, world, port, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from slim_b200 import Staged
        from slim_b200.dist import Communicator, all_gather_columns, sharded_learn

        rp, ri, rv = st.synth_zipf(20000, 1500, 40, seed=9)
        colcnt = np.bincount(ri, minlength=1500)
        comm = Communicator(rank)
        with Staged(rp, ri, rv, device=rank) as s:
            cols = np.sort(np.random.default_rng(1).choice(1500, 1200, replace=False)).astype(np.int32)
            loc, full = sharded_learn(s, dict(l1r=1.0, l2r=1.0, optTol=1e-7, niters=50), comm, cols=cols, colcnt=colcnt)
            got = full.to_host()
            # the torch.distributed restatement of the same exchange (padded, host reassembly)
            dev = torch.device("cuda", rank)
            cnt = torch.empty(max(loc.nsel, 1), dtype=torch.int32, device=dev)
            ind = torch.empty(max(loc.nnz, 1), dtype=torch.int32, device=dev)
            val = torch.empty(max(loc.nnz, 1), dtype=torch.float32, device=dev)
            loc.to_device(cnt, ind, val)
            from slim_b200.dist import shard_columns

            mine = shard_columns(cols, colcnt, rank, world)
            cp, ci, cv = all_gather_columns(mine, cnt[:loc.nsel], ind[:loc.nnz], val[:loc.nnz], len(cols))
            ok = (np.array_equal(got["colptr"], cp) and np.array_equal(got["colind"], ci) and
                  np.array_equal(got["colval"].view(np.uint32), cv.view(np.uint32)))
            full.close()
            loc.close()
        comm.close()
        q.put((rank, bool(ok), int(cp[-1])))
    except Exception as e:
        import traceback

        q.put((rank, False, traceback.format_exc() + repr(e)))
    finally:
        dist.destroy_process_group()


def test_two_rank_allgather_matches_torch_exchange(lib):
    if lib.SLIMB200_DeviceCount() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(120)
    assert all(ok for _, ok, _ in res), res
    assert res[0][2] == res[1][2] > 0
