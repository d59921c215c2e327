"""Column-sharded multi-GPU learning with one process per GPU: R replicated, the target item columns
dealt across ranks, and ONE exchange at the end that assembles W on every rank.

The reference's only parallelism is `#pragma omp for schedule(dynamic,32)` over target columns
(src/libslim/estimate.c:402-403) -- the columns are independent problems, so there is no data-path
collective during the solve; the final all-gather is the counterpart of SaveModel's concatenation of
per-column lists (estimate.c:570-588).

The exchange itself lives in the LIBRARY (slim_b200/csrc/gather.cuh behind SLIMB200_AllGatherColumns):
NCCL over NVLink, variable-length, reassembled on the device.  `Communicator` only carries the 128-byte
NCCL id from rank 0 to the other ranks through torch.distributed (plumbing).  `all_gather_columns` below
is the torch.distributed restatement of the same exchange: it runs with gloo on CPU tensors, which is how
the host-side logic (sharding, reassembly order) is tested without GPUs, and it is what the multi-rank GPU
test compares the library path with.  (Inside ONE process SLIM_Learn shards over the visible GPUs by
itself; see include/slim_b200.h.)
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_columns(cols: np.ndarray, colcnt: np.ndarray | None, rank: int, world: int) -> np.ndarray:
    """Columns owned by `rank`: deal the columns round-robin in descending-nnz order (cheap proxy of
    the per-column cost) so every rank gets the same head/tail mix.  Returns indices
    into `cols` (ascending)."""
    n = len(cols)
    if colcnt is None:
        order = np.arange(n)
    else:
        order = np.argsort(-colcnt[cols].astype(np.int64), kind="stable")
    return np.sort(order[rank::world])


def all_gather_columns(local_idx: np.ndarray, counts: torch.Tensor, colind: torch.Tensor,
                       colval: torch.Tensor, ncols_total: int, group=None):
    """Exchange the solved column segments of every rank.

    local_idx : positions (into the global column list, length ncols_total) of this rank's columns
    counts    : int32[nloc]  nnz per local column     (tensor on the collective's device)
    colind    : int32[nnz_loc], colval: float32[nnz_loc]  concatenated local segments
    Returns (colptr int64[ncols_total+1], colind int32[nnz], colval float32[nnz]) as numpy arrays in
    global column order, identical on every rank.
    """
    world = dist.get_world_size(group)
    dev = counts.device
    nloc = int(counts.numel())
    nnz_loc = int(colind.numel())
    meta = torch.tensor([nloc, nnz_loc], dtype=torch.int64, device=dev)
    metas = torch.empty(world * 2, dtype=torch.int64, device=dev)  # flat: gloo wants 1-D outputs
    dist.all_gather_into_tensor(metas, meta, group=group)
    metas_h = metas.cpu().numpy().reshape(world, 2)
    max_loc, max_nnz = int(metas_h[:, 0].max()), int(metas_h[:, 1].max())

    # header: [column position, count] per local column, padded to the largest shard
    head = torch.full((max(max_loc, 1), 2), -1, dtype=torch.int32, device=dev)
    if nloc:
        head[:nloc, 0] = torch.as_tensor(local_idx, dtype=torch.int32).to(dev)
        head[:nloc, 1] = counts
    heads = torch.empty(world * head.numel(), dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(heads, head.reshape(-1), group=group)

    # payload: (index, value) pairs packed as two int32 lanes, padded to the largest shard
    pay = torch.zeros((max(max_nnz, 1), 2), dtype=torch.int32, device=dev)
    if nnz_loc:
        pay[:nnz_loc, 0] = colind
        pay[:nnz_loc, 1] = colval.view(torch.int32)
    pays = torch.empty(world * pay.numel(), dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(pays, pay.reshape(-1), group=group)

    heads_h = heads.cpu().numpy().reshape(world, -1, 2)
    pays_h = pays.cpu().numpy().reshape(world, -1, 2)
    cnt = np.zeros(ncols_total, dtype=np.int64)
    for r in range(world):
        n = int(metas_h[r, 0])
        cnt[heads_h[r, :n, 0]] = heads_h[r, :n, 1]
    colptr = np.zeros(ncols_total + 1, dtype=np.int64)
    np.cumsum(cnt, out=colptr[1:])
    out_ind = np.empty(int(colptr[-1]), dtype=np.int32)
    out_val = np.empty(int(colptr[-1]), dtype=np.float32)
    for r in range(world):
        n = int(metas_h[r, 0])
        src = 0
        seg_ind = pays_h[r, :, 0]
        seg_val = pays_h[r, :, 1].view(np.float32)
        for k in range(n):
            j, c = int(heads_h[r, k, 0]), int(heads_h[r, k, 1])
            out_ind[colptr[j]:colptr[j] + c] = seg_ind[src:src + c]
            out_val[colptr[j]:colptr[j] + c] = seg_val[src:src + c]
            src += c
    return colptr, out_ind, out_val


class Communicator:
    """NCCL communicator owned by libslim.so (SLIMB200_CommInitRank), one per rank / GPU.  The 128-byte id
    created on rank 0 travels to the other ranks through torch.distributed (any backend)."""

    def __init__(self, device: int, group=None):
        import ctypes as C

        from . import _lib

        self._lib = _lib.load()
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        uid = np.zeros(128, dtype=np.uint8)
        if self.rank == 0:
            rc = self._lib.SLIMB200_CommUniqueId(uid.ctypes.data_as(C.c_void_p))
            if rc != 1:
                raise RuntimeError("SLIMB200_CommUniqueId failed: " + (self._lib.SLIMB200_LastError() or b"").decode())
        backend = dist.get_backend(group)
        t = torch.from_numpy(uid)
        if backend == "nccl":
            t = t.to(torch.device("cuda", device))
        dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        uid = t.cpu().numpy()
        st = C.c_int32(0)
        self.handle = self._lib.SLIMB200_CommInitRank(device, self.world, self.rank, uid.ctypes.data_as(C.c_void_p),
                                                      C.byref(st))
        if not self.handle:
            raise RuntimeError("SLIMB200_CommInitRank failed (%d): %s" %
                               (st.value, (self._lib.SLIMB200_LastError() or b"").decode()))

    def all_gather_columns(self, local, positions: np.ndarray, ncols_total: int):
        """Collective: `local` (ColumnResult) holds this rank's columns, positions[k] = index of local column k
        in the global list.  Returns a ColumnResult with ALL columns, resident on this rank's GPU."""
        import ctypes as C

        from .core import ColumnResult

        pos = np.ascontiguousarray(positions, dtype=np.int32)
        st = C.c_int32(0)
        h = self._lib.SLIMB200_AllGatherColumns(self.handle, local.handle, pos.ctypes.data_as(C.POINTER(C.c_int32)),
                                                int(ncols_total), C.byref(st))
        if not h:
            raise RuntimeError("SLIMB200_AllGatherColumns failed (%d): %s" %
                               (st.value, (self._lib.SLIMB200_LastError() or b"").decode()))
        return ColumnResult(self._lib, h)

    def close(self):
        import ctypes as C

        if getattr(self, "handle", None):
            h = C.c_void_p(self.handle)
            self._lib.SLIMB200_CommFree(C.byref(h))
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def sharded_learn(staged, params, comm: Communicator, cols: np.ndarray | None = None,
                  colcnt: np.ndarray | None = None, gather: bool = True):
    """Solve `cols` (None: all columns) across the ranks of `comm`; every rank holds a replica of R
    (`staged`).  Returns (local ColumnResult, ColumnResult with ALL columns on the device or None)."""
    from .core import learn_columns

    if cols is None:
        cols = np.arange(staged.ncols, dtype=np.int32)
    cols = np.ascontiguousarray(cols, dtype=np.int32)
    mine = shard_columns(cols, colcnt, comm.rank, comm.world)
    res = learn_columns(staged, params, cols=cols[mine])
    if not gather:
        return res, None
    return res, comm.all_gather_columns(res, mine, len(cols))
