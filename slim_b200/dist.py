"""Column-sharded multi-GPU learning: one process per GPU (torch.distributed), R replicated, the
target item columns dealt across ranks, and ONE exchange at the end that assembles W on every rank.

The reference's only parallelism is `#pragma omp for schedule(dynamic,32)` over target columns
(src/libslim/estimate.c:402-403) -- the columns are independent problems, so there is no data-path
collective during the solve; the all-gather below is the NCCL (or gloo, in CPU tests) counterpart
of SaveModel's concatenation of per-column lists (estimate.c:570-588).
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


def sharded_learn(staged, params, cols: np.ndarray | None = None, colcnt: np.ndarray | None = None,
                  group=None, gather: bool = True):
    """Solve `cols` (None: all columns) across the ranks of `group`; every rank holds a replica of R
    (`staged`).  Returns (local ColumnResult, (colptr, colind, colval) of ALL columns or None)."""
    from .core import learn_columns

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if cols is None:
        cols = np.arange(staged.ncols, dtype=np.int32)
    cols = np.ascontiguousarray(cols, dtype=np.int32)
    mine = shard_columns(cols, colcnt, rank, world)
    res = learn_columns(staged, params, cols=cols[mine])
    if not gather:
        return res, None
    dev = torch.device("cuda", staged.device)
    counts = torch.empty(max(res.nsel, 1), dtype=torch.int32, device=dev)
    ind = torch.empty(max(res.nnz, 1), dtype=torch.int32, device=dev)
    val = torch.empty(max(res.nnz, 1), dtype=torch.float32, device=dev)
    res.to_device(counts, ind, val)
    full = all_gather_columns(mine, counts[:res.nsel], ind[:res.nnz], val[:res.nnz], len(cols), group)
    return res, full
