"""Synthetic user x item matrices of the BASELINE.json shapes (bench / test harness utility).

Every user rates exactly `per_user` DISTINCT items drawn without replacement from a Zipf(alpha)
popularity law (successive sampling: i.i.d. draws from p with already-seen items rejected, which
has the same law as Gumbel-top-k); item ids are then permuted so popularity is not monotone in id;
column ids are sorted inside each row; all ratings are 1.0 (or uniform in {1..5}).  torch is used
only as the array engine (runs on the GPU when there is one)."""
from __future__ import annotations

import numpy as np
import torch


def zipf_csr(nusers: int, nitems: int, per_user: int, *, alpha: float = 1.1, seed: int = 42,
             perm_seed: int = 43, ratings: bool = False, device: str | torch.device = "cpu",
             chunk_users: int = 1 << 16):
    """Returns torch tensors on `device`: rowptr int64[nusers+1], rowind int32[nnz], rowval fp32[nnz]."""
    assert per_user <= nitems
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    p = torch.arange(1, nitems + 1, dtype=torch.float64, device=dev).pow(-alpha)
    cdf = torch.cumsum(p / p.sum(), 0)
    cdf[-1] = 1.0
    gp = torch.Generator(device="cpu")
    gp.manual_seed(perm_seed)
    perm = torch.randperm(nitems, generator=gp).to(dev)
    out = torch.empty((nusers, per_user), dtype=torch.int32, device=dev)
    ndraw = min(max(4 * per_user, per_user + 64), 8 * per_user)
    for s in range(0, nusers, chunk_users):
        e = min(nusers, s + chunk_users)
        b = e - s
        got = torch.full((b, per_user), -1, dtype=torch.int64, device=dev)
        have = torch.zeros(b, dtype=torch.int64, device=dev)
        todo = torch.arange(b, device=dev)
        while todo.numel() > 0:
            nb = todo.numel()
            u = torch.rand((nb, ndraw), generator=g, device=dev, dtype=torch.float64)
            draws = torch.searchsorted(cdf, u.reshape(-1)).clamp_(max=nitems - 1).reshape(nb, ndraw)
            # prepend what the row already has so repeats of earlier picks are rejected too
            prev = got[todo]
            cat = torch.cat([prev, draws], dim=1)
            valid = cat >= 0
            key = torch.where(valid, cat, torch.full_like(cat, nitems) + torch.arange(cat.shape[1], device=dev))
            sv, si = torch.sort(key, dim=1, stable=True)
            dup_sorted = torch.zeros_like(sv, dtype=torch.bool)
            dup_sorted[:, 1:] = sv[:, 1:] == sv[:, :-1]
            dup = torch.zeros_like(dup_sorted)
            dup.scatter_(1, si, dup_sorted)
            keep = valid & ~dup
            rank = torch.cumsum(keep.to(torch.int64), 1) - 1
            sel = keep & (rank < per_user)
            rows, cols = torch.nonzero(sel, as_tuple=True)
            newgot = torch.full((nb, per_user), -1, dtype=torch.int64, device=dev)
            newgot[rows, rank[rows, cols]] = cat[rows, cols]
            got[todo] = newgot
            have_t = sel.sum(1)
            have[todo] = have_t
            todo = todo[have_t < per_user]
        out[s:e] = torch.sort(perm[got], dim=1).values.to(torch.int32)
    rowptr = torch.arange(0, (nusers + 1) * per_user, per_user, dtype=torch.int64, device=dev)
    rowind = out.reshape(-1)
    if ratings:
        rowval = torch.randint(1, 6, (nusers * per_user,), generator=g, device=dev).to(torch.float32)
    else:
        rowval = torch.ones(nusers * per_user, dtype=torch.float32, device=dev)
    return rowptr, rowind, rowval


def stratified_columns(colcnt: np.ndarray, ncols_sel: int, offset: int = 0) -> np.ndarray:
    """Every k-th column in nnz-sorted order (SURVEY.md 8d "stratified column sample"), taken from
    the MIDDLE of each stratum: with proportional allocation the sample mean of the per-column
    cost is an unbiased estimate of the full-matrix mean whatever the sample size, so samples of
    different sizes (GPU step vs CPU step) estimate the same columns/s.  `offset` shifts the comb
    so different steps use disjoint samples."""
    n = len(colcnt)
    ncols_sel = min(ncols_sel, n)
    order = np.argsort(-colcnt.astype(np.int64), kind="stable")
    stride = n / ncols_sel
    pos = (np.floor((np.arange(ncols_sel) + 0.5) * stride).astype(np.int64) + offset) % n
    return np.sort(order[pos]).astype(np.int32)
