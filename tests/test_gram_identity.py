"""CPU check of the identity the Gram-space kernels rest on (slim_b200/csrc/gram.cuh): sequential coordinate
descent in USER space (reference src/libslim/cd.c:101-142, restated by the oracle) and the same sweep carried
out in GRAM space -- <a_i, yhat> = sum_k x_k G[k][i] with G = R^T R, 32-coordinate blocks whose inner
dependence is resolved with the in-block Gram tile, only changing coordinates visited -- produce the same
iterates.  numpy restatement of the kernel's block procedure vs the oracle in the same visiting order."""
import numpy as np
import scipy.sparse as sp

import slimtest as st

EPS = 1e-7  # EPSILON, reference src/libslim/def.h:14


def gram_space_cd(G, cnorm, j, l1r, l2r, opttol, maxniters, colcnt_j, block=32, item_space=False):
    """Column j of W by the procedure of cd_gram_kernel (internal item order = order of G): blocks of `block`
    consecutive ACTIVE coordinates; item_space=True: blocks of `block` consecutive ITEM ids with the inactive ones
    masked out, the blocking of cd_gram_batch_kernel (64 items per block)."""
    aty_all = G[j].copy()
    act = np.nonzero((aty_all > l1r) & (np.arange(len(aty_all)) != j))[0]
    na = len(act)
    x = np.zeros(na)
    maxit = min(50 * colcnt_j, maxniters)
    niters = 1
    if na == 0 or maxit <= 0:
        return act, x, niters
    aty = aty_all[act].astype(np.float32).astype(np.float64)  # gk_fkv_t.key is a float
    den = cnorm[act].astype(np.float64) ** 2 + l2r
    sq = np.diag(G)[act]
    GA = G[np.ix_(act, act)]
    done = False
    t = 0
    while t < maxit and not done:
        dl = 0.0
        if item_space:  # [start, end) positions of the actives inside every block of `block` item ids
            edges = np.searchsorted(act, np.arange(0, G.shape[0] + block, block))
            spans = [(a, b) for a, b in zip(edges[:-1], edges[1:]) if b > a]
        else:
            spans = [(p0, min(na, p0 + block)) for p0 in range(0, na, block)]
        for p0, p1 in spans:
            sl = slice(p0, p1)
            xin = np.where(np.abs(x) > EPS, x, 0.0)  # AddSpVec's EPSILON rule (cd.c:27)
            ipf = GA[sl] @ xin                      # the gather over the nonzero list
            xb = x[sl].copy()
            gbb = GA[sl, sl]
            k = 0
            while True:                              # the chain: only coordinates whose value changes
                in_old = np.where(np.abs(xb) > EPS, xb, 0.0)
                num = aty[sl] - (ipf - in_old * sq[sl])
                nx = np.where(num > l1r, (num - l1r) / den[sl], 0.0)
                want = np.nonzero((np.arange(len(xb)) >= k) & (nx != xb))[0]
                if len(want) == 0:
                    break
                kk = want[0]
                in_new = nx[kk] if abs(nx[kk]) > EPS else 0.0
                d = in_new - in_old[kk]
                dl += (nx[kk] - xb[kk]) ** 2
                xb[kk] = nx[kk]
                if d != 0.0:
                    ipf = ipf + d * gbb[kk]
                k = kk + 1
            x[sl] = xb
        t += 1
        if dl < opttol:
            done = True
    niters = t if done else maxit + 1
    return act, x, niters


import pytest


@pytest.mark.parametrize("blocking", [dict(block=32), dict(block=64, item_space=True)],
                         ids=["position-blocks-32", "item-blocks-64"])
def test_gram_space_sweeps_equal_user_space_sweeps(oracle, ml100k, blocking):
    g = ml100k
    rp, ri, rv = g["trn_rowptr"], g["trn_rowind"], g["trn_rowval"]
    nitems = int(ri.max()) + 1
    R = sp.csr_matrix((rv.astype(np.float64), ri, rp), shape=(len(rp) - 1, nitems))
    cnt = np.bincount(ri, minlength=nitems)
    inv = np.lexsort((np.arange(nitems), -cnt))  # the engine's internal order: descending nnz, ascending id
    Ri = R[:, inv]
    G = (Ri.T @ Ri).toarray()
    m = oracle.setup(rp, ri, rv)
    cnorm = oracle.csc_arrays(m)["cnorms"][inv]
    oracle.free_csc(m)
    cols = np.array([0, 7, 50, 181, 300, 655, 1000, 1500], np.int32)
    for kw in (dict(l1r=1.0, l2r=1.0, opttol=1e-7, niters=50), dict(l1r=0.5, l2r=2.0, opttol=1e-12, niters=400)):
        ref = oracle.learn(rp, ri, rv, cols=cols, nthreads=1, want_stats=True, order=st.ORDER_POPULARITY, **kw)
        rank = np.empty(nitems, np.int64)
        rank[inv] = np.arange(nitems)
        for q, j in enumerate(cols):
            act, x, niters = gram_space_cd(G, cnorm, int(rank[j]), kw["l1r"], kw["l2r"], kw["opttol"], kw["niters"],
                                           int(cnt[j]), **blocking)
            keep = np.abs(x) > EPS
            got = dict(zip(inv[act[keep]].tolist(), x[keep].astype(np.float32).tolist()))
            a, b = ref["colptr"][q], ref["colptr"][q + 1]
            want = dict(zip(ref["colind"][a:b].tolist(), ref["colval"][a:b].tolist()))
            assert ref["stats"]["niters"][q] == niters, (j, niters)
            assert ref["stats"]["nactive"][q] == len(act)
            assert set(got) == set(want), j
            assert max((abs(got[k] - want[k]) for k in got), default=0.0) <= 1e-7, j
