"""GPU parity at the BENCHMARKED scale (BASELINE.json configs[3]: 1 M users x 100 K items, 100 M nnz).

The small-matrix tests of test_gpu_parity.py never reach the shapes the bench runs: a 100 K-item Gram matrix
(element offsets ~1e10), union nonzero lists of ~95 K entries in cd_gram_batch_kernel<.., 16, 8, 2, 512>, one-target
clusters at full size.  This test builds the C4 matrix exactly as bench.py does (slim_b200.synth.zipf_csr, same
seeds), solves a stratified column sample PLUS the heaviest columns through the C ABI and compares every column
with the oracle run in the engine's visiting order (reference estimate.c:405-505, cd.c:101-142):
|A_j|, sweeps executed, objective (1e-6 relative) and W (1e-5 per nonzero; the heavy columns stop at the 50-sweep
cap far from convergence, so the iterates themselves are compared, not a converged optimum).
"""
import numpy as np
import pytest

import slimtest as st

pytestmark = pytest.mark.gpu

PARAMS = dict(l1r=1.0, l2r=1.0, optTol=1e-7, niters=50)


@pytest.fixture(scope="module")
def c4():
    import torch

    from slim_b200 import Staged, _lib
    from slim_b200.synth import zipf_csr

    L = _lib.load()
    if L.SLIMB200_DeviceCount() < 1:
        pytest.skip("needs a CUDA device")
    rp, ri, rv = zipf_csr(1_000_000, 100_000, 100, device="cuda:0")
    staged = Staged(rp, ri, rv, device=0)
    host = (rp.cpu().numpy(), ri.cpu().numpy(), rv.cpu().numpy())
    del rp, ri, rv
    torch.cuda.empty_cache()
    yield staged, host
    staged.close()


def _compare(staged, host, cols, oracle, w_tol=1e-5):
    from slim_b200 import learn_columns

    res = learn_columns(staged, PARAMS, cols=cols)
    got, gs = res.to_host(), res.stats()
    res.close()
    import os

    ref = oracle.learn(*host, l1r=PARAMS["l1r"], l2r=PARAMS["l2r"], opttol=PARAMS["optTol"], niters=PARAMS["niters"],
                       order=st.ORDER_POPULARITY, nthreads=min(len(cols), os.cpu_count() or 1), cols=cols,
                       want_stats=True)
    rs = ref["stats"]
    assert np.array_equal(gs["nactive"], rs["nactive"])
    assert np.array_equal(gs["active_nnz"], rs["active_nnz"])
    assert np.array_equal(gs["niters"], rs["niters"]), (gs["niters"], rs["niters"])
    rel = np.abs(gs["objval"] - rs["objval"]) / np.maximum(np.abs(rs["objval"]), 1e-300)
    assert rel.max() <= 1e-6, rel.max()
    maxd, flips = st.compare_models(got, ref)
    assert maxd <= w_tol, maxd
    assert all(mag < w_tol for _, _, mag in flips), flips[:5]
    return gs


def test_c4_stratified_columns_match_oracle(c4, oracle):
    from slim_b200.synth import stratified_columns

    staged, host = c4
    colcnt = np.bincount(host[1], minlength=staged.ncols)
    cols = stratified_columns(colcnt, 32, offset=7)
    gs = _compare(staged, host, cols, oracle)
    assert gs["niters"].min() >= 1


def test_c4_heaviest_columns_match_oracle(c4, oracle):
    # the 8 most popular items: every one of them goes through the batched Gram kernel at its full shape
    # (16 CTAs x 512 threads, DMMA gather, ~all 100 K coordinates active, 50 capped sweeps); 4 more columns with
    # 2 000 <= nnz < 30 000 run next to them in the same call: the two lighter ones on one-target clusters of 4 CTAs,
    # the two heavier ones (>= 9 000 nonzeros) in a second, short batch
    staged, host = c4
    colcnt = np.bincount(host[1], minlength=staged.ncols)
    order = np.argsort(-colcnt, kind="stable")
    mid = order[(colcnt[order] < 30000) & (colcnt[order] >= 2000)]
    cols = np.sort(np.concatenate([order[:8], mid[[0, len(mid) // 3, 2 * len(mid) // 3, len(mid) - 1]]])).astype(np.int32)
    gs = _compare(staged, host, cols, oracle)
    assert gs["niters"].max() == 51  # the giants hit the cap (cd.c:140 reports maxniters + 1)
