"""GPU parity tests: the CUDA engine, called through the C ABI of libslim.so, against the oracle
(oracle/slim_oracle.c) on the same inputs and against the committed reference goldens.

Tolerances: at the converged setting |W_gpu - W_ref| <= 2e-6 per nonzero
over the union support, support may differ only where |w| < 2e-6, top-10 lists identical, HR/ARHR
equal to 4 decimals.  Staging (CSR -> CSC, norms) is integer / ordered-float work: bit-exact.
"""
import ctypes as C
import os
from pathlib import Path

import numpy as np
import pytest

import slimtest as st

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
CONV = dict(opttol=1e-14, niters=100000)
TOL = 2e-6


@pytest.fixture(scope="module")
def lib():
    from slim_b200 import _lib

    L = _lib.load()  # raises if slim_b200/lib/libslim.so has not been built: no fallback
    assert L.SLIMB200_DeviceCount() > 0, "these tests need a CUDA device"
    return L


@pytest.fixture(scope="module")
def ours():
    return st.SlimLib(ROOT / "slim_b200" / "lib" / "libslim.so")


def _golden_model(g, tag):
    return dict(colptr=g[f"W_{tag}_colptr"], colind=g[f"W_{tag}_colind"], colval=g[f"W_{tag}_colval"])


def _learn(ours, rp, ri, rv, imodel=None, **kw):
    io, do = st.options(**kw)
    h, status = ours.learn(rp, ri, rv, io, do, imodel)
    assert h and status == st.SLIM_OK
    return h


def _check_close(w, ref, tol=TOL):
    maxd, flips = st.compare_models(w, ref)
    assert maxd <= tol, maxd
    assert all(mag < tol for _, _, mag in flips), flips[:5]


@pytest.mark.parametrize("name", ["ml100k", "automotive"])
@pytest.mark.parametrize("binary", [False, True])
def test_staging_is_bit_exact(lib, oracle, name, binary):
    from slim_b200 import Staged

    g = st.load_golden(name)
    rv = None if binary else g["trn_rowval"]
    m = oracle.setup(g["trn_rowptr"], g["trn_rowind"], rv)
    ref = oracle.csc_arrays(m)
    oracle.free_csc(m)
    with Staged(g["trn_rowptr"], g["trn_rowind"], rv) as s:
        assert (s.nrows, s.ncols, s.nnz) == (ref["nrows"], ref["ncols"], len(ref["colind"]))
        got = s.csc()
    assert np.array_equal(got["colptr"], ref["colptr"])
    assert np.array_equal(got["colind"], ref["colind"])
    if not binary:
        assert np.array_equal(got["colval"].view(np.uint32), ref["colval"].view(np.uint32))
    assert np.array_equal(got["cnorms"].view(np.uint32), ref["cnorms"].view(np.uint32))


def test_staging_float_norm_rounding(lib, oracle):
    # non-integer ratings: the float accumulation order of gk_fdot matters in the last bit
    rp, ri, _ = st.synth_zipf(4000, 300, 20, seed=5)
    rv = np.random.default_rng(9).random(len(ri)).astype(np.float32) * 5 + 0.01
    from slim_b200 import Staged

    m = oracle.setup(rp, ri, rv)
    ref = oracle.csc_arrays(m)
    oracle.free_csc(m)
    with Staged(rp, ri, rv) as s:
        got = s.csc()
    assert np.array_equal(got["colind"], ref["colind"])
    assert np.array_equal(got["cnorms"].view(np.uint32), ref["cnorms"].view(np.uint32))


@pytest.mark.parametrize("ratings", [False, True])
def test_window_gram_blocks(lib, ratings):
    # staged Gram blocks of 32 consecutive item columns == the diagonal blocks of R^T R (exact)
    import scipy.sparse as sp

    from slim_b200 import Staged

    rp, ri, rv = st.synth_zipf(3000, 200, 25, seed=8, ratings=ratings)
    R = sp.csr_matrix((rv.astype(np.float64), ri, rp), shape=(3000, 200))
    G = (R.T @ R).toarray()
    with Staged(rp, ri, rv) as s:
        got = s.window_gram()
        rank = s.item_order()
    cnt = np.bincount(ri, minlength=200)
    inv = np.lexsort((np.arange(200), -cnt))  # descending nnz, ties by ascending id
    assert np.array_equal(rank[inv], np.arange(200))
    G = G[np.ix_(inv, inv)]  # Gram matrix in the engine's internal item order
    assert got.shape == (7, 32, 32)
    for w in range(7):
        n = min(32, 200 - 32 * w)
        ref = G[32 * w:32 * w + n, 32 * w:32 * w + n].copy()
        np.fill_diagonal(ref, 0.0)
        assert np.array_equal(got[w, :n, :n], ref), w
        assert not got[w, n:, :].any() and not got[w, :, n:].any()


@pytest.mark.parametrize("name", ["ml100k", "automotive"])
def test_learn_matches_reference_golden(lib, ours, name):
    g = st.load_golden(name)
    h = _learn(ours, g["trn_rowptr"], g["trn_rowind"], g["trn_rowval"], **CONV)
    mv = st.model_views(h)
    ref = _golden_model(g, "conv")
    assert mv["nrows"] == mv["ncols"] == len(ref["colptr"]) - 1
    _check_close(mv, ref)
    assert abs(len(mv["colind"]) - len(ref["colind"])) <= 2
    # both views are consistent, ascending, no diagonal, positive weights
    for j in range(mv["ncols"]):
        seg = mv["colind"][mv["colptr"][j]:mv["colptr"][j + 1]]
        assert (np.diff(seg) > 0).all() and j not in seg
    assert (mv["colval"] > 0).all()
    ids, _ = ours.topn_all(h, g["trn_rowptr"], g["trn_rowind"], g["trn_rowval"], 10)
    bad = np.nonzero((ids != g["top10_ids"]).any(axis=1))[0]
    if name == "ml100k":  # user 277: 10th/11th scores 5.7e-7 apart
        assert set(bad.tolist()) <= {277}, bad
    else:
        assert len(bad) == 0, bad
    ev = st.evaluate(ids, (g["trn_rowptr"], g["trn_rowind"]), (g["tst_rowptr"], g["tst_rowind"]),
                     mv["ncols"], g["fmarker"])
    got = np.array([ev["hr"], ev["hr_head"], ev["hr_tail"], ev["arhr"]])
    assert np.array_equal(np.round(got, 4), np.round(g["metrics"], 4)), (got, g["metrics"])
    ours.free(h)


@pytest.mark.parametrize("name", ["ml100k", "automotive"])
def test_default_setting_matches_oracle_same_order(lib, ours, oracle, name):
    # library defaults (optTol 1e-7, 10000 sweeps): compare with the oracle run in the SAME fixed
    # ascending order; only fp64 summation order differs
    g = st.load_golden(name)
    h = _learn(ours, g["trn_rowptr"], g["trn_rowind"], g["trn_rowval"])
    mv = st.model_views(h)
    w = oracle.learn(g["trn_rowptr"], g["trn_rowind"], g["trn_rowval"], nthreads=8, order=st.ORDER_POPULARITY)
    _check_close(mv, w, tol=1e-6)
    # and against the reference's own (shuffled-order) result at its self-noise level
    maxd, _ = st.compare_models(mv, _golden_model(g, "default"))
    assert maxd <= 5e-3
    ours.free(h)


@pytest.mark.parametrize("nt", ["32", "128", "512"])
@pytest.mark.parametrize("yglobal", ["0", "1"])
@pytest.mark.parametrize("ratings,null_vals", [(False, False), (True, False), (False, True)])
def test_kernel_variants_small(lib, ours, oracle, monkeypatch, nt, yglobal, ratings, null_vals):
    monkeypatch.setenv("SLIMB200_GRAM", "0")  # the user-space team kernel
    monkeypatch.setenv("SLIMB200_NT", nt)
    monkeypatch.setenv("SLIMB200_YHAT_GLOBAL", yglobal)
    rp, ri, rv = st.synth_zipf(700, 260, 24, seed=13, ratings=ratings)
    if null_vals:
        rv = None
    kw = dict(l1r=0.7, l2r=1.5, **CONV)
    h = _learn(ours, rp, ri, rv, **kw)
    w = oracle.learn(rp, ri, rv, nthreads=8, **kw, order=st.ORDER_POPULARITY)
    _check_close(st.model_views(h), w)
    ours.free(h)


@pytest.mark.parametrize("cs", ["1", "2", "4", "8", "16"])
@pytest.mark.parametrize("window", ["0", "1"])
@pytest.mark.parametrize("ratings,null_vals", [(False, False), (True, False), (False, True)])
def test_cluster_kernel_variants(lib, ours, oracle, monkeypatch, cs, window, ratings, null_vals):
    # the thread-block-cluster kernel (used when yhat does not fit shared memory), forced on a small R;
    # window=1: 32-coordinate exact block updates through the staged Gram blocks, window=0: one
    # coordinate per cluster barrier
    monkeypatch.setenv("SLIMB200_GRAM", "0")  # the user-space kernels
    monkeypatch.setenv("SLIMB200_CLUSTER", cs)
    monkeypatch.setenv("SLIMB200_WINDOW", window)
    rp, ri, rv = st.synth_zipf(1000, 260, 24, seed=29, ratings=ratings)
    if null_vals:
        rv = None
    kw = dict(l1r=0.7, l2r=1.5, **CONV)
    h = _learn(ours, rp, ri, rv, **kw)
    w = oracle.learn(rp, ri, rv, nthreads=8, **kw, order=st.ORDER_POPULARITY)
    _check_close(st.model_views(h), w)
    ours.free(h)


@pytest.mark.parametrize("cs", ["0", "16"])
def test_unit_values_take_the_index_only_path(lib, ours, monkeypatch, cs):
    # all-ones ratings: dropping the value stream on the device must not change a single bit
    monkeypatch.setenv("SLIMB200_GRAM", "0")  # the user-space kernels
    monkeypatch.setenv("SLIMB200_CLUSTER", cs)
    rp, ri, rv = st.synth_zipf(3000, 300, 30, seed=31)
    h0 = _learn(ours, rp, ri, rv, niters=50)
    monkeypatch.setenv("SLIMB200_KEEP_VALUES", "1")
    h1 = _learn(ours, rp, ri, rv, niters=50)
    a, b = st.model_views(h0), st.model_views(h1)
    assert np.array_equal(a["colind"], b["colind"])
    assert np.array_equal(a["colval"].view(np.uint32), b["colval"].view(np.uint32))
    ours.free(h0)
    ours.free(h1)


@pytest.mark.parametrize("cs,window", [("0", "1"), ("8", "0"), ("8", "1"), ("16", "1"), ("2", "1")])
def test_long_columns_and_iteration_cap(lib, ours, oracle, monkeypatch, cs, window):
    monkeypatch.setenv("SLIMB200_GRAM", "0")  # the user-space kernels
    monkeypatch.setenv("SLIMB200_CLUSTER", cs)
    monkeypatch.setenv("SLIMB200_WINDOW", window)
    # dense head columns (nnz ~ nusers) exercise the multi-chunk path; niters=50 caps head targets
    rp, ri, rv = st.synth_zipf(6000, 400, 40, seed=21)
    from slim_b200 import Staged, learn_columns

    cols = np.arange(0, 400, 7, dtype=np.int32)
    with Staged(rp, ri, rv) as s:
        r = learn_columns(s, dict(niters=50), cols=cols)
        got, stats = r.to_host(), r.stats()
    ref = oracle.learn(rp, ri, rv, niters=50, cols=cols, nthreads=8, want_stats=True, order=st.ORDER_POPULARITY)
    assert np.array_equal(stats["nactive"], ref["stats"]["nactive"])
    assert np.array_equal(stats["active_nnz"], ref["stats"]["active_nnz"])
    assert np.array_equal(stats["expand_nnz"], ref["stats"]["expand_nnz"])
    assert np.array_equal(stats["niters"], ref["stats"]["niters"])
    _check_close(got, ref, tol=1e-6)
    assert np.allclose(stats["objval"], ref["stats"]["objval"], rtol=1e-9, atol=1e-9)
    assert np.allclose(stats["rnorm"], ref["stats"]["rnorm"], rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("cs", ["0", "16"])
def test_warm_start(lib, ours, oracle, monkeypatch, cs):
    monkeypatch.setenv("SLIMB200_GRAM", "0")  # the user-space kernels
    monkeypatch.setenv("SLIMB200_CLUSTER", cs)
    rp, ri, rv = st.synth_zipf(900, 200, 20, seed=17, ratings=True)
    h0 = _learn(ours, rp, ri, rv, l1r=3.0, l2r=1.0, niters=30)
    m0 = st.model_views(h0)
    h1 = _learn(ours, rp, ri, rv, imodel=h0, l1r=1.0, l2r=1.0, niters=5)
    w1 = oracle.learn(rp, ri, rv, l1r=1.0, l2r=1.0, niters=5, nthreads=4,
                      imodel=(m0["ncols"], m0["colptr"], m0["colind"], m0["colval"]), order=st.ORDER_POPULARITY)
    _check_close(st.model_views(h1), w1, tol=1e-6)
    # 5 sweeps from a warm start differ from 5 cold sweeps: the warm start was really used
    hc = _learn(ours, rp, ri, rv, l1r=1.0, l2r=1.0, niters=5)
    maxd, _ = st.compare_models(st.model_views(hc), w1)
    assert maxd > 1e-4
    for h in (h0, h1, hc):
        ours.free(h)


def test_window_sweep_large_and_small_columns(lib, ours, oracle, monkeypatch):
    # 40K users: head columns exceed the one-warp threshold (4096 entries per CTA range) with small
    # clusters, tail columns stay below it -- both gather paths of the window sweep in one window
    from slim_b200 import Staged, learn_columns

    rp, ri, rv = st.synth_zipf(40000, 600, 30, seed=77)
    cols = np.arange(0, 600, 13, dtype=np.int32)
    ref = oracle.learn(rp, ri, rv, niters=30, cols=cols, nthreads=8, want_stats=True, order=st.ORDER_POPULARITY)
    monkeypatch.setenv("SLIMB200_GRAM", "0")  # the user-space cluster kernel
    for cs in ("1", "4", "16"):
        monkeypatch.setenv("SLIMB200_CLUSTER", cs)
        with Staged(rp, ri, rv) as s:
            r = learn_columns(s, dict(niters=30), cols=cols)
            got, stats = r.to_host(), r.stats()
        assert np.array_equal(stats["niters"], ref["stats"]["niters"])
        assert np.array_equal(stats["nactive"], ref["stats"]["nactive"])
        _check_close(got, ref, tol=1e-6)
        assert np.allclose(stats["objval"], ref["stats"]["objval"], rtol=1e-9)


# ---- the Gram-space solver (slim_b200/csrc/gram.cuh): the default whenever G = R^T R fits in HBM ----------

@pytest.mark.parametrize("width", ["0", "2", "4"])
def test_gram_matrix_is_exact(lib, monkeypatch, width):
    # G staged by gram_build_kernel == R^T R in the engine's internal item order, bit for bit (integer ratings).
    # The packed layout stores a column in 8, 16 or 32 bits depending on its bound rmax * csq_i (gram.cuh);
    # width "0" is that natural layout, "2" forbids the 8-bit range, "4" stores everything in 32 bits.
    import scipy.sparse as sp

    from slim_b200 import Staged

    monkeypatch.setenv("SLIMB200_GRAM_WIDTH", width)
    for ratings, nu, ni, per in ((False, 3000, 200, 25), (True, 3000, 200, 25), (True, 9000, 1200, 12)):
        rp, ri, rv = st.synth_zipf(nu, ni, per, seed=8, ratings=ratings)
        R = sp.csr_matrix((rv.astype(np.float64), ri, rp), shape=(nu, ni))
        G = (R.T @ R).toarray()
        with Staged(rp, ri, rv) as s:
            got = s.gram()
            rank = s.item_order()
            nbytes, h32, h16 = s.gram_layout()
        assert got is not None and got.dtype == np.float32
        inv = np.argsort(rank)
        assert np.array_equal(got.astype(np.float64), G[np.ix_(inv, inv)])
        ld = (ni + 127) // 128 * 128
        assert 0 <= h32 <= h16 <= ld and h32 % 64 == 0 and h16 % 64 == 0
        assert nbytes == ni * (4 * h32 + 2 * (h16 - h32) + (ld - h16)) + 16
        if width == "4":
            assert h32 == ld
        elif width == "2":
            assert h16 == ld
        else:  # the bound that picks a column's width: rmax * (sum of squares of the column)
            csq = np.asarray(R.multiply(R).sum(axis=0)).ravel()[inv] * rv.max()
            assert all(csq[i] <= 65535 for i in range(h32, ni)) and all(csq[i] <= 255 for i in range(h16, ni))
            assert h32 == 0 or csq[h32 - 64:h32].max() > 65535  # ... and the ranges are as large as they can be
            assert h16 == h32 or csq[h16 - 64:h16].max() > 255
    if width == "0":
        assert h32 > 0 and h16 > h32 and h16 < ld  # the last matrix has all three ranges


@pytest.mark.parametrize("width", ["0", "2", "4"])
@pytest.mark.parametrize("route", ["single", "cluster", "batch"])
def test_gram_packed_widths_all_kernels(lib, oracle, monkeypatch, width, route):
    # every Gram-space kernel reads the packed layout through the same accessors: all three element widths, per kernel
    monkeypatch.setenv("SLIMB200_GRAM_WIDTH", width)
    monkeypatch.setenv("SLIMB200_GRAM_CS", "4")
    monkeypatch.setenv("SLIMB200_GRAM_HEAVY", "0" if route == "cluster" else "1000000000")
    monkeypatch.setenv("SLIMB200_GRAM_BATCH", "0" if route == "batch" else "1000000000")
    from slim_b200 import Staged, learn_columns

    rp, ri, rv = st.synth_zipf(9000, 1200, 12, seed=8, ratings=True)  # 32-, 16- and 8-bit columns (see above)
    cols = np.arange(0, 1200, 5, dtype=np.int32)
    with Staged(rp, ri, rv) as s:
        r = learn_columns(s, dict(niters=40), cols=cols)
        got, stats = r.to_host(), r.stats()
    ref = oracle.learn(rp, ri, rv, niters=40, cols=cols, nthreads=8, want_stats=True, order=st.ORDER_POPULARITY)
    assert np.array_equal(stats["nactive"], ref["stats"]["nactive"])
    assert np.array_equal(stats["niters"], ref["stats"]["niters"])
    _check_close(got, ref, tol=1e-6)
    assert np.allclose(stats["objval"], ref["stats"]["objval"], rtol=1e-9)


@pytest.mark.parametrize("cs", ["1", "2", "4", "8", "16"])
@pytest.mark.parametrize("heavy", ["0", "30", "1000000"])  # all targets on clusters / mixed / all on single CTAs
@pytest.mark.parametrize("ratings,null_vals", [(False, False), (True, False), (False, True)])
def test_gram_kernel_variants(lib, ours, oracle, monkeypatch, cs, heavy, ratings, null_vals):
    monkeypatch.setenv("SLIMB200_GRAM_CS", cs)
    monkeypatch.setenv("SLIMB200_GRAM_HEAVY", heavy)
    monkeypatch.setenv("SLIMB200_GRAM_BATCH", "1000000000")  # not the batched kernel (tested below)
    rp, ri, rv = st.synth_zipf(1000, 260, 24, seed=29, ratings=ratings)
    if null_vals:
        rv = None
    kw = dict(l1r=0.7, l2r=1.5, **CONV)
    h = _learn(ours, rp, ri, rv, **kw)
    w = oracle.learn(rp, ri, rv, nthreads=8, **kw, order=st.ORDER_POPULARITY)
    _check_close(st.model_views(h), w)
    ours.free(h)


@pytest.mark.parametrize("cs,heavy", [("1", "0"), ("8", "0"), ("16", "200")])
@pytest.mark.parametrize("f64", ["0", "1"])
def test_gram_real_valued_ratings(lib, ours, oracle, monkeypatch, cs, heavy, f64):
    # non-integer ratings: fp32 sums would not be exact, G is staged in fp64 (also forced on integer data)
    monkeypatch.setenv("SLIMB200_GRAM_CS", cs)
    monkeypatch.setenv("SLIMB200_GRAM_HEAVY", heavy)
    monkeypatch.setenv("SLIMB200_GRAM_BATCH", "1000000000")  # not the batched kernel (tested below)
    monkeypatch.setenv("SLIMB200_GRAM_F64", f64)
    rp, ri, rv = st.synth_zipf(1200, 240, 20, seed=19, ratings=True)
    if f64 == "0":
        rv = (rv * np.random.default_rng(3).uniform(0.2, 1.0, len(rv))).astype(np.float32)
    kw = dict(l1r=0.4, l2r=1.0, **CONV)
    h = _learn(ours, rp, ri, rv, **kw)
    w = oracle.learn(rp, ri, rv, nthreads=8, **kw, order=st.ORDER_POPULARITY)
    _check_close(st.model_views(h), w)
    ours.free(h)


@pytest.mark.parametrize("cs,heavy", [("1", "0"), ("8", "0"), ("16", "0"), ("4", "300")])
def test_gram_long_columns_and_iteration_cap(lib, ours, oracle, monkeypatch, cs, heavy):
    # same case as the user-space kernels: capped head targets must follow the oracle sweep by sweep
    monkeypatch.setenv("SLIMB200_GRAM_CS", cs)
    monkeypatch.setenv("SLIMB200_GRAM_HEAVY", heavy)
    monkeypatch.setenv("SLIMB200_GRAM_BATCH", "1000000000")  # not the batched kernel (tested below)
    rp, ri, rv = st.synth_zipf(6000, 400, 40, seed=21)
    from slim_b200 import Staged, learn_columns

    cols = np.arange(0, 400, 7, dtype=np.int32)
    with Staged(rp, ri, rv) as s:
        assert s.gram() is not None
        r = learn_columns(s, dict(niters=50), cols=cols)
        got, stats = r.to_host(), r.stats()
    ref = oracle.learn(rp, ri, rv, niters=50, cols=cols, nthreads=8, want_stats=True, order=st.ORDER_POPULARITY)
    assert np.array_equal(stats["nactive"], ref["stats"]["nactive"])
    assert np.array_equal(stats["active_nnz"], ref["stats"]["active_nnz"])
    assert np.array_equal(stats["expand_nnz"], ref["stats"]["expand_nnz"])
    assert np.array_equal(stats["niters"], ref["stats"]["niters"])
    _check_close(got, ref, tol=1e-6)
    assert np.allclose(stats["objval"], ref["stats"]["objval"], rtol=1e-9, atol=1e-9)
    assert np.allclose(stats["rnorm"], ref["stats"]["rnorm"], rtol=1e-9, atol=1e-7)


@pytest.mark.parametrize("cs,heavy", [("1", "0"), ("8", "0")])
def test_gram_warm_start(lib, ours, oracle, monkeypatch, cs, heavy):
    monkeypatch.setenv("SLIMB200_GRAM_CS", cs)
    monkeypatch.setenv("SLIMB200_GRAM_HEAVY", heavy)
    monkeypatch.setenv("SLIMB200_GRAM_BATCH", "1000000000")  # not the batched kernel (tested below)
    rp, ri, rv = st.synth_zipf(900, 200, 20, seed=17, ratings=True)
    h0 = _learn(ours, rp, ri, rv, l1r=3.0, l2r=1.0, niters=30)
    m0 = st.model_views(h0)
    h1 = _learn(ours, rp, ri, rv, imodel=h0, l1r=1.0, l2r=1.0, niters=5)
    w1 = oracle.learn(rp, ri, rv, l1r=1.0, l2r=1.0, niters=5, nthreads=4,
                      imodel=(m0["ncols"], m0["colptr"], m0["colind"], m0["colval"]), order=st.ORDER_POPULARITY)
    _check_close(st.model_views(h1), w1, tol=1e-6)
    for h in (h0, h1):
        ours.free(h)


def test_gram_and_user_space_kernels_agree(lib, monkeypatch):
    # same visiting order, same update rule: the two formulations differ by fp64 rounding only
    from slim_b200 import Staged, learn_columns

    rp, ri, rv = st.synth_zipf(40000, 600, 30, seed=77)
    cols = np.arange(0, 600, 13, dtype=np.int32)
    out = {}
    for gram in ("1", "0"):
        monkeypatch.setenv("SLIMB200_GRAM", gram)
        with Staged(rp, ri, rv) as s:
            r = learn_columns(s, dict(niters=30), cols=cols)
            out[gram] = (r.to_host(), r.stats())
    assert np.array_equal(out["1"][1]["niters"], out["0"][1]["niters"])
    assert np.array_equal(out["1"][1]["nactive"], out["0"][1]["nactive"])
    assert np.array_equal(out["1"][1]["expand_nnz"], out["0"][1]["expand_nnz"])
    _check_close(out["1"][0], out["0"][0], tol=1e-6)
    assert np.allclose(out["1"][1]["objval"], out["0"][1]["objval"], rtol=1e-9)


def test_gram_medium_matrix_cluster_class(lib, oracle, monkeypatch):
    # 2 000 items: element offsets beyond one panel range, 8-bit and 16-bit columns side by side in one block
    monkeypatch.setenv("SLIMB200_GRAM_HEAVY", "0")
    monkeypatch.setenv("SLIMB200_GRAM_BATCH", "1000000000")
    from slim_b200 import Staged, learn_columns

    rp, ri, rv = st.synth_zipf(20000, 2000, 50, seed=42)
    cols = np.arange(0, 2000, 9, dtype=np.int32)
    with Staged(rp, ri, rv) as s:
        r = learn_columns(s, dict(niters=50), cols=cols)
        got, stats = r.to_host(), r.stats()
        _, h32, h16 = s.gram_layout()
    assert h32 < h16 < 2048
    ref = oracle.learn(rp, ri, rv, niters=50, cols=cols, nthreads=8, want_stats=True, order=st.ORDER_POPULARITY)
    assert np.array_equal(stats["niters"], ref["stats"]["niters"])
    _check_close(got, ref, tol=1e-6)
    assert np.allclose(stats["objval"], ref["stats"]["objval"], rtol=1e-9)


# ---- STAIR layout of G (gram.cuh: GaStair): the layout when the full N x N matrix does not fit (BASELINE configs[4]) ----

@pytest.mark.parametrize("hd", ["64", "256", "1000000"])
def test_gram_stair_layout_is_exact(lib, monkeypatch, hd):
    # panel p keeps rows [0, max(64 (p + 1), hd)); every element read back through the mirror rule == R^T R
    import scipy.sparse as sp

    from slim_b200 import Staged

    monkeypatch.setenv("SLIMB200_GRAM_LAYOUT", "stair")
    monkeypatch.setenv("SLIMB200_GRAM_HD", hd)
    for ratings, nu, ni, per in ((False, 3000, 200, 25), (True, 9000, 1200, 12)):
        rp, ri, rv = st.synth_zipf(nu, ni, per, seed=8, ratings=ratings)
        R = sp.csr_matrix((rv.astype(np.float64), ri, rp), shape=(nu, ni))
        G = (R.T @ R).toarray()
        with Staged(rp, ri, rv) as s:
            got = s.gram()
            rank = s.item_order()
            nbytes, h32, h16 = s.gram_layout()
            stair, hdv = s.gram_stair()
        ld = (ni + 127) // 128 * 128
        assert stair == 1 and hdv == min((int(hd) + 63) // 64 * 64, ld)
        inv = np.argsort(rank)
        assert np.array_equal(got.astype(np.float64), G[np.ix_(inv, inv)])
        pan = np.arange(ld // 64)
        rows = np.minimum(ni, np.maximum((pan + 1) * 64, hdv))
        width = np.where(pan * 64 < h32, 4, np.where(pan * 64 < h16, 2, 1))
        assert nbytes == int((rows * 64 * width).sum()) + 16


@pytest.mark.parametrize("hd", ["64", "256"])
@pytest.mark.parametrize("route", ["single", "cluster", "user", "mixed", "user-cluster", "mixed-cluster"])
def test_gram_stair_kernels(lib, oracle, monkeypatch, hd, route):
    # one-target Gram kernels on the stair layout (direct and mirrored elements, all three widths); "user": every
    # target above the giant threshold goes to cd_hybrid_kernel in the same call ("-cluster": to the user-space
    # cd_cluster_kernel); "mixed": both
    if route.endswith("-cluster"):
        monkeypatch.setenv("SLIMB200_GIANT_KERNEL", "cluster")
        route = route[:-8]
    monkeypatch.setenv("SLIMB200_GRAM_LAYOUT", "stair")
    monkeypatch.setenv("SLIMB200_GRAM_HD", hd)
    monkeypatch.setenv("SLIMB200_GRAM_HEAVY", "0" if route == "cluster" else ("60" if route == "mixed" else "1000000000"))
    monkeypatch.setenv("SLIMB200_STAIR_USER", {"single": "1000000000", "cluster": "1000000000", "user": "0",
                                               "mixed": "150"}[route])
    from slim_b200 import Staged, learn_columns

    rp, ri, rv = st.synth_zipf(9000, 1200, 12, seed=8, ratings=True)
    cols = np.arange(0, 1200, 5, dtype=np.int32)
    with Staged(rp, ri, rv) as s:
        assert s.gram_stair()[0] == 1
        r = learn_columns(s, dict(niters=40), cols=cols)
        got, stats = r.to_host(), r.stats()
    ref = oracle.learn(rp, ri, rv, niters=40, cols=cols, nthreads=8, want_stats=True, order=st.ORDER_POPULARITY)
    assert np.array_equal(stats["nactive"], ref["stats"]["nactive"])
    assert np.array_equal(stats["niters"], ref["stats"]["niters"])
    _check_close(got, ref, tol=1e-6)
    assert np.allclose(stats["objval"], ref["stats"]["objval"], rtol=1e-9)


def test_gram_stair_full_model_golden(lib, ours, monkeypatch):
    # SLIM_Learn on the stair layout with giants in user space == the reference golden (ml100k, converged)
    monkeypatch.setenv("SLIMB200_GRAM_LAYOUT", "stair")
    monkeypatch.setenv("SLIMB200_GRAM_HD", "128")
    monkeypatch.setenv("SLIMB200_STAIR_USER", "250")
    g = st.load_golden("ml100k")
    h = _learn(ours, g["trn_rowptr"], g["trn_rowind"], g["trn_rowval"], l1r=1.0, l2r=1.0, **CONV)
    _check_close(st.model_views(h), _golden_model(g, "conv"))
    ours.free(h)


def test_gram_stair_warm_start_and_fslim(lib, ours, oracle, monkeypatch):
    monkeypatch.setenv("SLIMB200_GRAM_LAYOUT", "stair")
    monkeypatch.setenv("SLIMB200_GRAM_HD", "64")
    monkeypatch.setenv("SLIMB200_STAIR_USER", "120")
    rp, ri, rv = st.synth_zipf(900, 200, 20, seed=17, ratings=True)
    h0 = _learn(ours, rp, ri, rv, l1r=3.0, l2r=1.0, niters=30)
    m0 = st.model_views(h0)
    h1 = _learn(ours, rp, ri, rv, imodel=h0, l1r=1.0, l2r=1.0, niters=5)
    w1 = oracle.learn(rp, ri, rv, l1r=1.0, l2r=1.0, niters=5, nthreads=4,
                      imodel=(m0["ncols"], m0["colptr"], m0["colind"], m0["colval"]), order=st.ORDER_POPULARITY)
    _check_close(st.model_views(h1), w1, tol=1e-6)
    for h in (h0, h1):
        ours.free(h)
    # fSLIM on the stair layout: the reference's golden model (neighbour sets incl. boundary ties)
    g, f = st.load_golden("ml100k"), st.load_golden("fslim")
    io, do = st.options(l1r=1.0, l2r=1.0, nnbrs=int(f["nnbrs"]), simtype="cos", **CONV)
    h, status = ours.learn(g["trn_rowptr"], g["trn_rowind"], g["trn_rowval"], io, do)
    assert h and status == st.SLIM_OK
    ref = dict(colptr=f["ml100k_cos_colptr"], colind=f["ml100k_cos_colind"], colval=f["ml100k_cos_colval"])
    _check_close(st.model_views(h), ref)
    ours.free(h)


# ---- giant targets next to a resident Gram matrix (slim_b200/csrc/hybrid.cuh): user-space inner products, Gram tiles ----

@pytest.mark.parametrize("layout", ["stair", "full"])
@pytest.mark.parametrize("cs", ["1", "4", "8", "16"])
@pytest.mark.parametrize("ratings,null_vals", [(False, False), (True, False), (False, True)])
def test_hybrid_kernel_variants(lib, ours, oracle, monkeypatch, layout, cs, ratings, null_vals):
    monkeypatch.setenv("SLIMB200_GRAM_LAYOUT", layout)
    monkeypatch.setenv("SLIMB200_GRAM_HD", "64")
    monkeypatch.setenv("SLIMB200_HYBRID_MIN", "0")  # every target
    monkeypatch.setenv("SLIMB200_HYBRID_CS", cs)
    rp, ri, rv = st.synth_zipf(1000, 260, 24, seed=29, ratings=ratings)
    if null_vals:
        rv = None
    kw = dict(l1r=0.7, l2r=1.5, **CONV)
    h = _learn(ours, rp, ri, rv, **kw)
    w = oracle.learn(rp, ri, rv, nthreads=8, **kw, order=st.ORDER_POPULARITY)
    _check_close(st.model_views(h), w)
    ours.free(h)


@pytest.mark.parametrize("cs,hmin", [("1", "0"), ("2", "0"), ("16", "0"), ("16", "2000")])
def test_hybrid_long_columns_and_iteration_cap(lib, oracle, monkeypatch, cs, hmin):
    # 40 000 users: column ranges of every class (one lane, 8 lanes, one warp, the whole CTA), several blocks of 128
    # active coordinates, capped head targets that must follow the oracle sweep by sweep; "2000": the lighter targets
    # run on the Gram kernels in the same call
    monkeypatch.setenv("SLIMB200_GRAM_LAYOUT", "stair")
    monkeypatch.setenv("SLIMB200_GRAM_HD", "128")
    monkeypatch.setenv("SLIMB200_HYBRID_MIN", hmin)
    monkeypatch.setenv("SLIMB200_HYBRID_CS", cs)
    from slim_b200 import Staged, learn_columns

    rp, ri, rv = st.synth_zipf(40000, 600, 30, seed=77)
    cols = np.arange(0, 600, 13, dtype=np.int32)
    with Staged(rp, ri, rv) as s:
        r = learn_columns(s, dict(niters=30), cols=cols)
        got, stats = r.to_host(), r.stats()
    ref = oracle.learn(rp, ri, rv, niters=30, cols=cols, nthreads=8, want_stats=True, order=st.ORDER_POPULARITY)
    assert np.array_equal(stats["nactive"], ref["stats"]["nactive"])
    assert np.array_equal(stats["active_nnz"], ref["stats"]["active_nnz"])
    assert np.array_equal(stats["niters"], ref["stats"]["niters"])
    _check_close(got, ref, tol=1e-6)
    assert np.allclose(stats["objval"], ref["stats"]["objval"], rtol=1e-9, atol=1e-9)
    assert np.allclose(stats["rnorm"], ref["stats"]["rnorm"], rtol=1e-9, atol=1e-7)


@pytest.mark.parametrize("cs", ["1", "16"])
def test_hybrid_warm_start_and_real_ratings(lib, ours, oracle, monkeypatch, cs):
    monkeypatch.setenv("SLIMB200_HYBRID_MIN", "0")
    monkeypatch.setenv("SLIMB200_HYBRID_CS", cs)
    rp, ri, rv = st.synth_zipf(900, 200, 20, seed=17, ratings=True)
    h0 = _learn(ours, rp, ri, rv, l1r=3.0, l2r=1.0, niters=30)
    m0 = st.model_views(h0)
    h1 = _learn(ours, rp, ri, rv, imodel=h0, l1r=1.0, l2r=1.0, niters=5)
    w1 = oracle.learn(rp, ri, rv, l1r=1.0, l2r=1.0, niters=5, nthreads=4,
                      imodel=(m0["ncols"], m0["colptr"], m0["colind"], m0["colval"]), order=st.ORDER_POPULARITY)
    _check_close(st.model_views(h1), w1, tol=1e-6)
    for h in (h0, h1):
        ours.free(h)


# ---- batched heavy-target kernel (slim_b200/csrc/gram_batch.cuh): 8 targets per cluster, item-space blocks ----

@pytest.mark.parametrize("bcs", ["1", "4", "8", "16"])
@pytest.mark.parametrize("ratings,null_vals", [(False, False), (True, False), (False, True)])
def test_gram_batch_kernel_variants(lib, ours, oracle, monkeypatch, bcs, ratings, null_vals):
    monkeypatch.setenv("SLIMB200_BATCH_CS", bcs)
    monkeypatch.setenv("SLIMB200_GRAM_BATCH", "0")  # every target goes through the batched kernel
    monkeypatch.setenv("SLIMB200_GRAM_HEAVY", "0")
    rp, ri, rv = st.synth_zipf(1000, 260, 24, seed=29, ratings=ratings)
    if null_vals:
        rv = None
    kw = dict(l1r=0.7, l2r=1.5, **CONV)
    h = _learn(ours, rp, ri, rv, **kw)
    w = oracle.learn(rp, ri, rv, nthreads=8, **kw, order=st.ORDER_POPULARITY)
    _check_close(st.model_views(h), w)
    ours.free(h)


@pytest.mark.parametrize("bcs,batch,heavy", [("1", "0", "0"), ("8", "0", "0"), ("16", "0", "0"), ("4", "900", "200"),
                                             ("8", "2000", "300")])
def test_gram_batch_long_columns_and_iteration_cap(lib, ours, oracle, monkeypatch, bcs, batch, heavy):
    # capped head targets follow the oracle sweep by sweep; mixed launches: batched + cluster + single CTA classes
    monkeypatch.setenv("SLIMB200_BATCH_CS", bcs)
    monkeypatch.setenv("SLIMB200_GRAM_BATCH", batch)
    monkeypatch.setenv("SLIMB200_GRAM_HEAVY", heavy)
    rp, ri, rv = st.synth_zipf(6000, 400, 40, seed=21)
    from slim_b200 import Staged, learn_columns

    cols = np.arange(0, 400, 7, dtype=np.int32)  # 58 targets: the last batch is short
    with Staged(rp, ri, rv) as s:
        r = learn_columns(s, dict(niters=50), cols=cols)
        got, stats = r.to_host(), r.stats()
    ref = oracle.learn(rp, ri, rv, niters=50, cols=cols, nthreads=8, want_stats=True, order=st.ORDER_POPULARITY)
    assert np.array_equal(stats["nactive"], ref["stats"]["nactive"])
    assert np.array_equal(stats["active_nnz"], ref["stats"]["active_nnz"])
    assert np.array_equal(stats["expand_nnz"], ref["stats"]["expand_nnz"])
    assert np.array_equal(stats["niters"], ref["stats"]["niters"])
    _check_close(got, ref, tol=1e-6)
    assert np.allclose(stats["objval"], ref["stats"]["objval"], rtol=1e-9, atol=1e-9)
    assert np.allclose(stats["rnorm"], ref["stats"]["rnorm"], rtol=1e-9, atol=1e-7)


@pytest.mark.parametrize("bcs", ["1", "8"])
def test_gram_batch_warm_start_and_real_values(lib, ours, oracle, monkeypatch, bcs):
    monkeypatch.setenv("SLIMB200_BATCH_CS", bcs)
    monkeypatch.setenv("SLIMB200_GRAM_BATCH", "0")
    monkeypatch.setenv("SLIMB200_GRAM_HEAVY", "0")
    rp, ri, rv = st.synth_zipf(900, 200, 20, seed=17, ratings=True)
    rv = (rv * np.random.default_rng(3).uniform(0.2, 1.0, len(rv))).astype(np.float32)  # fp64 Gram matrix
    h0 = _learn(ours, rp, ri, rv, l1r=3.0, l2r=1.0, niters=30)
    m0 = st.model_views(h0)
    w0 = oracle.learn(rp, ri, rv, l1r=3.0, l2r=1.0, niters=30, nthreads=4, order=st.ORDER_POPULARITY)
    _check_close(m0, w0, tol=1e-6)
    h1 = _learn(ours, rp, ri, rv, imodel=h0, l1r=1.0, l2r=1.0, niters=5)
    w1 = oracle.learn(rp, ri, rv, l1r=1.0, l2r=1.0, niters=5, nthreads=4,
                      imodel=(m0["ncols"], m0["colptr"], m0["colind"], m0["colval"]), order=st.ORDER_POPULARITY)
    _check_close(st.model_views(h1), w1, tol=1e-6)
    for h in (h0, h1):
        ours.free(h)


def test_edge_cases(lib, ours, oracle):
    # empty rows, an empty column, single-entry columns, and maxniters = 0
    rp = np.array([0, 0, 2, 2, 5, 6], np.int64)
    ri = np.array([1, 3, 1, 2, 3, 3], np.int32)
    rv = np.array([1, 2, 3, 1, 1, 4], np.float32)
    for kw in (dict(l1r=0.1, l2r=0.5), dict(l1r=0.1, l2r=0.5, niters=0), dict(l1r=0.0, l2r=0.0)):
        h = _learn(ours, rp, ri, rv, **kw)
        mv = st.model_views(h)
        w = oracle.learn(rp, ri, rv, **{**dict(opttol=1e-7, niters=10000, order=st.ORDER_POPULARITY), **kw})
        assert mv["ncols"] == 4
        _check_close(mv, w, tol=1e-6)
        ours.free(h)
    # a matrix with no nonzeros at all
    h = _learn(ours, np.zeros(4, np.int64), np.zeros(0, np.int32), np.zeros(0, np.float32))
    mv = st.model_views(h)
    assert mv["ncols"] == 0 and mv["nrows"] == 0
    ours.free(h)


def test_column_shards_reassemble_to_the_full_model(lib, oracle):
    from slim_b200 import Staged, learn_columns

    rp, ri, rv = st.synth_zipf(1500, 300, 16, seed=3)
    with Staged(rp, ri, rv) as s:
        full = learn_columns(s, dict(niters=40)).to_host()
        parts = [learn_columns(s, dict(niters=40), cols=np.arange(r, 300, 3, dtype=np.int32)).to_host()
                 for r in range(3)]
        with pytest.raises(RuntimeError):
            learn_columns(s, cols=np.array([300], np.int32))
    for r, p in enumerate(parts):
        for q, j in enumerate(range(r, 300, 3)):
            a, b = full["colptr"][j], full["colptr"][j + 1]
            c, d = p["colptr"][q], p["colptr"][q + 1]
            assert np.array_equal(full["colind"][a:b], p["colind"][c:d])
            assert np.array_equal(full["colval"][a:b], p["colval"][c:d])  # bit-identical


def test_medium_synthetic_objective_property(lib, oracle):
    # 20K x 2K, 1M nnz Zipf(1.1) (the survey's probe): size-independent properties on all columns,
    # oracle comparison on a stratified column sample
    from slim_b200 import Staged, learn_columns

    rp, ri, rv = st.synth_zipf(20000, 2000, 50, seed=42)
    with Staged(rp, ri, rv) as s:
        r = learn_columns(s, dict(niters=50))
        w, stats = r.to_host(), r.stats()
        cnt = np.diff(s.csc()["colptr"])
    assert (w["colval"] > 0).all()
    assert (stats["objval"] >= stats["rnorm"] - 1e-9).all()
    assert (stats["niters"] >= 1).all() and (stats["niters"] <= 51).all()
    # the objective at x = 0 is 1/2 |y|^2: CD never ends above it
    assert (stats["objval"] <= 0.5 * cnt + 1e-6).all()
    order = np.argsort(cnt, kind="stable")
    sample = np.sort(order[:: len(order) // 48]).astype(np.int32)
    ref = oracle.learn(rp, ri, rv, niters=50, cols=sample, nthreads=8, want_stats=True, order=st.ORDER_POPULARITY)
    assert np.allclose(stats["objval"][sample], ref["stats"]["objval"], rtol=1e-6)
    sub = dict(colptr=np.concatenate([[0], np.cumsum(np.diff(w["colptr"])[sample])]),
               colind=np.concatenate([w["colind"][w["colptr"][j]:w["colptr"][j + 1]] for j in sample]),
               colval=np.concatenate([w["colval"][w["colptr"][j]:w["colptr"][j + 1]] for j in sample]))
    _check_close(sub, ref, tol=1e-5)


def test_python_mirror_train_predict(lib, ml100k):
    import scipy.sparse as sp

    from slim_b200 import SLIM, SLIMatrix

    g = ml100k
    R = sp.csr_matrix((g["trn_rowval"], g["trn_rowind"], g["trn_rowptr"]), shape=(934, 1683))
    mat = SLIMatrix(R)
    model = SLIM()
    model.train({"algo": "cd", "l1r": 1.0, "l2r": 1.0, "optTol": 1e-14, "niters": 100000}, mat)
    out = model.predict(mat, nrcmds=10)
    got = np.stack([out[u] for u in range(934)])
    bad = np.nonzero((got != g["top10_ids"]).any(axis=1))[0]
    assert set(bad.tolist()) <= {277}
    assert abs(model.to_csr().nnz - 65909) <= 2


def test_gpu_topn_matches_host_loop(lib, ours, monkeypatch):
    # Py_SLIM_Predict: the batched GPU top-N (predict.cuh) against the per-user host loop of the reference
    # restatement -- identical lists AND bit-identical float scores, short lists keep the caller's filler
    from slim_b200 import SLIM, SLIMatrix
    import scipy.sparse as sp

    for name, nr in (("ml100k", 10), ("automotive", 25)):
        g = st.load_golden(name)
        shape = (len(g["trn_rowptr"]) - 1, int(g["trn_rowind"].max()) + 1)
        trn = SLIMatrix(sp.csr_matrix((g["trn_rowval"], g["trn_rowind"], g["trn_rowptr"]), shape=shape))
        model = SLIM()
        model.train({"algo": "cd", "l1r": 1.0, "l2r": 1.0, "niters": 30}, trn)
        monkeypatch.setenv("SLIMB200_PREDICT_HOST", "1")
        host_ids, host_sc = model.predict(trn, nrcmds=nr, returnscores=True)
        monkeypatch.setenv("SLIMB200_PREDICT_HOST", "0")
        gpu_ids, gpu_sc = model.predict(trn, nrcmds=nr, returnscores=True)
        for u in range(shape[0]):
            assert np.array_equal(np.asarray(host_ids[u]), np.asarray(gpu_ids[u])), (name, u)
            assert np.array_equal(np.asarray(host_sc[u], np.float32).view(np.uint32),
                                  np.asarray(gpu_sc[u], np.float32).view(np.uint32)), (name, u)


def test_model_selection_keeps_matrix_resident(lib, automotive, capfd):
    # Py_SLIM_Mselect (reference pyapi.c:214-412): (l1, l2) grid with warm starts; here R is staged once
    from slim_b200 import SLIM, SLIMatrix
    import scipy.sparse as sp

    g = automotive
    trn = SLIMatrix(sp.csr_matrix((g["trn_rowval"], g["trn_rowind"], g["trn_rowptr"]), shape=(2928, 1835)))
    tst = SLIMatrix(sp.csr_matrix((g["tst_rowval"], g["tst_rowind"], g["tst_rowptr"]), shape=(2928, 1835)))
    model = SLIM()
    best = model.mselect({"niters": 100, "nthreads": 1}, trn, tst, [1.0, 5.0], [1.0], nrcmds=10)
    # the (1, 1) cell is the golden configuration: HR 0.1059 at convergence, ~0.106 after 100 sweeps
    assert 0.09 < best["bestHRHR"] < 0.13 and best["bestl2HR"] == 1.0
    assert best["bestl1HR"] in (1.0, 5.0) and best["bestARAR"] > 0.04
    out = capfd.readouterr().out
    assert out.count("Using Coordinate Descent!") == 2 and "hr_head" in out


# ---- fSLIM (nnbrs > 0): fslim.cuh + cd_gram_kernel; reference neighbors.c:16-125, estimate.c:424-431 ----------

@pytest.mark.parametrize("name", ["ml100k", "automotive"])
@pytest.mark.parametrize("sim", ["cos", "jac", "dotp"])
def test_fslim_matches_reference_golden(lib, ours, name, sim):
    # same neighbour SETS as the reference, boundary ties included (the kernel rebuilds the reference's candidate
    # order and runs its selection), weights within the converged tolerance
    g, f = st.load_golden(name), st.load_golden("fslim")
    nn = int(f["nnbrs"])
    io, do = st.options(l1r=1.0, l2r=1.0, nnbrs=nn, simtype=sim, **CONV)
    h, status = ours.learn(g["trn_rowptr"], g["trn_rowind"], g["trn_rowval"], io, do)
    assert h and status == st.SLIM_OK
    mv = st.model_views(h)
    ref = dict(colptr=f[f"{name}_{sim}_colptr"], colind=f[f"{name}_{sim}_colind"], colval=f[f"{name}_{sim}_colval"])
    _check_close(mv, ref)
    assert np.diff(mv["colptr"]).max() <= nn
    ours.free(h)


@pytest.mark.parametrize("binary", [False, True])
def test_fslim_default_setting_matches_oracle(lib, ours, oracle, binary):
    # capped / default tolerance, same visiting order as the oracle: iterate-level agreement
    rp, ri, rv = st.synth_zipf(6000, 500, 30, seed=5, ratings=not binary)
    io, do = st.options(l1r=0.3, l2r=1.0, niters=50, nnbrs=25, simtype="cos")
    h, status = ours.learn(rp, ri, None if binary else rv, io, do)
    assert h and status == st.SLIM_OK
    w = oracle.learn(rp, ri, None if binary else rv, l1r=0.3, l2r=1.0, niters=50, nthreads=8, nnbrs=25,
                     simtype="cos", nbr_ties=st.TIES_REFERENCE, order=st.ORDER_POPULARITY)
    _check_close(st.model_views(h), w, tol=1e-6)
    ours.free(h)


def test_ordered_flag_keeps_the_plain_slim_path(lib, ours):
    # api.c:54-60 + estimate.c:396,424: `ordered` only relabels the model type; with nnbrs > 0 it even switches the
    # neighbour restriction OFF (only SLIM_MTYPE_FSLIM is special-cased)
    g = st.load_golden("automotive")
    base = _learn(ours, g["trn_rowptr"], g["trn_rowind"], g["trn_rowval"], niters=50)
    ref = st.model_views(base)
    for nnbrs in (0, 10):
        io, do = st.options(l1r=1.0, l2r=1.0, niters=50, nnbrs=nnbrs)
        io[st.OPT_ORDERED] = 1
        h, status = ours.learn(g["trn_rowptr"], g["trn_rowind"], g["trn_rowval"], io, do)
        assert h and status == st.SLIM_OK
        mv = st.model_views(h)
        assert np.array_equal(mv["colind"], ref["colind"]) and np.array_equal(mv["colval"], ref["colval"])
        ours.free(h)
    ours.free(base)


# ---- model selection pinned to the reference (pyapi.c:214-412) -----------------------------------------------

def test_mselect_matches_reference_golden(lib, ours):
    g, ms = st.load_golden("automotive"), st.load_golden("mselect")
    trn = (g["trn_rowptr"], g["trn_rowind"], g["trn_rowval"])
    tst = (g["tst_rowptr"], g["tst_rowind"], g["tst_rowval"])
    rc, best, cells = st.mselect(ours, trn, tst, ms["l1"], ms["l2"], nrcmds=10, **CONV)
    assert rc == st.SLIM_OK
    assert cells.shape == ms["cells"].shape
    assert np.array_equal(cells[:, :2], ms["cells"][:, :2])                       # the (l1, l2) of every cell
    assert np.abs(cells[:, 2] - ms["cells"][:, 2]).max() <= 2                     # nnz(W): support flips below 2e-6
    assert np.array_equal(cells[:, 3:], ms["cells"][:, 3:]), (cells, ms["cells"])  # HR / head / tail / ARHR, 4 dp
    assert np.array_equal(best[[0, 1, 4, 5]], ms["best"][[0, 1, 4, 5]])           # the chosen (l1, l2) pairs
    assert np.allclose(best[[2, 3, 6, 7]], ms["best"][[2, 3, 6, 7]], rtol=0, atol=5e-5)


def test_automotive_gpu_topn_matches_golden_lists(lib, ours):
    # batched GPU top-N (predict.cuh) on the second fixture against the reference's own lists
    g = st.load_golden("automotive")
    h = _learn(ours, g["trn_rowptr"], g["trn_rowind"], g["trn_rowval"], **CONV)
    from slim_b200 import _lib as L_

    L = L_.load()
    mats = []
    for rp, ri, rv in ((g["trn_rowptr"], g["trn_rowind"], g["trn_rowval"]),):
        rp, ri, rv = (np.ascontiguousarray(rp, np.int64), np.ascontiguousarray(ri, np.int32),
                      np.ascontiguousarray(rv, np.float32))
        hm = C.c_void_p()
        assert L.Py_csr_wrapper(len(rp) - 1, rp.ctypes.data_as(C.POINTER(C.c_ssize_t)),
                                ri.ctypes.data_as(C.POINTER(C.c_int32)), rv.ctypes.data_as(C.POINTER(C.c_float)),
                                C.byref(hm)) == st.SLIM_OK
        mats.append((hm, rp, ri, rv))
    nu = len(g["trn_rowptr"]) - 1
    ids = np.full((nu, 10), -1, np.int32)
    sc = np.zeros((nu, 10), np.float32)
    assert L.Py_SLIM_Predict(10, h, mats[0][0], ids.ctypes.data_as(C.POINTER(C.c_int32)),
                             sc.ctypes.data_as(C.POINTER(C.c_float))) == st.SLIM_OK
    assert np.array_equal(ids, g["top10_ids"])
    L.Py_csr_free(mats[0][0])
    ours.free(h)
