"""CPU tests that PIN the oracle (oracle/slim_oracle.c) to the reference.

 * against tests/golden/*.npz, produced by the unmodified reference (make_golden.py);
 * against the reference itself, live, when oracle/_ref/libslim_ref.so is present.
"""
import numpy as np
import pytest

import slimtest as st

CONV = dict(opttol=1e-14, niters=100000)
TOL = 2e-6  # per-nonzero tolerance at the converged setting


def _golden_model(g, tag):
    return dict(colptr=g[f"W_{tag}_colptr"], colind=g[f"W_{tag}_colind"], colval=g[f"W_{tag}_colval"])


@pytest.mark.parametrize("name", ["ml100k", "automotive"])
def test_reference_order_mode_is_bit_identical_to_reference_defaults(oracle, name):
    g = st.load_golden(name)
    st.libc_srand(1)
    w = oracle.learn(g["trn_rowptr"], g["trn_rowind"], g["trn_rowval"], order=st.ORDER_REF_RAND)
    ref = _golden_model(g, "default")
    assert np.array_equal(w["colptr"], ref["colptr"])
    assert np.array_equal(w["colind"], ref["colind"])
    assert np.array_equal(w["colval"].view(np.uint32), ref["colval"].view(np.uint32))


@pytest.mark.parametrize("name", ["ml100k", "automotive"])
def test_ascending_mode_matches_converged_golden(oracle, name):
    g = st.load_golden(name)
    w = oracle.learn(g["trn_rowptr"], g["trn_rowind"], g["trn_rowval"], nthreads=8, **CONV)
    ref = _golden_model(g, "conv")
    maxd, flips = st.compare_models(w, ref)
    assert maxd <= TOL, maxd
    assert all(mag < TOL for _, _, mag in flips), flips[:5]
    # predictions: in-memory top-10 lists and the CLI's HR/ARHR
    n = w["ncols"]
    rp, ri, rv = oracle.transpose(n, w["colptr"], w["colind"], w["colval"])
    ids, _ = oracle.topn_all(n, rp, ri, rv, g["trn_rowptr"], g["trn_rowind"], g["trn_rowval"], 10)
    bad = np.nonzero((ids != g["top10_ids"]).any(axis=1))[0]
    if name == "ml100k":
        # user 277's 10th/11th scores differ by 5.7e-7 (items 748 / 176)
        assert set(bad.tolist()) <= {277}, bad
    else:
        assert len(bad) == 0, bad
    ev = st.evaluate(ids, (g["trn_rowptr"], g["trn_rowind"]), (g["tst_rowptr"], g["tst_rowind"]),
                     n, g["fmarker"])
    got = np.array([ev["hr"], ev["hr_head"], ev["hr_tail"], ev["arhr"]])
    assert np.array_equal(np.round(got, 4), np.round(g["metrics"], 4)), (got, g["metrics"])


def test_popularity_order_mode(oracle):
    # a different FIXED visiting order (the CUDA engine's) reaches the same optimum
    g = st.load_golden("automotive")
    w = oracle.learn(g["trn_rowptr"], g["trn_rowind"], g["trn_rowval"], nthreads=8,
                     order=st.ORDER_POPULARITY, **CONV)
    maxd, flips = st.compare_models(w, _golden_model(g, "conv"))
    assert maxd <= TOL and all(mag < TOL for _, _, mag in flips)
    # ... and differs from the ascending order when the sweeps are capped (order matters off the optimum)
    a = oracle.learn(g["trn_rowptr"], g["trn_rowind"], g["trn_rowval"], niters=2, nthreads=8)
    b = oracle.learn(g["trn_rowptr"], g["trn_rowind"], g["trn_rowval"], niters=2, nthreads=8,
                     order=st.ORDER_POPULARITY)
    assert st.compare_models(a, b)[0] > 1e-6


def test_golden_values_are_the_surveyed_ones(ml100k, automotive):
    assert len(ml100k["W_conv_colind"]) == 65909
    assert len(automotive["W_conv_colind"]) == 84317
    assert np.allclose(ml100k["metrics"], [0.3191, 0.5119, 0.0930, 0.1504], atol=5e-5)
    assert np.allclose(automotive["metrics"], [0.1059, 0.1664, 0.0544, 0.0530], atol=5e-5)


def _small(seed, ratings):
    return st.synth_zipf(300, 120, 12, seed=seed, ratings=ratings)


@pytest.mark.skipif(not st.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("ratings", [False, True])
@pytest.mark.parametrize("binary_null", [False, True])
def test_live_reference_bit_identical(oracle, ratings, binary_null):
    rp, ri, rv = _small(7, ratings)
    if binary_null:
        rv = None
    ref = st.load_ref()
    io, do = st.options(l1r=0.5, l2r=2.0, nthreads=1, niters=40)
    st.libc_srand(1)
    h, status = ref.learn(rp, ri, rv, io, do)
    assert status == st.SLIM_OK
    mv = st.model_views(h)
    st.libc_srand(1)
    w = oracle.learn(rp, ri, rv, l1r=0.5, l2r=2.0, niters=40, order=st.ORDER_REF_RAND)
    assert np.array_equal(w["colptr"], mv["colptr"])
    assert np.array_equal(w["colind"], mv["colind"])
    assert np.array_equal(w["colval"].view(np.uint32), mv["colval"].view(np.uint32))
    # CSR view of the model (SaveModel + CreateIndex(ROW))
    tp, ti, tv = oracle.transpose(mv["ncols"], w["colptr"], w["colind"], w["colval"])
    assert np.array_equal(tp, mv["rowptr"]) and np.array_equal(ti, mv["rowind"])
    assert np.array_equal(tv.view(np.uint32), mv["rowval"].view(np.uint32))
    # top-N restatement vs the reference's GetRecommendations
    ids_r, sc_r = ref.topn_all(h, rp, ri, rv, 5)
    ids_o, sc_o = oracle.topn_all(mv["ncols"], tp, ti, tv, rp, ri, rv, 5)
    assert np.array_equal(sc_r.view(np.uint32), sc_o.view(np.uint32))
    same = ids_r == ids_o
    # ids may differ only inside exact score ties (reference sort is not stable)
    for u, k in zip(*np.nonzero(~same)):
        assert (sc_r[u] == sc_r[u, k]).sum() > 1
    ref.free(h)


@pytest.mark.skipif(not st.have_ref(), reason="oracle/_ref not built")
def test_live_reference_warm_start(oracle):
    rp, ri, rv = _small(11, True)
    ref = st.load_ref()
    io, do = st.options(l1r=2.0, l2r=1.0, nthreads=1, niters=30)
    st.libc_srand(1)
    h0, _ = ref.learn(rp, ri, rv, io, do)
    m0 = st.model_views(h0)
    io2, do2 = st.options(l1r=1.0, l2r=1.0, nthreads=1, niters=30)
    st.libc_srand(1)
    h1, _ = ref.learn(rp, ri, rv, io2, do2, imodel=h0)
    m1 = st.model_views(h1)
    st.libc_srand(1)
    w = oracle.learn(rp, ri, rv, l1r=1.0, l2r=1.0, niters=30, order=st.ORDER_REF_RAND,
                     imodel=(m0["ncols"], m0["colptr"], m0["colind"], m0["colval"]))
    assert np.array_equal(w["colind"], m1["colind"])
    assert np.array_equal(w["colval"].view(np.uint32), m1["colval"].view(np.uint32))
    ref.free(h0)
    ref.free(h1)


def test_column_subset_and_stats(oracle):
    rp, ri, rv = _small(3, False)
    full = oracle.learn(rp, ri, rv, niters=50, want_stats=True)
    cols = np.array([5, 0, 77, 119], np.int32)
    sub = oracle.learn(rp, ri, rv, niters=50, cols=cols, want_stats=True)
    for q, j in enumerate(cols):
        a, b = full["colptr"][j], full["colptr"][j + 1]
        c, d = sub["colptr"][q], sub["colptr"][q + 1]
        assert np.array_equal(full["colind"][a:b], sub["colind"][c:d])
        assert np.array_equal(full["colval"][a:b], sub["colval"][c:d])
        assert full["stats"]["niters"][j] == sub["stats"]["niters"][q]
    assert (full["stats"]["objval"] >= full["stats"]["rnorm"]).all()


def test_edge_cases(oracle):
    # empty rows, an empty column (id 0 unused), a single-user column; maxniters = 0
    rp = np.array([0, 0, 2, 2, 5, 6], np.int64)
    ri = np.array([1, 3, 1, 2, 3, 3], np.int32)
    rv = np.array([1, 2, 3, 1, 1, 4], np.float32)
    w = oracle.learn(rp, ri, rv, l1r=0.1, l2r=0.5)
    assert w["ncols"] == 4 and w["colptr"][1] == 0  # empty column 0 -> empty model column
    w0 = oracle.learn(rp, ri, rv, l1r=0.1, l2r=0.5, niters=0)
    assert w0["colptr"][-1] == 0
    # strict zero diagonal and non-negativity
    for j in range(4):
        seg = slice(w["colptr"][j], w["colptr"][j + 1])
        assert j not in w["colind"][seg].tolist()
        assert (w["colval"][seg] > 0).all()


# ---- fSLIM (nnbrs > 0): reference src/libslim/neighbors.c:16-125 + estimate.c:424-431 ------------------------

@pytest.mark.parametrize("name", ["ml100k", "automotive"])
@pytest.mark.parametrize("sim", ["cos", "jac", "dotp"])
def test_fslim_matches_reference_golden(oracle, name, sim):
    # neighbour search restated with the reference's own selection (first-encounter candidate order + the
    # quickselect of lib/GKlib/fkvkselect.c): identical support, weights within the converged tolerance
    g, f = st.load_golden(name), st.load_golden("fslim")
    nn = int(f["nnbrs"])
    w = oracle.learn(g["trn_rowptr"], g["trn_rowind"], g["trn_rowval"], nthreads=8, nnbrs=nn, simtype=sim,
                     nbr_ties=st.TIES_REFERENCE, **CONV)
    ref = dict(colptr=f[f"{name}_{sim}_colptr"], colind=f[f"{name}_{sim}_colind"], colval=f[f"{name}_{sim}_colval"])
    assert np.diff(ref["colptr"]).max() <= nn
    maxd, flips = st.compare_models(w, ref)
    assert maxd <= TOL, maxd
    assert all(mag < TOL for _, _, mag in flips), flips[:5]
    # the visiting order is free at convergence
    w2 = oracle.learn(g["trn_rowptr"], g["trn_rowind"], g["trn_rowval"], nthreads=8, nnbrs=nn, simtype=sim,
                      nbr_ties=st.TIES_REFERENCE, order=st.ORDER_POPULARITY, **CONV)
    maxd2, flips2 = st.compare_models(w2, ref)
    assert maxd2 <= TOL and all(mag < TOL for _, _, mag in flips2)


def test_fslim_boundary_ties_are_common(oracle):
    # why the selection has to be restated literally: with a deterministic tie rule (similarity, then popularity,
    # then id) a large share of the ml100k columns ends up with a different -- equally similar -- neighbour set
    g, f = st.load_golden("ml100k"), st.load_golden("fslim")
    w = oracle.learn(g["trn_rowptr"], g["trn_rowind"], g["trn_rowval"], nthreads=8, nnbrs=10, simtype="cos",
                     nbr_ties=st.TIES_POPULARITY, niters=200)
    ref = dict(colptr=f["ml100k_cos_colptr"], colind=f["ml100k_cos_colind"], colval=f["ml100k_cos_colval"])
    _, flips = st.compare_models(w, ref)
    assert len({j for j, _, _ in flips}) > 100


def test_fslim_live_reference(oracle):
    if not st.have_ref():
        pytest.skip("oracle/_ref not built (no /root/reference on this box)")
    ref = st.load_ref()
    rp, ri, rv = st.synth_zipf(1500, 300, 15, seed=13, ratings=True)
    for sim in ("cos", "jac", "dotp"):
        io, do = st.options(l1r=0.5, l2r=2.0, nthreads=1, nnbrs=7, simtype=sim, **CONV)
        (h, status), _ = st.capture_stdout(lambda: ref.learn(rp, ri, rv, io, do))
        assert status == st.SLIM_OK
        mv = st.model_views(h)
        ref.free(h)
        w = oracle.learn(rp, ri, rv, l1r=0.5, l2r=2.0, nthreads=4, nnbrs=7, simtype=sim, **CONV)
        maxd, flips = st.compare_models(w, mv)
        assert maxd <= TOL and all(mag < TOL for _, _, mag in flips), (sim, maxd, flips[:3])
