"""CPU tests of the drop-in boundary: libslim.so builds, exports every symbol include/*.h declares,
and its host-side entry points (model assembly, top-N, model I/O, matrix wrapping) behave like the
reference's.  No learner calls here except the check that the learner FAILS without a GPU."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

import slimtest as st

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    from slim_b200 import _lib, build

    build.build()
    return _lib.load()


def _declared():
    names = set()
    for h in ("slim.h", "slim_b200.h"):
        txt = (ROOT / "include" / h).read_text()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        names |= set(re.findall(r"\b((?:SLIM|SLIMB200|Py)_\w+)\s*\(", txt))
    return names


def test_exports_every_declared_symbol(lib):
    from slim_b200 import _lib

    declared = _declared()
    assert len(declared) == 46
    assert declared == set(_lib.EXPORTED_SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None
    # the reference's 20 symbols
    ref20 = {n for n in declared if not n.startswith("SLIMB200_")}
    assert len(ref20) == 20


def test_defaults(lib):
    io = np.zeros(40, np.int32)
    do = np.zeros(40, np.float64)
    assert lib.SLIM_iSetDefaults(io.ctypes.data_as(C.POINTER(C.c_int32))) == 1
    assert lib.SLIM_dSetDefaults(do.ctypes.data_as(C.POINTER(C.c_double))) == 1
    assert (io == -1).all() and (do == -1).all()


def _assemble(lib, g, tag="conv"):
    cp = np.ascontiguousarray(g[f"W_{tag}_colptr"], np.int64)
    ci = np.ascontiguousarray(g[f"W_{tag}_colind"], np.int32)
    cv = np.ascontiguousarray(g[f"W_{tag}_colval"], np.float32)
    stt = C.c_int32(0)
    h = lib.SLIMB200_AssembleModel(len(cp) - 1, cp.ctypes.data_as(C.POINTER(C.c_int64)),
                                   ci.ctypes.data_as(C.POINTER(C.c_int32)),
                                   cv.ctypes.data_as(C.POINTER(C.c_float)), C.byref(stt))
    assert h and stt.value == 1
    return h


@pytest.mark.parametrize("name", ["ml100k", "automotive"])
def test_model_handle_layout_topn_and_metrics(lib, oracle, name):
    g = st.load_golden(name)
    h = _assemble(lib, g)
    mv = st.model_views(h)
    n = mv["ncols"]
    assert mv["nrows"] == n == len(g["W_conv_colptr"]) - 1
    tp, ti, tv = oracle.transpose(n, g["W_conv_colptr"], g["W_conv_colind"], g["W_conv_colval"])
    assert np.array_equal(mv["rowptr"], tp) and np.array_equal(mv["rowind"], ti)
    assert np.array_equal(mv["rowval"].view(np.uint32), tv.view(np.uint32))
    assert np.array_equal(mv["colind"], g["W_conv_colind"])
    # every pointer except the six model arrays is NULL
    m = C.cast(h, C.POINTER(st.GkCsr)).contents
    for f, _ in st.GkCsr._fields_[2:]:
        assert bool(getattr(m, f)) == (f in ("rowptr", "colptr", "rowind", "colind", "rowval", "colval"))
    # SLIM_GetTopN on the reference's own converged W reproduces the reference's lists
    ours = st.SlimLib(ROOT / "slim_b200" / "lib" / "libslim.so")
    ids, sc = ours.topn_all(h, g["trn_rowptr"], g["trn_rowind"], g["trn_rowval"], 10)
    assert np.array_equal(sc.view(np.uint32), g["top10_scores"].view(np.uint32))
    assert np.array_equal(ids, g["top10_ids"])
    ev = st.evaluate(ids, (g["trn_rowptr"], g["trn_rowind"]), (g["tst_rowptr"], g["tst_rowind"]), n,
                     g["fmarker"])
    got = np.array([ev["hr"], ev["hr_head"], ev["hr_tail"], ev["arhr"]])
    assert np.array_equal(np.round(got, 4), np.round(g["metrics"], 4))
    assert ours.free(h) is None  # SLIM_FreeModel nulls the caller's pointer


def test_model_io_roundtrip(lib, tmp_path, ml100k):
    h = _assemble(lib, ml100k)
    f = str(tmp_path / "model.binrow").encode()
    assert lib.SLIM_WriteModel(h, f) == 1
    h2 = lib.SLIM_ReadModel(f)
    a, b = st.model_views(h), st.model_views(h2)
    for k in ("rowptr", "rowind", "colptr", "colind"):
        assert np.array_equal(a[k], b[k])
    assert np.array_equal(a["rowval"], b["rowval"]) and np.array_equal(a["colval"], b["colval"])
    # text CSR used by the python wrapper's save_model / load_model ("%f": 6 decimals)
    t = str(tmp_path / "model.csr").encode()
    assert lib.Py_csr_save(h, t) == 1
    h3 = C.c_void_p()
    assert lib.Py_csr_load(C.byref(h3), t) == 1
    c = st.model_views(h3.value)
    assert np.array_equal(a["rowind"], c["rowind"])
    assert np.allclose(a["rowval"], c["rowval"], atol=5e-7)
    lib.Py_csr_free(h3)
    for x in (h, h2):
        hp = C.c_void_p(x)
        lib.SLIM_FreeModel(C.byref(hp))
        assert hp.value is None
    assert lib.SLIM_ReadModel(str(tmp_path / "missing").encode()) is None


def test_matrix_wrapper_stat_export(lib, automotive):
    rp = np.ascontiguousarray(automotive["trn_rowptr"], np.int64)
    ri = np.ascontiguousarray(automotive["trn_rowind"], np.int32)
    rv = np.ascontiguousarray(automotive["trn_rowval"], np.float32)
    h = C.c_void_p()
    assert lib.Py_csr_wrapper(len(rp) - 1, rp.ctypes.data_as(C.POINTER(C.c_ssize_t)),
                              ri.ctypes.data_as(C.POINTER(C.c_int32)),
                              rv.ctypes.data_as(C.POINTER(C.c_float)), C.byref(h)) == 1
    m = C.cast(h, C.POINTER(st.GkCsr)).contents
    assert (m.nrows, m.ncols) == (2928, 1835)
    nnz = C.c_int32(0)
    lib.Py_csr_stat(h, C.byref(nnz))
    assert nnz.value == 17545
    ip, ix, dv = np.zeros(len(rp), np.int32), np.zeros(nnz.value, np.int32), np.zeros(nnz.value, np.float32)
    lib.Py_csr_export(h, ip.ctypes.data_as(C.POINTER(C.c_int32)), ix.ctypes.data_as(C.POINTER(C.c_int32)),
                      dv.ctypes.data_as(C.POINTER(C.c_float)))
    assert np.array_equal(ip, rp) and np.array_equal(ix, ri) and np.array_equal(dv, rv)
    lib.Py_csr_free(h)


@pytest.mark.parametrize("name", ["ml100k", "automotive"])
def test_head_tail_split_matches_reference_marks(lib, name):
    # SLIM_DetermineHeadAndTail (api.c:215-245): the marks of the reference itself (golden `fmarker`), including the
    # items whose equal counts straddle the 50 % boundary -- their order comes from GKlib's unstable quicksort,
    # which api.cpp follows step by step
    g = st.load_golden(name)
    rp = np.ascontiguousarray(g["trn_rowptr"], np.int64)
    ri = np.ascontiguousarray(g["trn_rowind"], np.int32)
    ours = st.SlimLib(ROOT / "slim_b200" / "lib" / "libslim.so")
    fm = ours.head_tail(len(rp) - 1, len(g["fmarker"]), rp, ri)
    cnt = np.bincount(ri, minlength=len(fm))
    assert cnt[fm == 0].sum() >= rp[-1] // 2
    assert cnt[fm == 0].min() >= cnt[fm == 1].max()
    assert np.array_equal(fm, g["fmarker"])


def test_learner_fails_loudly_without_a_gpu(lib, ml100k, capfd):
    if lib.SLIMB200_DeviceCount() > 0:
        pytest.skip("a CUDA device is present")
    ours = st.SlimLib(ROOT / "slim_b200" / "lib" / "libslim.so")
    io, do = st.options()
    h, status = ours.learn(ml100k["trn_rowptr"], ml100k["trn_rowind"], ml100k["trn_rowval"], io, do)
    assert h is None and status != st.SLIM_OK
    assert b"no usable CUDA device" in lib.SLIMB200_LastError()
    stt = C.c_int32(0)
    rp = np.ascontiguousarray(ml100k["trn_rowptr"], np.int64)
    ri = np.ascontiguousarray(ml100k["trn_rowind"], np.int32)
    m = lib.SLIMB200_Stage(0, len(rp) - 1, rp.ctypes.data_as(C.POINTER(C.c_ssize_t)),
                           ri.ctypes.data_as(C.POINTER(C.c_int32)), None, C.byref(stt))
    assert m is None and stt.value == -4


def test_python_mirror_host_paths(lib, tmp_path, automotive):
    import scipy.sparse as sp

    from slim_b200 import SLIM, SLIMatrix
    from slim_b200.core import make_options

    g = automotive
    R = sp.csr_matrix((g["trn_rowval"], g["trn_rowind"], g["trn_rowptr"]), shape=(2928, 1835))
    mat = SLIMatrix(R)
    assert (mat.nUsers, mat.nItems) == (2928, 1835)
    # a model assembled from the golden W, driven through the mirror's predict / save / load / to_csr
    model = SLIM()
    model.handle = C.c_void_p(_assemble(lib, g))
    model.ismodel, model.nItems = 1, 1835
    model.id2item = np.arange(1835)
    model.item2id = model.id2item
    out, scores = model.predict(mat, nrcmds=10, returnscores=True)
    got = np.stack([out[u] for u in range(2928)])
    assert np.array_equal(got, g["top10_ids"])
    W = model.to_csr()
    assert W.nnz == 84317 and W.shape == (1835, 1835)
    model.save_model(str(tmp_path / "m.csr"), str(tmp_path / "m.map"))
    m2 = SLIM()
    m2.load_model(str(tmp_path / "m.csr"), str(tmp_path / "m.map"))
    assert m2.to_csr().nnz == 84317
    # triplet input with id remapping
    trip = [("u1", "a", 5.0), ("u1", "b", 3.0), ("u2", "b", 1.0)]
    t = SLIMatrix(trip)
    assert (t.nUsers, t.nItems) == (2, 2) and t.item2id == {"a": 0, "b": 1}
    with pytest.raises(TypeError):
        make_options({"l1r": -1.0})
    io, do = make_options({"niters": 7, "l1r": 2})
    assert io[st.OPT_MAXNITERS] == 7 and do[st.OPT_L1R] == 2.0 and do[st.OPT_OPTTOL] == 1e-7
