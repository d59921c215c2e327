"""Shared helpers for the test-suite, bench.py's CPU legs and tests/golden/make_golden.py.

Checker-side code only: this module is what loads ``oracle/`` (the CPU restatement and, when
present, the compiled reference in ``oracle/_ref``).  Nothing under ``slim_b200/`` imports it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"
GOLDEN_DIR = ROOT / "tests" / "golden"

# slim.h option slots (reference include/slim.h:215-230)
OPT_DBGLVL, OPT_NNBRS, OPT_SIMTYPE, OPT_NTHREADS, OPT_MAXNITERS = 0, 1, 2, 3, 4
OPT_ALGO, OPT_ORDERED, OPT_L1R, OPT_L2R, OPT_OPTTOL, OPT_NRCMDS = 5, 6, 7, 8, 9, 10
NOPTIONS = 40
SLIM_OK = 1


class GkCsr(C.Structure):
    """Layout of the model / matrix handle (reference lib/GKlib/gk_struct.h:75-88), 184 bytes."""

    _fields_ = [("nrows", C.c_int32), ("ncols", C.c_int32)] + [
        (n, C.c_void_p)
        for n in (
            "rowptr colptr rowind colind rowids colids rlabels clabels rmap cmap "
            "rowval colval rnorms cnorms rsums csums rsizes csizes rvols cvols rwgts cwgts"
        ).split()
    ]


assert C.sizeof(GkCsr) == 184


def _arr(ptr, n, dtype):
    if not ptr or n == 0:
        return np.zeros(0, dtype=dtype)
    ct = {np.int64: C.c_int64, np.int32: C.c_int32, np.float32: C.c_float}[dtype]
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(n,)).copy()


def model_views(handle):
    """Copy both views of a gk_csr_t-layout model handle into numpy arrays."""
    m = C.cast(handle, C.POINTER(GkCsr)).contents
    out = {"nrows": m.nrows, "ncols": m.ncols}
    if m.colptr:
        cp = _arr(m.colptr, m.ncols + 1, np.int64)
        out.update(colptr=cp, colind=_arr(m.colind, int(cp[-1]), np.int32),
                   colval=_arr(m.colval, int(cp[-1]), np.float32))
    if m.rowptr:
        rp = _arr(m.rowptr, m.nrows + 1, np.int64)
        out.update(rowptr=rp, rowind=_arr(m.rowind, int(rp[-1]), np.int32),
                   rowval=_arr(m.rowval, int(rp[-1]), np.float32))
    return out


def options(l1r=1.0, l2r=1.0, opttol=None, niters=None, nthreads=None, dbglvl=0, nnbrs=None, simtype=None):
    io = np.full(NOPTIONS, -1, dtype=np.int32)
    do = np.full(NOPTIONS, -1.0, dtype=np.float64)
    io[OPT_DBGLVL] = dbglvl
    if nnbrs is not None:
        io[OPT_NNBRS] = nnbrs
    if simtype is not None:
        io[OPT_SIMTYPE] = SIMTYPES[simtype]
    if nthreads is not None:
        io[OPT_NTHREADS] = nthreads
    if niters is not None:
        io[OPT_MAXNITERS] = niters
    do[OPT_L1R], do[OPT_L2R] = l1r, l2r
    if opttol is not None:
        do[OPT_OPTTOL] = opttol
    return io, do


def _p(a, ct):
    return a.ctypes.data_as(C.POINTER(ct)) if a is not None else None


class SlimLib:
    """ctypes binding of a libslim.so-compatible library (the reference build in oracle/_ref or
    our own libslim.so) -- the same calls the reference python-package makes."""

    def __init__(self, path):
        self.path = str(path)
        self.lib = C.CDLL(self.path)
        L = self.lib
        L.SLIM_Learn.restype = C.c_void_p
        L.SLIM_Learn.argtypes = [C.c_int32, C.POINTER(C.c_ssize_t), C.POINTER(C.c_int32),
                                 C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_double),
                                 C.c_void_p, C.POINTER(C.c_int32)]
        L.SLIM_GetTopN.restype = C.c_int32
        L.SLIM_GetTopN.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_float),
                                   C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_int32),
                                   C.POINTER(C.c_float)]
        L.SLIM_FreeModel.restype = None
        L.SLIM_FreeModel.argtypes = [C.POINTER(C.c_void_p)]
        L.SLIM_DetermineHeadAndTail.restype = C.POINTER(C.c_int32)
        L.SLIM_DetermineHeadAndTail.argtypes = [C.c_int32, C.c_int32, C.POINTER(C.c_ssize_t),
                                                C.POINTER(C.c_int32)]

    def head_tail(self, nrows, ncols, rowptr, rowind):
        rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
        rowind = np.ascontiguousarray(rowind, dtype=np.int32)
        p = self.lib.SLIM_DetermineHeadAndTail(nrows, ncols, _p(rowptr, C.c_ssize_t),
                                               _p(rowind, C.c_int32))
        fm = np.ctypeslib.as_array(p, shape=(ncols,)).copy()
        C.CDLL("libc.so.6").free(p)
        return fm

    def learn(self, rowptr, rowind, rowval, io, do, imodel=None):
        rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
        rowind = np.ascontiguousarray(rowind, dtype=np.int32)
        rv = None if rowval is None else np.ascontiguousarray(rowval, dtype=np.float32)
        st = C.c_int32(0)
        h = self.lib.SLIM_Learn(len(rowptr) - 1, _p(rowptr, C.c_ssize_t), _p(rowind, C.c_int32),
                                _p(rv, C.c_float), _p(io, C.c_int32), _p(do, C.c_double),
                                imodel, C.byref(st))
        return h, st.value

    def free(self, handle):
        hp = C.c_void_p(handle)
        self.lib.SLIM_FreeModel(C.byref(hp))
        return hp.value

    def topn_all(self, handle, rowptr, rowind, rowval, n=10):
        """Top-n ids/scores for every user (row of the history CSR); -1 pads short lists."""
        rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
        rowind = np.ascontiguousarray(rowind, dtype=np.int32)
        rv = None if rowval is None else np.ascontiguousarray(rowval, dtype=np.float32)
        nu = len(rowptr) - 1
        ids = np.full((nu, n), -1, dtype=np.int32)
        sc = np.zeros((nu, n), dtype=np.float32)
        io = np.full(NOPTIONS, -1, dtype=np.int32)
        for u in range(nu):
            a, b = int(rowptr[u]), int(rowptr[u + 1])
            it = rowind[a:b]
            self.lib.SLIM_GetTopN(
                handle, b - a, it.ctypes.data_as(C.POINTER(C.c_int32)),
                None if rv is None else rv[a:b].ctypes.data_as(C.POINTER(C.c_float)),
                _p(io, C.c_int32), n, ids[u].ctypes.data_as(C.POINTER(C.c_int32)),
                sc[u].ctypes.data_as(C.POINTER(C.c_float)))
        return ids, sc


def capture_stdout(fn):
    """Run fn() with the process-level stdout (C printf included) redirected; returns (result, text)."""
    import sys
    import tempfile

    sys.stdout.flush()
    libc = C.CDLL("libc.so.6")
    libc.fflush(None)
    saved = os.dup(1)
    with tempfile.TemporaryFile("w+b") as tf:
        os.dup2(tf.fileno(), 1)
        try:
            out = fn()
        finally:
            libc.fflush(None)
            os.dup2(saved, 1)
            os.close(saved)
        tf.seek(0)
        return out, tf.read().decode(errors="replace")


def mselect(lib, trn, tst, l1s, l2s, nrcmds=10, **opt):
    """Py_SLIM_Mselect of a libslim.so-compatible library (reference pyapi.c:214-412) on CSR triples.
    Returns (rc, best[8] = l1HR, l2HR, HRHR, ARHR, l1AR, l2AR, HRAR, ARAR, per-cell rows parsed from the
    library's own printout: l1r, l2r, nnz, hr, hr_head, hr_tail, arhr)."""
    import re

    L = lib.lib
    L.Py_csr_wrapper.restype = C.c_int32
    L.Py_csr_wrapper.argtypes = [C.c_int32, C.POINTER(C.c_ssize_t), C.POINTER(C.c_int32), C.POINTER(C.c_float),
                                 C.POINTER(C.c_void_p)]
    L.Py_csr_free.argtypes = [C.c_void_p]
    L.Py_SLIM_Mselect.restype = C.c_int32
    L.Py_SLIM_Mselect.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_double),
                                  C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int32, C.c_int32] + \
        [C.POINTER(C.c_double)] * 8
    hs = []
    keep = []
    for rp, ri, rv in (trn, tst):
        rp = np.ascontiguousarray(rp, np.int64)
        ri = np.ascontiguousarray(ri, np.int32)
        rv = np.ascontiguousarray(rv, np.float32)
        keep.append((rp, ri, rv))
        h = C.c_void_p()
        assert L.Py_csr_wrapper(len(rp) - 1, _p(rp, C.c_ssize_t), _p(ri, C.c_int32), _p(rv, C.c_float),
                                C.byref(h)) == SLIM_OK
        hs.append(h)
    io, do = options(**opt)
    io[OPT_NRCMDS] = nrcmds
    a1 = np.ascontiguousarray(l1s, np.float64)
    a2 = np.ascontiguousarray(l2s, np.float64)
    best = [C.c_double(0.0) for _ in range(8)]
    rc, text = capture_stdout(lambda: L.Py_SLIM_Mselect(hs[0], hs[1], _p(io, C.c_int32), _p(do, C.c_double),
                                                        _p(a1, C.c_double), _p(a2, C.c_double), len(a1), len(a2),
                                                        *[C.byref(b) for b in best]))
    for h in hs:
        L.Py_csr_free(h)
    pat = re.compile(r"l1r:\s*(\S+)\s+l2r:\s*(\S+)\s+nnz:\s*(\d+)\s+hr:\s*(\S+)\s+hr_head:\s*(\S+)\s+hr_tail:\s*(\S+)"
                     r"\s+arhr:\s*(\S+)")
    cells = np.array([[float(x) for x in m.groups()] for m in pat.finditer(text)], dtype=np.float64)
    return rc, np.array([b.value for b in best]), cells


# ----------------------------------------------------------------------------------------------
# oracle/ loaders
# ----------------------------------------------------------------------------------------------
def build_oracle(ref=True):
    """Compile oracle/ (port always; oracle/_ref only when /root/reference exists)."""
    tgt = ["all"] if ref else ["port"]
    subprocess.run(["make", "-C", str(ORACLE_DIR), "-s"] + tgt, check=True)


def ref_lib_path(cols=False):
    return ORACLE_DIR / "_ref" / ("libslim_ref_cols.so" if cols else "libslim_ref.so")


def have_ref():
    return ref_lib_path().exists()


def load_ref(cols=False):
    return SlimLib(ref_lib_path(cols))


def libc_srand(seed=1):
    """Reset glibc rand() (the reference's ShuffleList uses the never-seeded global stream)."""
    C.CDLL("libc.so.6").srand(C.c_uint(seed))


class _OParams(C.Structure):
    _fields_ = [("l1r", C.c_double), ("l2r", C.c_double), ("optTol", C.c_double),
                ("maxniters", C.c_int32), ("order", C.c_int32), ("nthreads", C.c_int32),
                ("nnbrs", C.c_int32), ("simtype", C.c_int32), ("nbr_ties", C.c_int32)]


class _OCsc(C.Structure):
    _fields_ = [("nrows", C.c_int32), ("ncols", C.c_int32), ("colptr", C.POINTER(C.c_int64)),
                ("colind", C.POINTER(C.c_int32)), ("colval", C.POINTER(C.c_float)),
                ("cnorms", C.POINTER(C.c_float)), ("rowptr", C.POINTER(C.c_int64)),
                ("rowind", C.POINTER(C.c_int32)), ("rowval", C.POINTER(C.c_float))]


class _OStats(C.Structure):
    _fields_ = [("niters", C.POINTER(C.c_int32)), ("nactive", C.POINTER(C.c_int32)),
                ("active_nnz", C.POINTER(C.c_int64)), ("expand_nnz", C.POINTER(C.c_int64)),
                ("rnorm", C.POINTER(C.c_double)), ("objval", C.POINTER(C.c_double))]


ORDER_ASCENDING, ORDER_REF_RAND, ORDER_POPULARITY = 0, 1, 2
TIES_REFERENCE, TIES_POPULARITY = 0, 1
SIMTYPES = {"cos": 0, "jac": 1, "dotp": 2}


class Oracle:
    """ctypes binding of oracle/libslim_oracle.so (the plain-C restatement)."""

    def __init__(self):
        path = ORACLE_DIR / "libslim_oracle.so"
        if not path.exists():
            build_oracle(ref=False)
        self.lib = L = C.CDLL(str(path))
        L.oracle_setup.restype = C.POINTER(_OCsc)
        L.oracle_setup.argtypes = [C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int32),
                                   C.POINTER(C.c_float)]
        L.oracle_free_csc.argtypes = [C.POINTER(_OCsc)]
        L.oracle_learn.restype = C.c_int
        L.oracle_learn.argtypes = [C.POINTER(_OCsc), C.POINTER(_OParams), C.POINTER(C.c_int32),
                                   C.c_int32, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int32),
                                   C.POINTER(C.c_float), C.POINTER(C.POINTER(C.c_int64)),
                                   C.POINTER(C.POINTER(C.c_int32)), C.POINTER(C.POINTER(C.c_float)),
                                   C.POINTER(_OStats)]
        L.oracle_free.argtypes = [C.c_void_p]
        L.oracle_topn.restype = C.c_int32
        L.oracle_topn.argtypes = [C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int32),
                                  C.POINTER(C.c_float), C.c_int32, C.POINTER(C.c_int32),
                                  C.POINTER(C.c_float), C.c_int32, C.POINTER(C.c_int32),
                                  C.POINTER(C.c_float)]
        L.oracle_transpose.argtypes = [C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int32),
                                       C.POINTER(C.c_float), C.POINTER(C.c_int64),
                                       C.POINTER(C.c_int32), C.POINTER(C.c_float)]

    def setup(self, rowptr, rowind, rowval):
        rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
        rowind = np.ascontiguousarray(rowind, dtype=np.int32)
        rv = None if rowval is None else np.ascontiguousarray(rowval, dtype=np.float32)
        m = self.lib.oracle_setup(len(rowptr) - 1, _p(rowptr, C.c_int64), _p(rowind, C.c_int32),
                                  _p(rv, C.c_float))
        return m

    def csc_arrays(self, m):
        mm = m.contents
        nnz = int(mm.colptr[mm.ncols]) if mm.ncols > 0 else 0
        cp = np.ctypeslib.as_array(mm.colptr, shape=(mm.ncols + 1,)).copy()
        ci = np.ctypeslib.as_array(mm.colind, shape=(max(nnz, 1),))[:nnz].copy()
        cv = None
        if mm.colval:
            cv = np.ctypeslib.as_array(mm.colval, shape=(max(nnz, 1),))[:nnz].copy()
        cn = np.ctypeslib.as_array(mm.cnorms, shape=(max(mm.ncols, 1),))[:mm.ncols].copy()
        return dict(nrows=mm.nrows, ncols=mm.ncols, colptr=cp, colind=ci, colval=cv, cnorms=cn)

    def free_csc(self, m):
        self.lib.oracle_free_csc(m)

    def learn(self, rowptr, rowind, rowval, l1r=1.0, l2r=1.0, opttol=1e-7, niters=10000,
              order=ORDER_ASCENDING, nthreads=1, cols=None, imodel=None, want_stats=False,
              nnbrs=0, simtype="cos", nbr_ties=TIES_REFERENCE):
        """Returns dict(colptr, colind, colval[, stats]) for the solved columns (in `cols` order).
        imodel: optional (ncols, colptr, colind, colval) CSC of a warm-start model."""
        m = self.setup(rowptr, rowind, rowval)
        try:
            ncols = m.contents.ncols
            p = _OParams(l1r, l2r, opttol, niters, order, nthreads, nnbrs, SIMTYPES[simtype], nbr_ties)
            cs = None if cols is None else np.ascontiguousarray(cols, dtype=np.int32)
            nsel = ncols if cs is None else len(cs)
            st = None
            sarr = {}
            if want_stats:
                sarr = dict(niters=np.zeros(nsel, np.int32), nactive=np.zeros(nsel, np.int32),
                            active_nnz=np.zeros(nsel, np.int64), expand_nnz=np.zeros(nsel, np.int64),
                            rnorm=np.zeros(nsel, np.float64), objval=np.zeros(nsel, np.float64))
                st = _OStats(_p(sarr["niters"], C.c_int32), _p(sarr["nactive"], C.c_int32),
                             _p(sarr["active_nnz"], C.c_int64), _p(sarr["expand_nnz"], C.c_int64),
                             _p(sarr["rnorm"], C.c_double), _p(sarr["objval"], C.c_double))
            inc, icp, ici, icv = 0, None, None, None
            if imodel is not None:
                inc = int(imodel[0])
                icp = np.ascontiguousarray(imodel[1], dtype=np.int64)
                ici = np.ascontiguousarray(imodel[2], dtype=np.int32)
                icv = np.ascontiguousarray(imodel[3], dtype=np.float32)
            wp, wi, wv = C.POINTER(C.c_int64)(), C.POINTER(C.c_int32)(), C.POINTER(C.c_float)()
            rc = self.lib.oracle_learn(m, C.byref(p), _p(cs, C.c_int32), nsel, inc,
                                       _p(icp, C.c_int64), _p(ici, C.c_int32), _p(icv, C.c_float),
                                       C.byref(wp), C.byref(wi), C.byref(wv),
                                       C.byref(st) if st is not None else None)
            assert rc == 0
            colptr = np.ctypeslib.as_array(wp, shape=(nsel + 1,)).copy()
            nnz = int(colptr[-1])
            colind = np.ctypeslib.as_array(wi, shape=(max(nnz, 1),))[:nnz].copy()
            colval = np.ctypeslib.as_array(wv, shape=(max(nnz, 1),))[:nnz].copy()
            for q in (wp, wi, wv):
                self.lib.oracle_free(C.cast(q, C.c_void_p))
            out = dict(ncols=ncols, colptr=colptr, colind=colind, colval=colval)
            if want_stats:
                out["stats"] = sarr
            return out
        finally:
            self.free_csc(m)

    def transpose(self, n, ptr, ind, val):
        ptr = np.ascontiguousarray(ptr, dtype=np.int64)
        ind = np.ascontiguousarray(ind, dtype=np.int32)
        val = np.ascontiguousarray(val, dtype=np.float32)
        tp = np.zeros(n + 1, np.int64)
        ti = np.zeros(max(len(ind), 1), np.int32)
        tv = np.zeros(max(len(ind), 1), np.float32)
        self.lib.oracle_transpose(n, _p(ptr, C.c_int64), _p(ind, C.c_int32), _p(val, C.c_float),
                                  _p(tp, C.c_int64), _p(ti, C.c_int32), _p(tv, C.c_float))
        return tp, ti[:len(ind)], tv[:len(ind)]

    def topn_all(self, nitems, wrowptr, wrowind, wrowval, rowptr, rowind, rowval, n=10):
        wrowptr = np.ascontiguousarray(wrowptr, dtype=np.int64)
        wrowind = np.ascontiguousarray(wrowind, dtype=np.int32)
        wrowval = np.ascontiguousarray(wrowval, dtype=np.float32)
        rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
        rowind = np.ascontiguousarray(rowind, dtype=np.int32)
        rv = None if rowval is None else np.ascontiguousarray(rowval, dtype=np.float32)
        nu = len(rowptr) - 1
        ids = np.full((nu, n), -1, dtype=np.int32)
        sc = np.zeros((nu, n), dtype=np.float32)
        for u in range(nu):
            a, b = int(rowptr[u]), int(rowptr[u + 1])
            self.lib.oracle_topn(
                nitems, _p(wrowptr, C.c_int64), _p(wrowind, C.c_int32), _p(wrowval, C.c_float),
                b - a, rowind[a:b].ctypes.data_as(C.POINTER(C.c_int32)),
                None if rv is None else rv[a:b].ctypes.data_as(C.POINTER(C.c_float)),
                n, ids[u].ctypes.data_as(C.POINTER(C.c_int32)),
                sc[u].ctypes.data_as(C.POINTER(C.c_float)))
        return ids, sc


# ----------------------------------------------------------------------------------------------
# fixtures / file formats / evaluation
# ----------------------------------------------------------------------------------------------
def read_text_csr(path):
    """Text CSR as the reference CLI reads it (lib/GKlib/csr.c:655-769 with numbering=0):
    one row per line, `col val` pairs, 0-based column ids."""
    rowptr, rowind, rowval = [0], [], []
    with open(path) as f:
        for line in f:
            t = line.split()
            rowind.extend(int(x) for x in t[0::2])
            rowval.extend(float(x) for x in t[1::2])
            rowptr.append(len(rowind))
    return (np.asarray(rowptr, np.int64), np.asarray(rowind, np.int32),
            np.asarray(rowval, np.float32))


def read_ijv(path, nrows=None):
    """IJV triplets `row col val`, 0-based (lib/GKlib/csr.c:487-548)."""
    d = np.loadtxt(path, dtype=np.float64, ndmin=2)
    r, c, v = d[:, 0].astype(np.int64), d[:, 1].astype(np.int32), d[:, 2].astype(np.float32)
    n = int(r.max()) + 1 if nrows is None else nrows
    order = np.argsort(r, kind="stable")
    r, c, v = r[order], c[order], v[order]
    rowptr = np.zeros(n + 1, np.int64)
    np.add.at(rowptr, r + 1, 1)
    return np.cumsum(rowptr), c, v


def head_tail(nrows, ncols, rowptr, rowind):
    """fmarker restating SLIM_DetermineHeadAndTail (src/libslim/api.c:215-245): 0 = head (most
    frequent items holding half of the ratings), 1 = tail.  Ties in the frequency sort are
    implementation-defined in the reference; broken here by ascending item id."""
    cnt = np.bincount(rowind, minlength=ncols).astype(np.int64)
    order = np.lexsort((np.arange(ncols), -cnt))
    fm = np.ones(ncols, np.int32)
    left = int(rowptr[nrows]) // 2
    for c in order:
        if left <= 0:
            break
        fm[c] = 0
        left -= int(cnt[c])
    return fm


def evaluate(ids, trn, tst, ncols_model, fmarker=None):
    """HR / head HR / tail HR / ARHR as printed by the reference CLI
    (src/programs/slim_predict.c:181-242).  fmarker: head/tail marks (0/1) per item; computed
    with head_tail() when absent."""
    trp, tri = trn[0], trn[1]
    tsp, tsi = tst[0], tst[1]
    nu = len(trp) - 1
    ncols = max(int(tri.max()) + 1, int(tsi.max()) + 1, ncols_model)
    fm = head_tail(nu, ncols, trp, tri) if fmarker is None else fmarker
    hr = [0.0, 0.0, 0.0]
    arhr = 0.0
    nvalid = nhead = ntail = 0
    for u in range(nu):
        t = tsi[tsp[u]:tsp[u + 1]]
        nvalid += 1
        if len(t) == 0:
            continue
        tset = set(int(x) for x in t)
        ntrue = [0, 0]
        for x in t:
            ntrue[fm[x]] += 1
        nhead += 1 if ntrue[0] else 0
        ntail += 1 if ntrue[1] else 0
        base = sum(1.0 / (1.0 + k) for k in range(len(t)))
        nh = [0, 0, 0]
        l = 0.0
        for r, it in enumerate(ids[u]):
            if it >= 0 and int(it) in tset:
                nh[fm[it]] += 1
                nh[2] += 1
                l += 1.0 / (1.0 + r)
        hr[0] += nh[0] / ntrue[0] if nh[0] > 0 else 0.0
        hr[1] += nh[1] / ntrue[1] if nh[1] > 0 else 0.0
        hr[2] += nh[2] / len(t)
        arhr += l / base
    return dict(hr=hr[2] / max(nvalid, 1), hr_head=hr[0] / max(nhead, 1),
                hr_tail=hr[1] / max(ntail, 1), arhr=arhr / max(nvalid, 1))


def load_golden(name):
    z = np.load(GOLDEN_DIR / f"{name}.npz")
    return {k: z[k] for k in z.files}


def csc_to_dense(ncols, colptr, colind, colval, nrows=None):
    n = ncols if nrows is None else nrows
    W = np.zeros((n, len(colptr) - 1), dtype=np.float64)
    for j in range(len(colptr) - 1):
        a, b = int(colptr[j]), int(colptr[j + 1])
        W[colind[a:b], j] = colval[a:b]
    return W


def compare_models(a, b):
    """max |dW| over the union support and the list of support mismatches with their magnitude.
    a, b: dicts with colptr/colind/colval for the same columns."""
    assert len(a["colptr"]) == len(b["colptr"])
    maxd, flips = 0.0, []
    for j in range(len(a["colptr"]) - 1):
        a0, a1 = int(a["colptr"][j]), int(a["colptr"][j + 1])
        b0, b1 = int(b["colptr"][j]), int(b["colptr"][j + 1])
        ia, va = a["colind"][a0:a1], a["colval"][a0:a1].astype(np.float64)
        ib, vb = b["colind"][b0:b1], b["colval"][b0:b1].astype(np.float64)
        if len(ia) == len(ib) and np.array_equal(ia, ib):
            if len(ia):
                maxd = max(maxd, float(np.max(np.abs(va - vb))))
            continue
        da = dict(zip(ia.tolist(), va.tolist()))
        db = dict(zip(ib.tolist(), vb.tolist()))
        for k in set(da) | set(db):
            d = abs(da.get(k, 0.0) - db.get(k, 0.0))
            maxd = max(maxd, d)
            if (k in da) != (k in db):
                flips.append((j, k, max(abs(da.get(k, 0.0)), abs(db.get(k, 0.0)))))
    return maxd, flips


def synth_zipf(nusers, nitems, per_user, seed=42, alpha=1.1, ratings=False, perm_seed=43):
    """SURVEY.md section 8d synthetic R: every user draws `per_user` distinct items from
    Zipf(alpha) without replacement (Gumbel top-k), item ids permuted, columns sorted per row.
    Returns (rowptr int64, rowind int32, rowval float32)."""
    rng = np.random.default_rng(seed)
    logp = -alpha * np.log(np.arange(1, nitems + 1, dtype=np.float64))
    perm = np.random.default_rng(perm_seed).permutation(nitems).astype(np.int32)
    rowind = np.empty((nusers, per_user), dtype=np.int32)
    chunk = max(1, min(nusers, (1 << 24) // max(nitems, 1)))
    for s in range(0, nusers, chunk):
        e = min(nusers, s + chunk)
        g = rng.gumbel(size=(e - s, nitems)) + logp
        top = np.argpartition(-g, per_user - 1, axis=1)[:, :per_user]
        rowind[s:e] = np.sort(perm[top], axis=1)
    rowptr = np.arange(0, (nusers + 1) * per_user, per_user, dtype=np.int64)
    if ratings:
        rowval = rng.integers(1, 6, size=nusers * per_user).astype(np.float32)
    else:
        rowval = np.ones(nusers * per_user, dtype=np.float32)
    return rowptr, rowind.reshape(-1), rowval
