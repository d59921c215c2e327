"""Host-side mirror of the reference python wrapper (python-package/SLIM/core.py) over libslim.so.

Same public names and argument meaning as the reference package -- ``SLIMatrix`` (core.py:245-385)
and ``SLIM`` with ``train / mselect / predict / save_model / load_model / to_csr`` (core.py:388-804)
-- so code written against ``from SLIM import SLIM, SLIMatrix`` runs against the CUDA engine by
changing the import.  (The unmodified reference package also works: it only needs this repo's
libslim.so in site-packages/SLIM/, see INTEGRATION.md.)

Extensions without a reference counterpart: ``Staged`` keeps R resident in HBM between learn calls
and ``learn_columns`` solves a subset of target columns (the unit the multi-GPU driver shards).
"""
from __future__ import annotations

import ctypes as C
import os
import time

import numpy as np

from . import _lib

SLIM_NOPTIONS = 40
SLIM_OK = 1
(OPT_DBGLVL, OPT_NNBRS, OPT_SIMTYPE, OPT_NTHREADS, OPT_MAXNITERS, OPT_ALGO, OPT_ORDERED, OPT_L1R,
 OPT_L2R, OPT_OPTTOL, OPT_NRCMDS) = range(11)
_SIMTYPES = {"cos": 0, "jac": 1, "dotp": 2}
_ALGOS = {"admm": 0, "cd": 1}

# python-level defaults of the reference wrapper (core.py:123-198); note niters=50, not 10000
_DEFAULTS = dict(dbglvl=0, nnbrs=0, simtype="cos", algo="cd", nthreads=1, niters=50, nrcmds=10,
                 l1r=1.0, l2r=1.0, optTol=1e-7)


def _ptr(a, ct):
    return None if a is None else a.ctypes.data_as(C.POINTER(ct))


def make_options(params=None, **kw):
    """dict / attribute-object / keywords -> (ioptions int32[40], doptions float64[40]) with the
    validation of check_dict_params (reference core.py:123-198)."""
    p = dict(_DEFAULTS)
    if params is not None:
        src = params if isinstance(params, dict) else {k: getattr(params, k) for k in _DEFAULTS
                                                       if hasattr(params, k)}
        p.update(src)
    p.update(kw)
    for key in ("dbglvl", "nnbrs", "niters", "nrcmds"):
        if not isinstance(p[key], (int, np.integer)) or p[key] < 0:
            raise TypeError(f"Please provide a non-negative integer value for {key}.")
    if not isinstance(p["nthreads"], (int, np.integer)) or p["nthreads"] <= 0:
        raise TypeError("Please provide positive integer value for nthreads.")
    if p["simtype"] not in _SIMTYPES:
        raise TypeError("Please select simtype from {'cos', 'jac', 'dotp'}.")
    if p["algo"] not in _ALGOS:
        raise TypeError("Please select algo from {'admm', 'cd'}.")
    for key in ("l1r", "l2r", "optTol"):
        if not isinstance(p[key], (int, float, np.floating, np.integer)) or p[key] < 0:
            raise TypeError(f"Please provide non-negative value for {key}.")
    io = np.full(SLIM_NOPTIONS, -1, dtype=np.int32)
    do = np.full(SLIM_NOPTIONS, -1.0, dtype=np.float64)
    io[OPT_DBGLVL], io[OPT_NNBRS] = p["dbglvl"], p["nnbrs"]
    io[OPT_SIMTYPE], io[OPT_ALGO] = _SIMTYPES[p["simtype"]], _ALGOS[p["algo"]]
    io[OPT_NTHREADS], io[OPT_ORDERED] = p["nthreads"], 0
    io[OPT_MAXNITERS], io[OPT_NRCMDS] = p["niters"], p["nrcmds"]
    do[OPT_L1R], do[OPT_L2R], do[OPT_OPTTOL] = p["l1r"], p["l2r"], p["optTol"]
    return io, do


class SLIMatrix:
    """User x item matrix handed to the library (reference core.py:245-385).  `data` is a scipy
    csr_matrix, or ijv triplets (list / ndarray / DataFrame) whose user and item keys are mapped to
    dense ids; `oldmat` (a SLIMatrix or SLIM) supplies an existing mapping."""

    def __init__(self, data, oldmat=None):
        import scipy.sparse as sp

        self._lib = _lib.load()
        self.handle = C.c_void_p()
        if sp.issparse(data):
            R = sp.csr_matrix(data)
            self.nUsers, self.nItems = R.shape
            if oldmat is not None and isinstance(oldmat, SLIMatrix) and \
                    (self.nUsers, self.nItems) != (oldmat.nUsers, oldmat.nItems):
                raise TypeError("The size of the input matrix does not match the size of oldmat.")
            if oldmat is not None and isinstance(oldmat, SLIM) and self.nItems != len(oldmat.id2item):
                raise TypeError("The size of the input matrix does not match the size of oldmat.")
            self.id2item = np.arange(self.nItems)
            self.item2id = self.id2item
            self.id2user = np.arange(self.nUsers)
            self.user2id = self.id2user
        else:
            if hasattr(data, "values"):
                data = data.values
            R = self._from_triplets(list(data), oldmat)
        self._wrap(R)

    def _from_triplets(self, data, oldmat):
        import scipy.sparse as sp

        if oldmat is not None:
            if not isinstance(oldmat, (SLIMatrix, SLIM)):
                raise AssertionError("Please feed in a SLIMatrix object or a SLIM model for oldmat.")
            self.id2item = np.array(oldmat.id2item).copy()
            self.item2id = (dict(oldmat.item2id) if isinstance(oldmat.item2id, dict)
                            else {k: i for i, k in enumerate(self.id2item.tolist())})
        else:
            self.item2id, self.id2item = {}, []
        if isinstance(oldmat, SLIMatrix):
            self.id2user = np.array(oldmat.id2user).copy()
            self.user2id = (dict(oldmat.user2id) if isinstance(oldmat.user2id, dict)
                            else {k: i for i, k in enumerate(self.id2user.tolist())})
            grow_users = False
        else:
            self.user2id, self.id2user = {}, []
            grow_users = True
        grow_items = oldmat is None
        row, col, val, miss = [], [], [], 0
        for u, i, v in data:
            if grow_users and u not in self.user2id:
                self.user2id[u] = len(self.id2user)
                self.id2user.append(u)
            if grow_items and i not in self.item2id:
                self.item2id[i] = len(self.id2item)
                self.id2item.append(i)
            if u in self.user2id and i in self.item2id:
                row.append(self.user2id[u])
                col.append(self.item2id[i])
                val.append(v)
            else:
                miss += 1
        if miss:
            print("%d of the events fall out of the range of oldmat. Partial entries collected." % miss)
        self.id2item = np.array(self.id2item)
        self.id2user = np.array(self.id2user)
        self.nUsers, self.nItems = len(self.id2user), len(self.id2item)
        return sp.csr_matrix((val, (row, col)), shape=(self.nUsers, self.nItems))

    def _wrap(self, R):
        self.rowptr = np.ascontiguousarray(R.indptr, dtype=np.int64)
        self.rowind = np.ascontiguousarray(R.indices, dtype=np.int32)
        self.rowval = np.ascontiguousarray(R.data, dtype=np.float32)
        rc = self._lib.Py_csr_wrapper(R.shape[0], _ptr(self.rowptr, C.c_ssize_t),
                                      _ptr(self.rowind, C.c_int32), _ptr(self.rowval, C.c_float),
                                      C.byref(self.handle))
        if rc != SLIM_OK:
            raise MemoryError("Py_csr_wrapper failed")

    def __del__(self):
        try:
            if self.handle:
                self._lib.Py_csr_free(self.handle)
                self.handle = C.c_void_p()
        except Exception:
            pass


class SLIM:
    """A SLIM model (reference core.py:388-804)."""

    def __init__(self):
        self._lib = _lib.load()
        self.ismodel = 0
        self.handle = C.c_void_p()

    def _drop(self):
        if self.ismodel == SLIM_OK and self.handle:
            self._lib.Py_csr_free(self.handle)
        self.handle = C.c_void_p()
        self.ismodel = 0

    def __del__(self):
        try:
            self._drop()
        except Exception:
            pass

    def train(self, params, data):
        assert type(data) == SLIMatrix, "trndata must be a SLIMatrix object."
        io, do = make_options(params)
        self._drop()
        self.nItems = data.nItems
        start = time.time()
        rc = self._lib.Py_SLIM_Learn(data.handle, _ptr(io, C.c_int32), _ptr(do, C.c_double),
                                     C.byref(self.handle))
        self.ismodel = rc
        self.id2item = np.array(data.id2item).copy()
        self.item2id = data.item2id.copy() if hasattr(data.item2id, "copy") else data.item2id
        if rc != SLIM_OK:
            raise RuntimeError("Something went wrong with model estimation: " +
                               (self._lib.SLIMB200_LastError() or b"").decode())
        print("Learning takes %.3f secs." % (time.time() - start))

    def mselect(self, params, trndata, tstdata, arrayl1, arrayl2, nrcmds):
        assert type(trndata) == SLIMatrix and type(tstdata) == SLIMatrix
        if len(arrayl1) < 1 or len(arrayl2) < 1:
            raise TypeError("The l1 / l2 arrays must not be empty.")
        io, do = make_options(params, nrcmds=nrcmds)
        l1 = np.ascontiguousarray(np.sort(arrayl1), dtype=np.float64)
        l2 = np.ascontiguousarray(np.sort(arrayl2), dtype=np.float64)
        best = [C.c_double(0.0) for _ in range(8)]
        start = time.time()
        rc = self._lib.Py_SLIM_Mselect(trndata.handle, tstdata.handle, _ptr(io, C.c_int32),
                                       _ptr(do, C.c_double), _ptr(l1, C.c_double), _ptr(l2, C.c_double),
                                       len(l1), len(l2), *[C.byref(b) for b in best])
        if rc != SLIM_OK:
            raise RuntimeError("Something went wrong with model estimation or evaluation when "
                               "l1=%.4f, l2=%.4f." % (best[0].value, best[1].value))
        v = [b.value for b in best]
        print("Model selection takes %.3f secs." % (time.time() - start))
        print("The best HR is achieved by, l1: %.4f, l2:%.4f, HR:%.4f, AR:%.4f." % tuple(v[:4]))
        print("The best AR is achieved by, l1: %.4f, l2:%.4f, HR:%.4f, AR:%.4f." % tuple(v[4:]))
        return dict(bestl1HR=v[0], bestl2HR=v[1], bestHRHR=v[2], bestARHR=v[3],
                    bestl1AR=v[4], bestl2AR=v[5], bestHRAR=v[6], bestARAR=v[7])

    def predict(self, data, nrcmds=10, outfile=None, negitems=None, nnegs=0, returnscores=False):
        if self.ismodel != SLIM_OK:
            raise TypeError("Model not found. Please train a model.")
        assert self.nItems == data.nItems, "The shape of the input matrix should match the model."
        res = np.full(data.nUsers * nrcmds, -1, dtype=np.int32)
        scores = np.zeros(data.nUsers * nrcmds, dtype=np.float32)
        if negitems is not None:
            assert nnegs >= nrcmds
            neg = np.full(data.nUsers * nnegs, -1, dtype=np.int32)
            for key, value in negitems.items():
                assert len(value) == nnegs, "The number of negative items should match nnegs."
                for i in range(nnegs):
                    try:
                        neg[data.user2id[key] * nnegs + i] = self.item2id[value[i]]
                    except (KeyError, IndexError):
                        pass
            rc = self._lib.Py_SLIM_Predict_1vsk(nrcmds, nnegs, self.handle, data.handle,
                                                _ptr(neg, C.c_int32), _ptr(res, C.c_int32),
                                                _ptr(scores, C.c_float))
        else:
            rc = self._lib.Py_SLIM_Predict(nrcmds, self.handle, data.handle, _ptr(res, C.c_int32),
                                           _ptr(scores, C.c_float))
        if rc != SLIM_OK:
            raise RuntimeError("Something went wrong during prediction.")
        res = np.asarray(self.id2item)[res].reshape(data.nUsers, nrcmds)
        scores = scores.reshape(data.nUsers, nrcmds)
        items = data.user2id.items() if isinstance(data.user2id, dict) else ((k, k) for k in data.user2id)
        out, outscores = {}, {}
        for key, value in items:
            out[key] = res[value, :]
            outscores[key] = scores[value, :]
        if outfile:
            with open(outfile, "w") as f:
                for key, value in out.items():
                    f.write(str(key) + ": " + np.array2string(value, max_line_width=np.inf) + "\n")
                    if returnscores:
                        f.write(str(key) + ": " + np.array2string(outscores[key], max_line_width=np.inf) + "\n")
        return (out, outscores) if returnscores else out

    def save_model(self, modelfname, mapfname):
        if self.ismodel != SLIM_OK:
            raise RuntimeError("Not exist a model to save.")
        self._lib.Py_csr_save(self.handle, modelfname.encode("utf-8"))
        np.savetxt(mapfname, self.id2item, fmt="%s")

    def load_model(self, modelfname, mapfname):
        if not (os.path.isfile(modelfname) and os.path.isfile(mapfname)):
            raise RuntimeError("File does not exist or invalid filename.")
        self._drop()
        self.ismodel = self._lib.Py_csr_load(C.byref(self.handle), modelfname.encode("utf-8"))
        try:
            self.id2item = np.genfromtxt(mapfname, dtype=np.int32)
        except Exception:
            self.id2item = np.genfromtxt(mapfname)
        self.id2item = np.atleast_1d(self.id2item)
        self.item2id = {k: i for i, k in enumerate(self.id2item.tolist())}
        self.nItems = len(self.id2item)
        if self.ismodel != SLIM_OK:
            raise RuntimeError("Fail to load the model.")

    def to_csr(self, returnmap=False):
        import scipy.sparse as sp

        if self.ismodel != SLIM_OK:
            raise RuntimeError("Not exist a model to export.")
        nnz = C.c_int32(0)
        self._lib.Py_csr_stat(self.handle, C.byref(nnz))
        indptr = np.zeros(self.nItems + 1, dtype=np.int32)
        indices = np.zeros(nnz.value, dtype=np.int32)
        data = np.ones(nnz.value, dtype=np.float32)
        self._lib.Py_csr_export(self.handle, _ptr(indptr, C.c_int32), _ptr(indices, C.c_int32),
                                _ptr(data, C.c_float))
        m = sp.csr_matrix((data, indices, indptr), shape=(self.nItems, self.nItems))
        return (m, self.id2item[:]) if returnmap else m


# ------------------------------------------------------------------------------------------------
# extension: resident matrices and column subsets (include/slim_b200.h)
# ------------------------------------------------------------------------------------------------
class Staged:
    """R staged in HBM (CSR + padded CSC + norms): the result of CreateTrainingMatrix
    (reference src/libslim/setup.c:109-135), kept on the device until close()."""

    def __init__(self, rowptr, rowind, rowval, device=0):
        self._lib = _lib.load()
        st = C.c_int32(0)
        if hasattr(rowptr, "is_cuda") and rowptr.is_cuda:  # torch tensors already in HBM
            import torch

            assert rowptr.dtype == torch.int64 and rowind.dtype == torch.int32
            assert rowval is None or rowval.dtype == torch.float32
            nrows, nnz = rowptr.numel() - 1, rowind.numel()
            self.handle = self._lib.SLIMB200_StageDevice(
                device, nrows, nnz, rowptr.data_ptr(), rowind.data_ptr(),
                None if rowval is None else rowval.data_ptr(), C.byref(st))
        else:
            rp = np.ascontiguousarray(rowptr, dtype=np.int64)
            ri = np.ascontiguousarray(rowind, dtype=np.int32)
            rv = None if rowval is None else np.ascontiguousarray(rowval, dtype=np.float32)
            self.handle = self._lib.SLIMB200_Stage(device, len(rp) - 1, _ptr(rp, C.c_ssize_t),
                                                   _ptr(ri, C.c_int32), _ptr(rv, C.c_float), C.byref(st))
        if not self.handle:
            raise RuntimeError("SLIMB200_Stage failed (%d): %s" %
                               (st.value, (self._lib.SLIMB200_LastError() or b"").decode()))
        nr, nc, dev, nl = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        nnz, ms = C.c_int64(), C.c_double()
        self._lib.SLIMB200_MatrixInfo(self.handle, C.byref(nr), C.byref(nc), C.byref(nnz), C.byref(dev),
                                      C.byref(ms), C.byref(nl))
        self.nrows, self.ncols, self.nnz, self.device = nr.value, nc.value, nnz.value, dev.value
        self.stage_ms, self.stage_launches = ms.value, nl.value

    def csc(self):
        cp = np.zeros(self.ncols + 1, np.int64)
        ci = np.zeros(max(self.nnz, 1), np.int32)
        cv = np.zeros(max(self.nnz, 1), np.float32)
        cn = np.zeros(max(self.ncols, 1), np.float32)
        rc = self._lib.SLIMB200_MatrixCSC(self.handle, _ptr(cp, C.c_int64), _ptr(ci, C.c_int32),
                                          _ptr(cv, C.c_float), _ptr(cn, C.c_float))
        assert rc == SLIM_OK
        return dict(colptr=cp, colind=ci[:self.nnz], colval=cv[:self.nnz], cnorms=cn[:self.ncols])

    def item_order(self):
        """rank[original item id] = internal (popularity-ordered) id."""
        rank = np.zeros(max(self.ncols, 1), np.int32)
        assert self._lib.SLIMB200_MatrixItemOrder(self.handle, _ptr(rank, C.c_int32)) == SLIM_OK
        return rank[:self.ncols]

    def window_gram(self):
        nwin = (self.ncols + 31) // 32
        out = np.zeros((max(nwin, 1), 32, 32), np.float64)
        rc = self._lib.SLIMB200_MatrixWindowGram(self.handle, _ptr(out, C.c_double))
        assert rc == SLIM_OK
        return out[:nwin]

    def gram_info(self):
        """(element bytes, build ms) of the staged Gram matrix R^T R; element bytes 0 = not staged."""
        eb, ms = C.c_int32(0), C.c_double(0.0)
        assert self._lib.SLIMB200_MatrixGramInfo(self.handle, C.byref(eb), C.byref(ms)) == SLIM_OK
        return eb.value, ms.value

    def gram_layout(self):
        """(bytes in HBM, first 16-bit column, first 8-bit column) of the staged Gram matrix (packed layout)."""
        b, h32, h16 = C.c_int64(0), C.c_int32(0), C.c_int32(0)
        assert self._lib.SLIMB200_MatrixGramLayout(self.handle, C.byref(b), C.byref(h32), C.byref(h16)) == SLIM_OK
        return b.value, h32.value, h16.value

    def gram_stair(self):
        """(stair, hd): stair = 1 when G is held in the STAIR layout (panel p stores rows [0, max(64(p+1), hd)) only)."""
        st, hd = C.c_int32(0), C.c_int32(0)
        assert self._lib.SLIMB200_MatrixGramStair(self.handle, C.byref(st), C.byref(hd)) == SLIM_OK
        return st.value, hd.value

    def gram(self):
        """Dense copy of the staged Gram matrix (internal item order), None when it was not staged."""
        eb, _ = self.gram_info()
        if eb == 0:
            return None
        out = np.zeros((self.ncols, self.ncols), np.float32 if eb == 4 else np.float64)
        assert self._lib.SLIMB200_MatrixGram(self.handle, out.ctypes.data_as(C.c_void_p)) == SLIM_OK
        return out

    def close(self):
        if getattr(self, "handle", None):
            h = C.c_void_p(self.handle)
            self._lib.SLIMB200_FreeMatrix(C.byref(h))
            self.handle = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ColumnResult:
    """Solved target columns (device resident until fetched)."""

    def __init__(self, lib, handle):
        self._lib, self.handle = lib, handle
        nsel, nl, nnz = C.c_int32(), C.c_int32(), C.c_int64()
        sms, gms = C.c_double(), C.c_double()
        lib.SLIMB200_ResultInfo(handle, C.byref(nsel), C.byref(nnz), C.byref(sms), C.byref(gms), C.byref(nl))
        self.nsel, self.nnz = nsel.value, nnz.value
        self.solve_ms, self.gather_ms, self.launches = sms.value, gms.value, nl.value

    def to_host(self):
        cp = np.zeros(self.nsel + 1, np.int64)
        ci = np.zeros(max(self.nnz, 1), np.int32)
        cv = np.zeros(max(self.nnz, 1), np.float32)
        rc = self._lib.SLIMB200_ResultToHost(self.handle, _ptr(cp, C.c_int64), _ptr(ci, C.c_int32),
                                             _ptr(cv, C.c_float))
        if rc != SLIM_OK:
            raise RuntimeError("SLIMB200_ResultToHost failed")
        return dict(colptr=cp, colind=ci[:self.nnz], colval=cv[:self.nnz])

    def to_device(self, counts, colind, colval):
        """Copy into caller-owned CUDA tensors (int32[nsel], int32[>=nnz], float32[>=nnz])."""
        rc = self._lib.SLIMB200_ResultToDevice(self.handle, counts.data_ptr(), colind.data_ptr(),
                                               colval.data_ptr())
        if rc != SLIM_OK:
            raise RuntimeError("SLIMB200_ResultToDevice failed")

    def stats(self):
        n = max(self.nsel, 1)
        s = dict(niters=np.zeros(n, np.int32), nactive=np.zeros(n, np.int32),
                 active_nnz=np.zeros(n, np.int64), expand_nnz=np.zeros(n, np.int64),
                 rnorm=np.zeros(n, np.float64), objval=np.zeros(n, np.float64))
        self._lib.SLIMB200_ResultStats(self.handle, _ptr(s["niters"], C.c_int32), _ptr(s["nactive"], C.c_int32),
                                       _ptr(s["active_nnz"], C.c_int64), _ptr(s["expand_nnz"], C.c_int64),
                                       _ptr(s["rnorm"], C.c_double), _ptr(s["objval"], C.c_double))
        return {k: v[:self.nsel] for k, v in s.items()}

    def phases(self):
        ph = np.zeros((max(self.nsel, 1), 4), np.float32)
        ng = np.zeros(max(self.nsel, 1), np.int32)
        self._lib.SLIMB200_ResultPhases(self.handle, _ptr(ph, C.c_float), _ptr(ng, C.c_int32))
        return ph[:self.nsel], ng[:self.nsel]

    def close(self):
        if getattr(self, "handle", None):
            h = C.c_void_p(self.handle)
            self._lib.SLIMB200_FreeResult(C.byref(h))
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def learn_columns(staged: Staged, params=None, cols=None, imodel=None, **kw) -> ColumnResult:
    """Solve target columns `cols` (None: all) of a staged matrix; options as SLIM.train.
    `imodel` is an optional model handle (c_void_p / int) used as warm start."""
    lib = staged._lib
    io, do = make_options(params, **kw)
    cs = None if cols is None else np.ascontiguousarray(cols, dtype=np.int32)
    st = C.c_int32(0)
    h = lib.SLIMB200_LearnColumns(staged.handle, _ptr(io, C.c_int32), _ptr(do, C.c_double),
                                  _ptr(cs, C.c_int32), 0 if cs is None else len(cs), imodel, C.byref(st))
    if not h:
        raise RuntimeError("SLIMB200_LearnColumns failed (%d): %s" %
                           (st.value, (lib.SLIMB200_LastError() or b"").decode()))
    return ColumnResult(lib, h)
