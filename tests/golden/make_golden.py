#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference built into oracle/_ref/.

Run in the build container (needs /root/reference for the data fixtures and the reference
build):   make -C oracle && python tests/golden/make_golden.py

The reference ships no golden vectors for the CD learn path, so the pins
are outputs of the reference library itself (oracle/_ref/libslim_ref.so, built from the sources
where they lie by oracle/Makefile), called in memory through the same C ABI its python-package
uses, with nthreads=1 and glibc rand() reset to its initial state so every run is
bit-reproducible:

  W_conv    : SLIM_Learn at the converged setting (optTol=1e-14, niters=100000) -- the parity
              golden; top-10 lists and HR/ARHR are taken from it with
              the reference's own SLIM_GetTopN.
  W_default : SLIM_Learn at the library defaults (optTol=1e-7, niters=10000) -- pins the
              restatement bit-for-bit in its reference-order mode.

  fslim.npz   : fSLIM models (nnbrs = 10; cos / jac / dotp; converged setting) of both fixtures
                (reference src/libslim/neighbors.c:16-125, estimate.c:424-431).
  mselect.npz : Py_SLIM_Mselect (reference src/libslim/pyapi.c:214-412) on Automotive over a slice of the
                reference's test/l12file grid at the converged setting: per-cell nnz / HR / head / tail / ARHR as
                the library prints them, and the eight best-(l1, l2) outputs.

Each file also carries the training / test matrices (the reference's test/ data fixtures,
re-encoded as CSR arrays) so the GPU box needs nothing from /root/reference.
"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import slimtest as st  # noqa: E402

REF_TEST = Path("/root/reference/test")


def run(name, trn, tst):
    ref = st.load_ref()
    out = dict(trn_rowptr=trn[0], trn_rowind=trn[1], trn_rowval=trn[2],
               tst_rowptr=tst[0], tst_rowind=tst[1], tst_rowval=tst[2])
    for tag, kw in (("conv", dict(opttol=1e-14, niters=100000)), ("default", {})):
        io, do = st.options(l1r=1.0, l2r=1.0, nthreads=1, **kw)
        st.libc_srand(1)
        h, status = ref.learn(trn[0], trn[1], trn[2], io, do)
        assert status == st.SLIM_OK
        mv = st.model_views(h)
        out[f"W_{tag}_colptr"] = mv["colptr"]
        out[f"W_{tag}_colind"] = mv["colind"]
        out[f"W_{tag}_colval"] = mv["colval"]
        if tag == "conv":
            ids, sc = ref.topn_all(h, trn[0], trn[1], trn[2], 10)
            out["top10_ids"], out["top10_scores"] = ids, sc
            nc = max(int(trn[1].max()) + 1, int(tst[1].max()) + 1)
            fm = ref.head_tail(len(trn[0]) - 1, nc, trn[0], trn[1])  # slim_predict.c:91-95
            out["fmarker"] = fm
            ev = st.evaluate(ids, trn, tst, mv["ncols"], fm)
            out["metrics"] = np.array([ev["hr"], ev["hr_head"], ev["hr_tail"], ev["arhr"]])
            print(name, "ncols", mv["ncols"], "nnz(W)", len(mv["colind"]),
                  {k: round(v, 4) for k, v in ev.items()})
        else:
            print(name, "default nnz(W)", len(mv["colind"]))
        ref.free(h)
    np.savez_compressed(st.GOLDEN_DIR / f"{name}.npz", **out)


FSLIM_NNBRS = 10
MSELECT_L1 = [0.5, 1.0, 2.0, 4.0]   # values of the reference's test/l12file
MSELECT_L2 = [0.5, 1.0, 5.0]


def run_fslim(out, name, trn):
    ref = st.load_ref()
    for sim in ("cos", "jac", "dotp"):
        io, do = st.options(l1r=1.0, l2r=1.0, nthreads=1, opttol=1e-14, niters=100000, nnbrs=FSLIM_NNBRS, simtype=sim)
        st.libc_srand(1)
        h, status = ref.learn(trn[0], trn[1], trn[2], io, do)
        assert status == st.SLIM_OK
        mv = st.model_views(h)
        for k in ("colptr", "colind", "colval"):
            out[f"{name}_{sim}_{k}"] = mv[k]
        print("fslim", name, sim, "nnz(W)", len(mv["colind"]))
        ref.free(h)


def run_mselect(trn, tst):
    ref = st.load_ref()
    st.libc_srand(1)
    rc, best, cells = st.mselect(ref, trn, tst, MSELECT_L1, MSELECT_L2, nrcmds=10, nthreads=1, opttol=1e-14,
                                 niters=100000)
    assert rc == st.SLIM_OK and len(cells) == len(MSELECT_L1) * len(MSELECT_L2)
    print("mselect best", best)
    np.savez_compressed(st.GOLDEN_DIR / "mselect.npz", l1=np.array(MSELECT_L1), l2=np.array(MSELECT_L2), best=best,
                        cells=cells)


def main():
    assert st.have_ref(), "build oracle/_ref first: make -C oracle"
    only = sys.argv[1:]  # e.g. `make_golden.py fslim mselect` regenerates only the newer files
    fs = {}
    trn = st.read_text_csr(REF_TEST / "ml100k-train.csr")
    tst = st.read_text_csr(REF_TEST / "ml100k-test.csr")
    if not only:
        run("ml100k", trn, tst)
    if not only or "fslim" in only:
        run_fslim(fs, "ml100k", trn)
    trn = st.read_ijv(REF_TEST / "AutomotiveTrain.ijv")
    tst = st.read_ijv(REF_TEST / "AutomotiveTest.ijv", nrows=len(trn[0]) - 1)
    if not only:
        run("automotive", trn, tst)
    if not only or "fslim" in only:
        run_fslim(fs, "automotive", trn)
        np.savez_compressed(st.GOLDEN_DIR / "fslim.npz", nnbrs=np.array(FSLIM_NNBRS), **fs)
    if not only or "mselect" in only:
        run_mselect(trn, tst)


if __name__ == "__main__":
    main()
