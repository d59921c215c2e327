"""BASELINE.json configs[0] end to end: the reference's own command-line programs (src/programs/slim_learn.c,
slim_predict.c -- compiled UNMODIFIED by oracle/Makefile) linked against this repository's libslim.so instead of
the reference library, run on the fixtures the reference ships.  The CLIs dereference the model handle as a GKlib
matrix (slim_learn.c:83 hands it to gk_csr_Write, slim_predict.c:34 reads one back), free it with
SLIM_FreeModel, and print HR / ARHR with their own evaluation loop: whatever they print must be what the reference
build prints (SURVEY.md section 8c goldens)."""
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

import slimtest as st

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
BIN = ROOT / "oracle" / "_ref"


def _write_csr(path, rp, ri, rv):
    with open(path, "w") as f:
        for u in range(len(rp) - 1):
            f.write(" ".join(f"{int(c)} {float(v):g}" for c, v in zip(ri[rp[u]:rp[u + 1]], rv[rp[u]:rp[u + 1]])) + "\n")


def _write_ijv(path, rp, ri, rv):
    with open(path, "w") as f:
        for u in range(len(rp) - 1):
            for c, v in zip(ri[rp[u]:rp[u + 1]], rv[rp[u]:rp[u + 1]]):
                f.write(f"{u} {int(c)} {float(v):g}\n")


@pytest.mark.parametrize("name,fmt", [("ml100k", "csr"), ("automotive", "ijv")])
def test_reference_cli_on_cuda_library(tmp_path, name, fmt):
    learn, predict = BIN / "b200_slim_learn", BIN / "b200_slim_predict"
    if not (learn.exists() and predict.exists()):
        pytest.skip("oracle/_ref/b200_slim_* not built (needs /root/reference at build time)")
    from slim_b200 import _lib

    if _lib.load().SLIMB200_DeviceCount() < 1:
        pytest.skip("needs a CUDA device")
    g = st.load_golden(name)
    trn, tst, mdl = tmp_path / f"train.{fmt}", tmp_path / f"test.{fmt}", tmp_path / f"model.{fmt}"
    wr = _write_csr if fmt == "csr" else _write_ijv
    wr(trn, g["trn_rowptr"], g["trn_rowind"], g["trn_rowval"])
    wr(tst, g["tst_rowptr"], g["tst_rowind"], g["tst_rowval"])
    out = subprocess.run([str(learn), f"-ifmt={fmt}", "-algo=cd", "-l1r=1", "-l2r=1", "-optTol=1e-14", "-niters=100000",
                          "-nthreads=2", str(trn), str(mdl)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "Using Coordinate Descent!" in out.stdout and "ERROR" not in out.stdout, out
    assert mdl.exists()
    nnz = sum(len(line.split()) // 2 for line in open(mdl)) if fmt == "csr" else sum(1 for _ in open(mdl))
    assert abs(nnz - len(g["W_conv_colind"])) <= 2  # #nzs 65909 (ml100k) / 84317 (Automotive)
    out = subprocess.run([str(predict), f"-ifmt={fmt}", "-nrcmds=10", str(mdl), str(trn), str(tst)],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out
    m = re.search(r"hr:\s*([0-9.]+)\s+hr_head:\s*([0-9.]+)\s+hr_tail:\s*([0-9.]+)\s+arhr:\s*([0-9.]+)", out.stdout)
    assert m, out.stdout
    got = np.array([float(x) for x in m.groups()])
    assert np.array_equal(got, np.round(g["metrics"], 4)), (got, g["metrics"])
