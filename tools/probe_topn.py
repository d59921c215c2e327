import os, sys, time, numpy as np, scipy.sparse as sp
import pathlib; ROOT = pathlib.Path(__file__).resolve().parent.parent; sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / 'tests'))
import slimtest as st
from slim_b200 import SLIM, SLIMatrix
rp, ri, rv = st.synth_zipf(100000, 10000, 50, seed=7)
R = sp.csr_matrix((rv, ri, rp), shape=(100000, 10000))
mat = SLIMatrix(R)
m = SLIM()
t0 = time.time(); m.train({"algo": "cd", "l1r": 1.0, "l2r": 1.0, "niters": 20}, mat); print("train s %.2f" % (time.time() - t0), "nnz(W)", m.to_csr().nnz, flush=True)
for host in ("1", "0"):
    os.environ["SLIMB200_PREDICT_HOST"] = host
    t0 = time.time()
    rc_ids = m.predict(mat, nrcmds=10)
    dt = time.time() - t0
    print("predict", "host loop" if host == "1" else "GPU batched", "100000 users x top-10: %.2f s (python wrapper included)" % dt, flush=True)
    if host == "1": ref = rc_ids
print("identical lists:", all(np.array_equal(ref[u], rc_ids[u]) for u in range(0, 100000, 37)))
