import os, sys, numpy as np, torch, time
import pathlib; ROOT = pathlib.Path(__file__).resolve().parent.parent; sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / 'tests'))
from slim_b200 import Staged, learn_columns
from slim_b200.synth import zipf_csr, stratified_columns
rp, ri, rv = zipf_csr(1_000_000, 100_000, 100, device='cuda')
s = Staged(rp, ri, rv)
colcnt = torch.bincount(ri.to(torch.int64), minlength=s.ncols).cpu().numpy()
cols = stratified_columns(colcnt, 12288, offset=1)
base = dict(os.environ)
for cfg in sys.argv[1:]:
    os.environ.clear(); os.environ.update(base)
    if cfg != "default":
        for kv in cfg.split(","):
            k, v = kv.split("=")
            os.environ[k] = v
    r = learn_columns(s, dict(l1r=1.0, l2r=1.0, optTol=1e-7, niters=50), cols=cols)
    print(cfg, "solve_ms %.1f" % r.solve_ms, "-> %.0f cols/s" % (12288 / r.solve_ms * 1e3), "nnz", r.nnz, flush=True)
    r.close()
