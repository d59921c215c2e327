import os, sys, numpy as np, torch, time
import pathlib; ROOT = pathlib.Path(__file__).resolve().parent.parent; sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / 'tests'))
from slim_b200 import Staged, learn_columns
from slim_b200.synth import zipf_csr, stratified_columns
rp, ri, rv = zipf_csr(1_000_000, 100_000, 100, device='cuda')
s = Staged(rp, ri, rv)
colcnt = torch.bincount(ri.to(torch.int64), minlength=s.ncols).cpu().numpy()
order = np.argsort(-colcnt, kind='stable')
cols = np.sort(order[[40,80,120,160,200,240,280,320, 500,600,700,800,900,1000,1100,1200, 2000,2200,2400,2600,2800,3000,3200,3400]]).astype(np.int32)
niters = int(os.environ.get("DBG_NITERS", "50"))
for cfg in sys.argv[1:]:
    for kv in cfg.split(","):
        k, v = kv.split("=")
        os.environ[k] = v
    t0 = time.time()
    r = learn_columns(s, dict(l1r=1.0, l2r=1.0, optTol=1e-7, niters=niters), cols=cols)
    print(cfg, "solve_ms %.1f" % r.solve_ms, "nnz", r.nnz, flush=True)
    r.close()
