import os, sys, numpy as np, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from slim_b200 import Staged, learn_columns
from slim_b200.synth import zipf_csr, stratified_columns
nu, ni, d = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
ncs = int(sys.argv[4]); niters = int(sys.argv[5])
rp, ri, rv = zipf_csr(nu, ni, d, device='cuda')
s = Staged(rp, ri, rv)
colcnt = torch.bincount(ri.to(torch.int64), minlength=s.ncols).cpu().numpy()
cols = stratified_columns(colcnt, ncs)
print("max c", colcnt.max(), "sel heavy", (colcnt[cols] >= 2500).sum(), (colcnt[cols] >= 600).sum(), flush=True)
r = learn_columns(s, dict(l1r=1.0, l2r=1.0, optTol=1e-7, niters=niters), cols=cols)
st = r.stats()
print("ok nnz", r.nnz, "solve_ms", r.solve_ms, "niters", st["niters"][:8], flush=True)
