"""One C5 step (BASELINE configs[4]: 5M x 500K, 500M nnz) per plan configuration; the matrix is restaged only when a
staging-time variable (SLIMB200_GRAM*, layout / head square) changes.

    python tools/probe_c5_plan.py NCOLS default SLIMB200_STAIR_USER=5000 SLIMB200_GRAM_HD=65536,SLIMB200_STAIR_USER=5000

Prints the solve time and, per column-nnz class, the number of targets, their summed / largest kernel time, the mean
active-set size, sweeps and nnz(w_j) (= |S| at the end)."""
import os, sys, numpy as np, torch
import pathlib; ROOT = pathlib.Path(__file__).resolve().parent.parent; sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / 'tests'))
from slim_b200 import Staged, learn_columns
from slim_b200.synth import zipf_csr, stratified_columns

STAGE_KEYS = ("SLIMB200_GRAM_HD", "SLIMB200_GRAM_LAYOUT", "SLIMB200_GRAM_GB", "SLIMB200_GRAM")
ncols_step = int(sys.argv[1])
nu, ni, pu = (int(v) for v in os.environ.get("PROBE_SHAPE", "5000000,500000,100").split(","))
rp, ri, rv = zipf_csr(nu, ni, pu, device='cuda')
colcnt = torch.bincount(ri.to(torch.int64), minlength=ni).cpu().numpy()
cols = stratified_columns(colcnt, ncols_step, offset=1)
all_cols = cols
base = dict(os.environ)
staged, staged_key = None, None
for cfg in sys.argv[2:]:
    os.environ.clear(); os.environ.update(base)
    kv = {} if cfg == "default" else dict(x.split("=") for x in cfg.split(","))
    os.environ.update(kv)
    lo_nnz = int(kv.get("PROBE_MIN_NNZ", "0"))  # only the step's targets with at least this many nonzeros
    cols = all_cols[colcnt[all_cols] >= lo_nnz]
    key = tuple(sorted((k, v) for k, v in kv.items() if k in STAGE_KEYS))
    if staged is None or key != staged_key:
        if staged is not None:
            staged.close()
        staged, staged_key = Staged(rp, ri, rv), key
        print("# staged: %.0f ms, gram %s, layout %s, stair %s" % (staged.stage_ms, staged.gram_info(), staged.gram_layout(),
                                                                staged.gram_stair()), flush=True)
    r = learn_columns(staged, dict(l1r=1.0, l2r=1.0, optTol=1e-7, niters=50), cols=cols)
    st, (ph, ng), w = r.stats(), r.phases(), r.to_host()
    wn = np.diff(w["colptr"])
    print(cfg, "solve_ms %.1f" % r.solve_ms, "-> %.0f cols/s" % (len(cols) / r.solve_ms * 1e3), "nnz", r.nnz, flush=True)
    c = colcnt[cols]
    tot = ph.sum(1) * 1e-3
    for lo, hi in ((30000, 1 << 30), (9000, 30000), (5000, 9000), (2000, 5000), (500, 2000), (100, 500), (0, 100)):
        m = (c >= lo) & (c < hi)
        if m.any():
            print("   nnz [%6d,%9d): %5d targets, kernel time sum %9.1f ms max %9.1f ms | active-set phase %6.2f ms | "
                  "nactive %8.0f sweeps %5.1f nnz(w) %7.0f" % (lo, hi, m.sum(), tot[m].sum(), tot[m].max(), ph[m, 1].mean() * 1e-3,
                                                            st["nactive"][m].mean(), st["niters"][m].mean(), wn[m].mean()), flush=True)
    r.close()
