"""CPU model (no GPU): DRAM sectors touched by the Gram-space block gathers on C4 for the fp32 panel layout of
round 1 versus the column-typed layout (u32 / u16 / u8 panels chosen by the column's nnz).
For a few targets the oracle gives the support S of w_j, R^T R's row j gives the active set A;
bytes per sweep = |S| x 32 B x distinct 32-byte sectors the 32-coordinate position blocks touch."""
import sys, time, numpy as np
import pathlib; ROOT = pathlib.Path(__file__).resolve().parent.parent; sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / 'tests'))
import slimtest as st
from slim_b200.synth import zipf_csr
import scipy.sparse as sp
rp, ri, rv = (t.numpy() for t in zipf_csr(1_000_000, 100_000, 100))
N = 100000
cnt = np.bincount(ri, minlength=N)
order = np.lexsort((np.arange(N), -cnt))
rank = np.empty(N, np.int64); rank[order] = np.arange(N)
csort = cnt[order]                      # nnz by internal id (descending)
H32 = int(np.searchsorted(-csort, -65535, side='left'))   # first internal id with c <= 65535
H16 = int(np.searchsorted(-csort, -255, side='left'))     # first internal id with c <= 255
H32 = (H32 + 63) // 64 * 64; H16 = (H16 + 63) // 64 * 64
print('N', N, 'u32 cols', H32, 'u16 cols', H16 - H32, 'u8 cols', N - H16,
      'bytes per row: fp32 %d typed %d' % (4 * N, 4 * H32 + 2 * (H16 - H32) + (N - H16)))
def sector_typed(i):
    i = np.asarray(i)
    s = np.where(i < H32, i // 8, np.where(i < H16, 10_000_000 + (i - H32) // 16, 20_000_000 + (i - H16) // 32))
    return s
R = sp.csr_matrix((rv, ri, rp), shape=(len(rp) - 1, N)); Rc = R.tocsc()
O = st.Oracle()
ranks = [150, 400, 700, 2000, 3000, 5000, 8000, 15000, 30000, 60000]
cols = [int(order[r]) for r in ranks]
t = time.time()
res = O.learn(rp, ri, rv, opttol=1e-7, niters=50, order=st.ORDER_POPULARITY, nthreads=8, cols=np.array(cols, np.int32), want_stats=True)
print('oracle s', time.time() - t)
print('%8s %7s %7s %6s %5s | %9s %9s %6s %9s %6s' % ('c_j', '|A|', '|S|', 'sweeps', 'dens', 'useful4B', 'fp32', 'x', 'typed', 'x'))
for q, j in enumerate(cols):
    users = Rc.indices[Rc.indptr[j]:Rc.indptr[j + 1]]
    co = np.asarray(R[users].sum(axis=0)).ravel()
    A = np.nonzero((co > 1.0) & (np.arange(N) != j))[0]
    Ai = np.sort(rank[A])
    a, b = res['colptr'][q], res['colptr'][q + 1]
    S = b - a
    n = len(Ai)
    sw = min(int(res['stats']['niters'][q]), 50)
    blocks = [Ai[p:p + 32] for p in range(0, n, 32)]
    s32 = sum(len(np.unique(bk // 8)) for bk in blocks)
    sty = sum(len(np.unique(sector_typed(bk))) for bk in blocks)
    useful = n * S * 4
    print('%8d %7d %7d %6d %5.2f | %8.2fGB %8.2fGB %6.2f %8.2fGB %6.2f  nA<H32 %d nA<H16 %d' % (
        cnt[j], n, S, sw, n / N, useful * sw / 1e9, S * s32 * 32 * sw / 1e9, s32 * 32 / (n * 4), S * sty * 32 * sw / 1e9,
        s32 / max(sty, 1), (Ai < H32).sum(), (Ai < H16).sum()))
