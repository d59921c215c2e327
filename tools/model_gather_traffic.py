import sys, time, numpy as np
import pathlib; ROOT = pathlib.Path(__file__).resolve().parent.parent; sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / 'tests'))
import slimtest as st
from slim_b200.synth import zipf_csr
rp, ri, rv = (t.numpy() for t in zipf_csr(1_000_000, 100_000, 100))  # the C4 matrix (CPU, ~30 s)
N=100000
cnt=np.bincount(ri,minlength=N)
order=np.lexsort((np.arange(N),-cnt))      # internal id -> original
rank=np.empty(N,np.int64); rank[order]=np.arange(N)
# CSC for users of a column
import scipy.sparse as sp
R=sp.csr_matrix((rv,ri,rp),shape=(len(rp)-1,N)); Rc=R.tocsc()
O=st.Oracle()
ranks=[700,2000,3000,5000,8000,15000,30000]
cols=[int(order[r]) for r in ranks]
t=time.time()
res=O.learn(rp,ri,rv,opttol=1e-7,niters=50,order=st.ORDER_POPULARITY,nthreads=8,cols=np.array(cols,np.int32),want_stats=True)
print('oracle s',time.time()-t)
print('%8s %7s %7s %7s %6s | %9s %9s %9s %9s'%('c_j','n','S','sweeps','dens','useful','rowk@A','rowm@S','S-sect/S'))
for q,j in enumerate(cols):
    users=Rc.indices[Rc.indptr[j]:Rc.indptr[j+1]]
    co=np.asarray(R[users].sum(axis=0)).ravel()
    A=np.nonzero((co>1.0)&(np.arange(N)!=j))[0]
    Ai=np.sort(rank[A])                       # internal ids ascending
    a,b=res['colptr'][q],res['colptr'][q+1]
    Si=np.sort(rank[res['colind'][a:b]])
    n,S=len(Ai),len(Si)
    sw=min(int(res['stats']['niters'][q]),50)
    # current: per block of 32 consecutive actives, sectors = distinct (panel, 8-col group) => distinct Ai//8
    blocks=[Ai[p:p+32] for p in range(0,n,32)]
    sect_blocks=sum(len(np.unique(bk//8)) for bk in blocks)
    cur=S*sect_blocks*32          # bytes per sweep
    # symmetric: for every active row m, gather columns S: sectors = distinct Si//8
    ssect=len(np.unique(Si//8))
    sym=n*ssect*32
    useful=n*S*4
    print('%8d %7d %7d %7d %6.2f | %8.2fGB %8.2fGB %8.2fGB %9.2f'%(cnt[j],n,S,sw,n/N,useful*sw/1e9,cur*sw/1e9,sym*sw/1e9,ssect/S))
