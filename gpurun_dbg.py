import os, sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import slimtest as st
from slim_b200 import Staged, learn_columns
rp, ri, rv = st.synth_zipf(700, 260, 24, seed=13)
kw=dict(l1r=0.7,l2r=1.5,optTol=1e-14,niters=100000)
res={}
for win in ("0","1"):
    os.environ["SLIMB200_CLUSTER"]="16"; os.environ["SLIMB200_WINDOW"]=win
    with Staged(rp,ri,rv) as s:
        r=learn_columns(s,kw); res[win]=(r.to_host(), r.stats())
a,b=res["0"][0],res["1"][0]
sa,sb=res["0"][1],res["1"][1]
bad=[]
for j in range(260):
    ia=a["colind"][a["colptr"][j]:a["colptr"][j+1]]; ib=b["colind"][b["colptr"][j]:b["colptr"][j+1]]
    va=a["colval"][a["colptr"][j]:a["colptr"][j+1]]; vb=b["colval"][b["colptr"][j]:b["colptr"][j+1]]
    if len(ia)!=len(ib) or not np.array_equal(ia,ib) or np.abs(va-vb).max()>1e-6: bad.append(j)
print("bad cols", len(bad), bad[:20])
print("niters0", sa["niters"][:20]); print("niters1", sb["niters"][:20])
print("obj0", sa["objval"][:6]); print("obj1", sb["objval"][:6])
for niters in (1,2):
    kw2=dict(l1r=0.7,l2r=1.5,optTol=1e-14,niters=niters)
    out={}
    for win in ("0","1"):
        os.environ["SLIMB200_WINDOW"]=win
        with Staged(rp,ri,rv) as s:
            r=learn_columns(s,kw2,cols=np.array([5],np.int32)); out[win]=r.to_host()
    print("niters",niters)
    for win in ("0","1"):
        print(win, out[win]["colind"][:12], out[win]["colval"][:12])
