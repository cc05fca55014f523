import sys, os, math, numpy as np
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,'tests'))
import torch, bench
from sharp_b200 import api, _lib
dev=torch.device('cuda',0)
wl=bench.workload('cfg4'); m=wl['m']
lam=bench.type_profiles(torch,dev,m,wl['types'],wl['nnz_per_cell'])
part=bench.gen_part(torch,dev,lam,6000,0,False)
p=508
ctx=api.get_context(0)
rms=[_lib.r_ranm(m,p,50+2103+k) for k in range(1,3)]
rm=ctx.upload_rm(rms)
proj=ctx.rp_project(m,6000,rm,csc=(part['p'],part['i'],part['x']),normalize=2,logkind=2)
for kk in range(2):
  for blk in range(2):
    P=proj[kk][blk*2000:(blk+1)*2000].copy()
    P=P-P.mean(1,keepdims=True); P/=np.linalg.norm(P,axis=1,keepdims=True)
    D=1-P@P.T; n=2000
    np.fill_diagonal(D,np.inf)
    size=np.ones(n); active=np.ones(n,bool)
    seq=[]; aff_tot=0
    nnv=D.argmin(1)
    while active.sum()>1:
        idx=np.flatnonzero(active)
        pairs=[(i,nnv[i]) for i in idx if nnv[nnv[i]]==i and i<nnv[i]]
        M=set()
        for a,b in pairs: M.add(a); M.add(b)
        seq.append((len(idx),len(pairs)))
        for a,b in pairs:
            dab=D[a,b]
            k=np.flatnonzero(active); k=k[(k!=a)&(k!=b)]
            new=((size[a]+size[k])*D[a,k]+(size[b]+size[k])*D[b,k]-size[k]*dab)/(size[a]+size[b]+size[k])
            D[a,k]=new; D[k,a]=new
            active[b]=False; D[b,:]=np.inf; D[:,b]=np.inf
            size[a]+=size[b]
        idx=np.flatnonzero(active)
        aff=[k for k in idx if (k in M) or (nnv[k] in M)]
        for k in aff: nnv[k]=D[k].argmin()
        aff_tot+=len(aff)
    s2=sum(a*a for a,_ in seq)
    print('member',kk,'block',blk,'rounds',len(seq),'sum n_r^2 / n^2 = %.2f'%(s2/n/n),'affected scans',aff_tot, 'seq',seq[:12],'...',seq[-8:])
