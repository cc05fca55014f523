import sys, os, math
import os; R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R,'tests'))
import numpy as np, torch
import bench, orc
from sharp_b200 import Context
from sharp_b200.rrng import ranM2
wl = bench.workload('cfg3')
dev = torch.device('cuda', 0)
lam = bench.type_profiles(torch, dev, wl['m'], wl['types'], wl['nnz_per_cell'])
part = bench.gen_part(torch, dev, lam, 3000, 0, False)
m, n = wl['m'], 3000
csc = (part['p'].astype(np.int64), part['i'], part['x'])
colsum = np.add.reduceat(csc[2], csc[0][:-1])
ctx = Context(0)
for K, p in ((15, 325), (15, 416), (5, 325), (8, 325), (15, 300)):
    rms = [ranM2(m, p, 50 + 2103 + k) for k in range(1, K + 1)]
    per = np.zeros(m, dtype=int)
    for r in rms: per += np.bincount(r['i'], minlength=m)
    rm = ctx.upload_rm(rms)
    out = {}
    for v in (0, 2, 3):
        ctx.set_rp_variant(v)
        out[v] = ctx.rp_project(m, n, rm, csc=csc, normalize=2, logkind=2)
    ctx.set_rp_variant(0)
    ref = np.stack([orc.rp_project(m, n, rms[k], csc=csc, colsum=colsum, logkind=2) for k in range(K)])
    sc = np.max(np.abs(ref))
    e0 = np.abs(out[0] - ref) / sc; e2 = np.abs(out[2] - ref) / sc; e3 = np.abs(out[3] - ref) / sc
    print(f"K={K} p={p} KP={K*p} max entries/gene={per.max()} mean={per.mean():.1f}: err v0={e0.max():.3e} v2={e2.max():.3e} v3={e3.max():.3e}", flush=True)
    if e0.max() > 1e-9:
        bad = np.argwhere(e0 > 1e-9)
        print('  bad entries', len(bad), 'cells', len(np.unique(bad[:,1])), 'members', np.unique(bad[:,0])[:20], 'cols sample', bad[:5])
        kk, cc, jj = bad[0]
        print('  got', out[0][kk,cc,jj], 'ref', ref[kk,cc,jj], 'v2', out[2][kk,cc,jj])
    rm.close()
