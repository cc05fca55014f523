"""Simulate the round-parallel Ward agglomeration (reciprocal-NN merging) on a 1-cor distance block like the bench's,
to get per-round statistics: nr, m (pairs), rows needing rescan, and traffic under alternative designs."""
import numpy as np, sys, math
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
rng = np.random.default_rng(0)
n, p, g = 2000, 508, 20
centres = rng.normal(size=(g, p)) * 0.35
truth = rng.integers(0, g, size=n)
X = centres[truth] + rng.normal(size=(n, p))
Z = X - X.mean(1, keepdims=True); Z /= np.linalg.norm(Z, axis=1, keepdims=True)
D = 1.0 - Z @ Z.T
np.fill_diagonal(D, np.inf)
size = np.ones(n)
alive = np.arange(n)
stats = []
Dc = D.copy()
nr = n
rounds = 0
while nr > 1:
    nn = np.argmin(Dc, axis=1)
    recip = nn[nn] == np.arange(nr)
    keep = recip & (np.arange(nr) < nn)         # kept member of each pair
    ret = recip & (np.arange(nr) > nn)
    m = int(keep.sum())
    # rows whose NN was merged (excluding pair members)
    nn_merged = recip[nn] & ~recip
    resc = int(nn_merged.sum())
    stats.append((nr, m, resc))
    # build next matrix (Ward LW on unsquared dissimilarities, ward.D)
    a = np.nonzero(keep)[0]; b = nn[a]
    newrows = []
    # vectorised LW for all pairs vs all clusters
    sa, sb = size[a][:, None], size[b][:, None]
    sk = size[None, :]
    dab = Dc[a, b][:, None]
    Da, Db = Dc[a], Dc[b]
    Da = np.where(np.isinf(Da), 0, Da); Db = np.where(np.isinf(Db), 0, Db)
    merged = ((sa + sk) * Da + (sb + sk) * Db - sk * dab) / (sa + sb + sk)   # m x nr (distance of merged pair to every old cluster)
    surv = ~ret
    idx = np.nonzero(surv)[0]
    nnew = len(idx)
    pos = -np.ones(nr, dtype=int); pos[idx] = np.arange(nnew)
    Dn = Dc[np.ix_(idx, idx)].copy()
    # rows/cols of kept members become merged distances; pair-pair distances need two-step LW: approximate by applying sequentially
    for q in range(m):
        ip = pos[a[q]]
        row = merged[q, idx].copy()
        Dn[ip, :] = row; Dn[:, ip] = row
    # pair-pair: recompute properly via sequential LW
    if m > 1:
        for q in range(m):
            for r in range(q + 1, m):
                # distance between merged(q) and merged(r): LW of merged(q) against r's members, using merged[q] values
                dq_ar, dq_br = merged[q, a[r]], merged[q, b[r]]
                sr_a, sr_b = size[a[r]], size[b[r]]
                sq = size[a[q]] + size[b[q]]
                v = ((sr_a + sq) * dq_ar + (sr_b + sq) * dq_br - sq * Dc[a[r], b[r]]) / (sr_a + sr_b + sq)
                Dn[pos[a[q]], pos[a[r]]] = v; Dn[pos[a[r]], pos[a[q]]] = v
    np.fill_diagonal(Dn, np.inf)
    size2 = size[idx].copy()
    size2[pos[a]] = size[a] + size[b]
    size = size2; Dc = Dn; nr = nnew; rounds += 1
print('rounds', rounds)
tot_read = n * n  # initial NN scan
tot_write = 0
resc_rows = 0
for (nr, m, resc) in stats:
    tot_read += nr * nr
    tot_write += (nr - m) ** 2
    resc_rows += resc
print('streaming rebuild: reads %.2f n^2, writes %.2f n^2' % (tot_read / n**2, tot_write / n**2))
print('m/nr by round (first 12):', [round(m / nr, 3) for nr, m, _ in stats[:12]])
print('rescans/m (first 12):', [round(r / max(m, 1), 2) for nr, m, r in stats[:12]])
# in-place design without compaction every round: per round m*(2 reads + 1 write) rows + rescans, rows of length L (last compaction size)
for R in (1, 2, 4, 8, 1000):
    reads = n * n; writes = 0; L = n; since = 0
    for (nr, m, resc) in stats:
        if since == R:
            reads += L * L * (nr / L) ** 0 ; writes += nr * nr; L = nr; since = 0   # compaction pass: read L x L live part (~L*L), write nr*nr
        reads += (2 * m + resc) * L + m * nr * 4 * 0   # row reads
        writes += m * L + m * nr * 4                    # new row + mirrored column as 32-byte sectors (x4)
        since += 1
    print('in-place, compact every %d rounds: reads %.2f n^2, writes %.2f n^2, total %.2f' % (R, reads / n**2, writes / n**2, (reads + writes) / n**2))
