"""Second, independent restatement of wMetaC: a LITERAL numpy transcription of R/wMetaC.R (dense N x N co-association
matrices, string labels, every pair of clusters through getss/getnewk) against the C++ oracle, which computes the
same quantities without the N x N matrices.  The reference ships no known answers for this path (DESIGN.md section 2:
parity unpinned); two restatements written from the same R lines by different routes agreeing on every label is the
strongest pin available without R.  Only the get_opt_hclust call inside is shared (the oracle's, itself cross-checked
against scipy / scikit-learn in test_oracle_primitives.py)."""
import numpy as np
import pytest

import orc


def _geta(col):
    """R/wMetaC.R:242-283: A[i, j] = 1 when points i and j carry the same label"""
    col = np.asarray(col)
    return (col[:, None] == col[None, :]).astype(np.float64)


def _ldsum(values):
    """R's sum(): long double accumulator, in the order given"""
    acc = np.longdouble(0.0)
    for v in values:
        acc += np.longdouble(v)
    return float(acc)


def wmetac_transcribed(nC, prm):
    N, C = nC.shape
    AA = sum(_geta(nC[:, i]) for i in range(C)) / C                      # :24-25
    nd = np.where(AA != 0, AA * (1.0 - AA), 0.0)                         # :30-36 (newAA, dense here)
    w0 = 4.0 / N * np.array([_ldsum(nd[i, nd[i] != 0]) for i in range(N)])   # :40 rowSums
    w1 = (w0 + 0.01) / (1 + 0.01)                                        # :42-43
    x = [f"{nC[r, i]}_{i + 1}" for i in range(C) for r in range(N)]      # :60-62 column-major vector of labels
    R = list(dict.fromkeys(x))                                           # :66 unique(): first-appearance order
    allC = len(R)
    x_arr = np.array(x)

    def newk(k):                                                         # getnewk :315-320 (0-based point indices)
        k1 = np.nonzero(x_arr == R[k])[0]
        d = int(R[k].rsplit("_", 1)[1])
        return k1 - (d - 1) * N

    members = [newk(k) for k in range(allC)]
    S = np.eye(allC)
    for a in range(allC):                                                # combn(allC, 2) :70-71, getss :299-312
        for b in range(a + 1, allC):
            inter = [i for i in members[a] if i in set(members[b])]      # intersect(): order of the first argument
            if inter:
                seen = set(members[a])
                union = list(members[a]) + [i for i in members[b] if i not in seen]   # union(): x, then the new y
                S[a, b] = S[b, a] = _ldsum(w1[inter]) / _ldsum(w1[union])
    hres = orc.opt_hclust(S, symmetric=1, prm=prm)                       # :94-95
    tf = hres["f"]
    newnC = np.empty((N, C), dtype=np.int64)                             # :141 tf[match(q, R)]
    pos = {lab: i for i, lab in enumerate(R)}
    for i in range(C):
        for r in range(N):
            newnC[r, i] = tf[pos[x[i * N + r]]]

    def vote(row, second=False):                                         # :143, :147-157 sort(table(d), decreasing = TRUE)
        vals, cnt = np.unique(row, return_counts=True)                   # table(): ascending names
        order = np.argsort(-cnt, kind="stable")                          # stable: ties keep the smaller id first
        if second and len(order) > 1 and cnt[order[1]] >= 1 * 0.5:       # n0 = length(x[1]) = 1
            return vals[order[1]]
        return vals[order[0]]

    finalC = np.array([vote(newnC[r]) for r in range(N)])
    if len(np.unique(finalC)) == 1:
        finalC = np.array([vote(newnC[r], second=True) for r in range(N)])
    uC = list(dict.fromkeys(finalC.tolist()))                            # :180 unique(finalC)
    x0 = np.zeros((N, len(uC)))
    for r in range(N):                                                   # :182-208
        y0 = np.array([np.sum(newnC[r] == u) for u in uC], dtype=np.float64)
        xind = uC.index(finalC[r])
        x0[r, xind] = 1.0
        for j in np.nonzero(y0)[0]:
            if j != xind:
                x0[r, j] = 0.5 * y0[j] / y0[xind]
    return finalC, x0, hres["optN.cluster"]


@pytest.mark.parametrize("N,C,g,noise,seed", [(120, 4, 3, 0.15, 1), (90, 5, 4, 0.3, 2), (150, 3, 2, 0.05, 3), (80, 6, 5, 0.4, 4)])
def test_wmetac_literal_transcription_agrees_with_the_oracle(N, C, g, noise, seed):
    rng = np.random.default_rng(seed)
    truth = rng.integers(1, g + 1, size=N)
    nC = np.empty((N, C), dtype=np.int64)
    for i in range(C):
        perm = rng.permutation(g) + 1                                    # every solution numbers the clusters its own way
        lab = perm[truth - 1]
        flip = rng.random(N) < noise
        lab[flip] = rng.integers(1, g + 2, size=int(flip.sum()))          # noise, sometimes a spurious extra cluster
        nC[:, i] = lab
    prm = orc.hc_params(sil_thre=0.0)                                     # wMetaC's default when sil.thre is missing (:89-92)
    final_t, x0_t, optn_t = wmetac_transcribed(nC, prm)
    ref = orc.wmetac(nC, prm)
    assert np.array_equal(np.asarray(ref["finalC"]).astype(np.int64), final_t)
    assert ref["x0"].shape == x0_t.shape and np.allclose(ref["x0"], x0_t, rtol=0, atol=0)


# ---------------------------------------------------------------------------------------------------------
# sMetaC (R/sMetaC.R:17-210)
# ---------------------------------------------------------------------------------------------------------
def _r_cor(a, b):
    """stats::cor(x, y) for two vectors as src/library/stats/src/cov.c does it: long double means (with the second
    refinement pass), long double sums of products, clamped to [-1, 1]"""
    ld = np.longdouble
    a, b = a.astype(ld), b.astype(ld)
    n = len(a)

    def mean(x):
        m = x.sum(dtype=ld) / n
        return m + (x - m).sum(dtype=ld) / n

    xm, ym = mean(a), mean(b)
    sxy = ((a - xm) * (b - ym)).sum(dtype=ld)
    sxx = ((a - xm) ** 2).sum(dtype=ld)
    syy = ((b - ym) ** 2).sum(dtype=ld)
    r = sxy / (np.sqrt(sxx) * np.sqrt(syy))
    return float(min(max(r, ld(-1.0)), ld(1.0)))


def smetac_transcribed(lab, sE1, prm):
    lab = np.asarray(lab)
    R = list(dict.fromkeys(lab.tolist()))                                 # :21 unique(rerowColor)
    nC = len(R)
    aG = np.stack([(sE1[lab == R[t]].astype(np.longdouble).sum(axis=0) / np.sum(lab == R[t])).astype(np.float64)
                   for t in range(nC)])                                   # :58-63 colMeans
    S = np.eye(nC)
    for a in range(nC):                                                   # :67-85
        for b in range(a + 1, nC):
            S[a, b] = S[b, a] = _r_cor(aG[a], aG[b])
    ncells = len(lab)
    min_n, max_n = prm.min_n, prm.max_n
    if ncells < 1e6:                                                      # :101-109
        base = min(max(ncells // 10000, 2), 10)
        if min_n == 2 and min(max_n, nC) - base >= 3:
            min_n = base
    else:                                                                 # :110-119
        max_n = max(max_n, ncells // 5000)
        min_n = max(min_n, ncells // 50000)
    p2 = orc.hc_params(prm.hmethod, prm.n_cluster, min_n, max_n, prm.sil_thre, prm.height_ntimes)
    hres = orc.opt_hclust(S, symmetric=1, prm=p2)                         # :121-122
    s0 = hres["msil"]
    n = len(s0)
    if n > 1 and len(np.unique(hres["f"])) == 2 and hres["maxsil"] > prm.sil_thre:    # :139-147
        s1 = np.sort(s0)[n - 2]
        s2 = int(np.nonzero(s0 == s1)[0][0])                              # quirk B4: the first of several
        tf = hres["v"][:, s2]
    else:
        tf = hres["f"]
    pos = {c: i for i, c in enumerate(R)}
    final = np.array([tf[pos[c]] for c in lab.tolist()])                  # :182
    return final, np.asarray(tf)


@pytest.mark.parametrize("ncells,p,nclu,g,seed", [(3000, 40, 18, 4, 1), (5200, 64, 30, 6, 2), (1500, 24, 9, 2, 3)])
def test_smetac_literal_transcription_agrees_with_the_oracle(ncells, p, nclu, g, seed):
    rng = np.random.default_rng(seed)
    centres = rng.normal(size=(g, p)) * 2.0
    group_of_cluster = rng.integers(0, g, size=nclu)
    group_of_cluster[:g] = np.arange(g)
    lab = rng.integers(1, nclu + 1, size=ncells)
    sE1 = centres[group_of_cluster[lab - 1]] + rng.normal(size=(ncells, p))
    prm = orc.hc_params()
    final_t, tf_t = smetac_transcribed(lab, sE1, prm)
    ref = orc.smetac(lab, sE1, prm)
    assert np.array_equal(np.asarray(ref["tf"]).astype(np.int64), tf_t.astype(np.int64))
    assert np.array_equal(np.asarray(ref["finalColor"]).astype(np.int64), final_t.astype(np.int64))


# ---------------------------------------------------------------------------------------------------------
# get_opt_hclust, feature branch (R/get_opt_hclust.R:66-229), with scipy / scikit-learn doing the work of
# stats::hclust, cutree and cluster::silhouette
# ---------------------------------------------------------------------------------------------------------
def opt_hclust_transcribed(mat, min_n=2, max_n=40, sil_thre=0.35):
    from scipy.cluster.hierarchy import cut_tree, linkage
    from scipy.spatial.distance import squareform
    from sklearn.metrics import silhouette_samples
    z = (mat - mat.mean(axis=1, keepdims=True)) / mat.std(axis=1, ddof=1, keepdims=True)    # t(scale(t(mat))) :68
    d = 1.0 - np.corrcoef(z)                                                                # 1 - cor(t(mat)) :69
    np.fill_diagonal(d, 0.0)
    d = (d + d.T) / 2
    # stats::hclust(d, "ward.D") runs Ward's recurrence on the dissimilarities as given; scipy's ward runs it on squared
    # input: ward.D(d) == scipy ward(sqrt(d)) with heights squared (test_oracle_primitives.py)
    Z = linkage(np.sqrt(squareform(d, checks=False)), method="ward")
    n = mat.shape[0]
    ks = list(range(min_n, min(max_n, n - 1) + 1))                                          # :113
    v = np.empty((n, len(ks)), dtype=np.int64)
    msil = np.empty(len(ks))
    for i, k in enumerate(ks):
        lab = cut_tree(Z, n_clusters=k).ravel()
        _, first = np.unique(lab, return_index=True)                                        # cutree numbers by first appearance
        rank = np.empty(len(first), dtype=np.int64)
        rank[np.argsort(first)] = np.arange(1, len(first) + 1)
        v[:, i] = rank[lab]
        msil[i] = np.median(silhouette_samples(d, v[:, i], metric="precomputed"))           # :134-137
    tmp = np.nonzero(msil == msil.max())[0]                                                 # :162-167
    oind = tmp[int(np.ceil(len(tmp) / 2)) - 1]
    assert msil.max() > sil_thre, "this transcription only covers the silhouette route"
    return v[:, oind], len(np.unique(v[:, oind])), msil


@pytest.mark.parametrize("n,p,g,sep,seed", [(300, 40, 4, 2.0, 1), (450, 60, 7, 2.5, 2), (200, 25, 3, 3.0, 3)])
def test_opt_hclust_transcription_with_scipy_and_sklearn_agrees_with_the_oracle(n, p, g, sep, seed):
    rng = np.random.default_rng(seed)
    centres = rng.normal(size=(g, p)) * sep
    truth = rng.integers(0, g, size=n)
    mat = centres[truth] + rng.normal(size=(n, p))
    f_t, optn_t, msil_t = opt_hclust_transcribed(mat)
    ref = orc.opt_hclust(mat, symmetric=0)
    assert ref["optN.cluster"] == optn_t
    assert np.array_equal(np.asarray(ref["f"]).astype(np.int64), f_t)
    assert np.allclose(ref["msil"], msil_t, rtol=1e-10, atol=1e-12)


# ---------------------------------------------------------------------------------------------------------
# SHARP_large (R/SHARP.R:478-851): the orchestration -- shuffle, folds, gather, per-block wMetaC, sMetaC, un-shuffle,
# small-cluster merge, relabel -- transcribed in numpy around the oracle's STAGE functions, against the oracle's own
# driver (which is what the GPU pipeline is compared with)
# ---------------------------------------------------------------------------------------------------------
def sharp_large_transcribed(x, rms, p, K, ng, reind, hc):
    m, ncells = x.shape
    E = x[:, reind - 1] if ncells < 1e5 else x                                       # :493-505 E = E[, reind]
    T = int(np.ceil(ncells / ng))                                                    # :507-508
    if T > 1:                                                                        # :509-522
        folds = np.repeat(np.arange(1, T + 1), ng)                                   # cut(seq(1, T*ng), breaks = T)
        nt = ncells - (T - 2) * ng
        nind = np.nonzero(folds == T - 1)[0]
        folds[nind[nt // 2:]] = T                                                    # nind[floor(nt/2) + 1:ng]; NA indices ignored
        folds = folds[:ncells]
    else:
        folds = np.ones(ncells, dtype=np.int64)
    enrp = np.zeros((ncells, K), dtype=np.int64)
    enE = np.zeros((ncells, p))
    for k in range(K):                                                               # :554-618, gathered as in :627-635
        for t in range(1, T + 1):
            tind = np.nonzero(folds == t)[0]
            E1 = orc.rp_project(m, ncells, rms[k], dense=np.asfortranarray(E), cells=tind, logkind=2)   # :567-581
            color, _ = orc.getrowcolor(E1, hc)                                       # :584-585
            enrp[tind, k] = color
            enE[tind] = enE[tind] + E1
    fColor = np.empty(ncells, dtype=object)
    for t in range(1, T + 1):                                                        # :692-709
        tind = np.nonzero(folds == t)[0]
        f = orc.wmetac(enrp[tind], hc)["finalC"]
        fColor[tind] = [f"{c}en{t}" for c in f]
    assert T > 1
    codes = {c: i + 1 for i, c in enumerate(dict.fromkeys(fColor.tolist()))}        # sMetaC works on unique(fColor)
    s = orc.smetac(np.array([codes[c] for c in fColor.tolist()]), enE / K, hc)       # :749-753
    Srow = np.asarray(s["finalColor"]).astype(np.int64)
    final = np.empty(ncells, dtype=np.int64)
    viE = np.empty((ncells, p))
    if ncells < 1e5:                                                                 # :776-781
        final[reind - 1] = Srow
        viE[reind - 1] = enE / K
    else:
        final, viE = Srow, enE / K
    if ncells > 1e4:                                                                 # :816-825
        vals, cnt = np.unique(final, return_counts=True)
        small = vals[cnt < 10]
        if len(small):
            final[np.isin(final, small)] = small.min()
    uy = list(dict.fromkeys(final.tolist()))                                         # :828-832 match(y, unique(y))
    pos = {v: i + 1 for i, v in enumerate(uy)}
    return np.array([pos[v] for v in final.tolist()]), viE


@pytest.mark.parametrize("m,n,K,ng,seed", [(700, 1300, 3, 400, 5), (500, 11000, 2, 2000, 6)])
def test_sharp_large_orchestration_transcribed(m, n, K, ng, seed):
    import math

    import synth
    from sharp_b200.rrng import r_sample_perm, ranM2
    x, _ = synth.make_expression(m, n, n_types=4, seed=seed, kind="tpm", zero_frac=0.75, sep=2.0, frac=0.4)
    p = math.ceil(math.log2(n) / 0.04)
    p = min(p, 120)                                                                   # keep the CPU test quick
    rms = [ranM2(m, p, 50 + 11 + k) for k in range(1, K + 1)]
    reind = np.asarray(r_sample_perm(n, 50))
    hc = orc.hc_params()
    pred_t, vie_t = sharp_large_transcribed(np.asarray(x), rms, p, K, ng, reind, hc)
    prm = orc.SharpParams(1, 1, K, p, ng, 0, 0, 0, hc, 2, -1)
    ref = orc.sharp(m, n, rms, prm, dense=x, reind=reind)
    assert np.array_equal(np.asarray(ref["pred_clusters"]).astype(np.int64), pred_t)
    assert np.allclose(ref["viE"], vie_t, rtol=1e-13, atol=1e-13)


# ---------------------------------------------------------------------------------------------------------
# SHARP_unlimited's global step (R/SHARP_unlimited.R:125-183)
# ---------------------------------------------------------------------------------------------------------
def unlimited_combine_transcribed(part_of, pred, E1, hc, n_cluster=0):
    ncells = len(pred)
    fColor = [f"{c}s{i}" for c, i in zip(pred.tolist(), part_of.tolist())]             # :134 paste(pred, "s", i)
    codes = {c: k + 1 for k, c in enumerate(dict.fromkeys(fColor))}
    prm = orc.hc_params(hc.hmethod, n_cluster, hc.min_n, hc.max_n, hc.sil_thre, hc.height_ntimes)
    s = orc.smetac(np.array([codes[c] for c in fColor]), E1, prm)                       # :151-153
    # sMetaC assigns numbers INTO a character vector (R/sMetaC.R:182), so finalColor is character from here on:
    final = [str(int(v)) for v in np.asarray(s["finalColor"]).tolist()]
    if not n_cluster and ncells > 1e4:                                                  # :158-166
        names = sorted(set(final))                                                      # table(): names in string order
        cnt = {k: final.count(k) for k in names}
        small = [k for k in names if cnt[k] < 10]
        if small:
            tgt = str(min(int(k) for k in small))
            final = [tgt if k in small else k for k in final]
    names = sorted(set(final))                                                          # :168 sort(table(.), decreasing = TRUE):
    cnt = np.array([final.count(k) for k in names])                                     # stable, ties keep the STRING order
    order = np.argsort(-cnt, kind="stable")
    mp = {names[j]: r + 1 for r, j in enumerate(order)}                                 # :169-171 map[finalrowColor] by NAME
    return np.array([mp[k] for k in final])


@pytest.mark.parametrize("nparts,per_part,g,fixed,seed", [(3, 400, 12, 12, 1), (4, 300, 5, 0, 2), (2, 6000, 3, 0, 3)])
def test_unlimited_combine_transcribed(nparts, per_part, g, fixed, seed):
    """g = 12 equally large planted groups with N.cluster = 12: every cluster size ties, so the final numbering is decided
    by table()'s STRING order of the names ("1", "10", "11", "12", "2", ...)"""
    rng = np.random.default_rng(seed)
    p = 30
    centres = rng.normal(size=(g, p)) * 3.0
    part_of, pred, rows = [], [], []
    for i in range(1, nparts + 1):
        truth = np.arange(per_part) % g
        rng.shuffle(truth)
        relabel = rng.permutation(g) + 1                                                 # every part numbers its clusters its own way
        pred.append(relabel[truth])
        part_of.append(np.full(per_part, i))
        rows.append(centres[truth] + 0.3 * rng.normal(size=(per_part, p)))
    part_of, pred, E1 = np.concatenate(part_of), np.concatenate(pred), np.concatenate(rows)
    hc = orc.hc_params()
    got_t = unlimited_combine_transcribed(part_of, pred, E1, hc, fixed)
    got, nf = orc.unlimited_combine(part_of, pred, E1, hc, fixed)
    assert np.array_equal(np.asarray(got).astype(np.int64), got_t)
    assert nf == len(np.unique(got_t))
    if fixed:
        assert nf == fixed


# ---------------------------------------------------------------------------------------------------------
# SHARP_small (R/SHARP.R:339-454): K x (RPmat -> getrowColor), one wMetaC over all cells, relabel
# ---------------------------------------------------------------------------------------------------------
def sharp_small_transcribed(x, rms, p, K, hc):
    m, ncells = x.shape
    enrp = np.zeros((ncells, K), dtype=np.int64)
    enE = np.zeros((ncells, p))
    for k in range(K):                                                                # :349-372
        E1 = orc.rp_project(m, ncells, rms[k], dense=np.asfortranarray(x), logkind=2)   # log2(scExp + 1) :343-345, RPmat :356
        color, _ = orc.getrowcolor(E1, hc)
        enrp[:, k] = color
        enE = enE + E1                                                                # :380-385
    fC = orc.wmetac(enrp, hc)                                                         # :390-391
    final = np.asarray(fC["finalC"]).astype(np.int64)
    uy = list(dict.fromkeys(final.tolist()))                                          # :430-432
    pos = {v: i + 1 for i, v in enumerate(uy)}
    return np.array([pos[v] for v in final.tolist()]), enE / K, fC["x0"]


@pytest.mark.parametrize("m,n,K,seed", [(600, 350, 4, 7), (900, 479, 3, 8)])
def test_sharp_small_orchestration_transcribed(m, n, K, seed):
    import math

    import synth
    from sharp_b200.rrng import ranM2
    x, _ = synth.make_expression(m, n, n_types=3, seed=seed, kind="tpm", zero_frac=0.7, sep=2.0, frac=0.4)
    p = math.ceil(math.log2(n) / 0.04)
    rms = [ranM2(m, p, 50 + 2103 + k) for k in range(1, K + 1)]
    hc = orc.hc_params()
    pred_t, vie_t, x0_t = sharp_small_transcribed(np.asarray(x), rms, p, K, hc)
    prm = orc.SharpParams(0, 1, K, p, 2000, 0, 0, 0, hc, 2, -1)
    ref = orc.sharp(m, n, rms, prm, dense=x)
    assert np.array_equal(np.asarray(ref["pred_clusters"]).astype(np.int64), pred_t)
    assert np.allclose(ref["viE"], vie_t, rtol=1e-13, atol=1e-13)
    assert ref["x0"].shape == x0_t.shape and np.array_equal(ref["x0"], x0_t)


def test_getrowcolor_transcribed_including_the_colour_wrap():
    """R/getrowColor.R:35-68: colour j of the j-th entry of unique(f); beyond 40 colours the index wraps (j %% 40, 0 -> 40),
    so with indN.cluster = 45 the clusters 41..45 SHARE the colours (and therefore the labels) of clusters 1..5"""
    rng = np.random.default_rng(5)
    E = rng.normal(size=(260, 30))
    for ncl in (0, 7, 45):
        prm = orc.hc_params(n_cluster=ncl)
        f = np.asarray(orc.opt_hclust(E, symmetric=0, prm=prm)["f"]).astype(np.int64)
        unf = list(dict.fromkeys(f.tolist()))
        colour = np.zeros(len(f), dtype=np.int64)
        for j, u in enumerate(unf, start=1):
            jj = j
            if jj > 40:
                jj = jj % 40
                if jj == 0:
                    jj = 40
            colour[f == u] = jj
        got, _ = orc.getrowcolor(E, prm)
        assert np.array_equal(np.asarray(got).astype(np.int64), colour)
        if ncl == 45:
            assert len(np.unique(colour)) == 40
