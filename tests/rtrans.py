"""Literal numpy transcriptions of the reference's R DRIVERS around the oracle's STAGE functions (projection, getrowColor,
wMetaC, sMetaC) -- TEST INFRASTRUCTURE ONLY, like orc.py.  Strings are kept as strings where the reference's behaviour
depends on them (paste() / unique() / table() on character vectors).  Each function cites the R lines it follows; the
GPU tests (and the CPU tests of the host glue through tests/fakectx.py) compare sharp_b200.api against these."""
import math

import numpy as np

import orc


def folds_of(ncells, ng):
    """R/SHARP_unlimited2.R:337-358 (== R/SHARP.R:513-536)"""
    T = int(math.ceil(ncells / ng))
    if T > 1:
        folds = np.repeat(np.arange(1, T + 1), ng)                                   # cut(seq(1, T*ng), breaks = T)
        nt = ncells - (T - 2) * ng
        nind = np.nonzero(folds == T - 1)[0]
        folds[nind[nt // 2:]] = T                                                    # nind[floor(nt/2) + 1:ng]; NA indices ignored
        return folds[:ncells], T
    return np.ones(ncells, dtype=np.int64), 1


def fpart_transcribed(x, rms, p, K, ng, reind, hc40, flag=True, colsum=None, enp_hc=None):
    """SHARP_fpart, R/SHARP_unlimited2.R:297-544.  x: dense genes x cells; hc40: the per-block parameters with
    maxN.cluster = 40 (:421); enp_hc: parameters of the per-block wMetaC (the CALLER's maxN.cluster, :481).
    -> (fColor strings "<finalC>en<t>" un-shuffled, E1 un-shuffled, folds un-shuffled)"""
    m, ncells = x.shape
    shuffle = ncells < 1e5
    src = (reind - 1) if shuffle else np.arange(ncells)                              # :326-329 E = E[, reind]
    folds, T = folds_of(ncells, ng)
    enrp = np.zeros((ncells, K), dtype=np.int64)
    enE = np.zeros((ncells, p))
    for k in range(K):                                                               # :375-452
        for t in range(1, T + 1):
            tind = np.nonzero(folds == t)[0]
            E1 = orc.rp_project(m, ncells, rms[k], dense=np.asfortranarray(x), cells=src[tind], colsum=colsum,
                                logkind=10 if flag else 0, round_digits=1)           # log10 :391, round(., 1) :410
            color, _ = orc.getrowcolor(E1, hc40)                                     # :421-423
            enrp[tind, k] = color
            enE[tind] = enE[tind] + E1                                               # :462-470
    fColor = np.empty(ncells, dtype=object)
    for t in range(1, T + 1):                                                        # :477-494
        tind = np.nonzero(folds == t)[0]
        f = orc.wmetac(enrp[tind], enp_hc or hc40)["finalC"]
        fColor[tind] = [f"{int(c)}en{t}" for c in f]
    E1 = enE / K                                                                     # :517
    if shuffle:                                                                      # :520-524
        fc = np.empty(ncells, dtype=object)
        fc[reind - 1] = fColor
        e1 = np.empty_like(E1)
        e1[reind - 1] = E1
        fo = np.empty_like(folds)
        fo[reind - 1] = folds
        fColor, E1, folds = fc, e1, fo
    return fColor, E1, folds


def table_merge_and_size_relabel(final_num, ncells, n_cluster):
    """R/SHARP_unlimited2.R:189-204 (== R/SHARP_unlimited.R:166-183): finalrowColor is a CHARACTER vector here (sMetaC
    assigns numbers into one); table() orders names as strings, sort(decreasing = TRUE) is stable"""
    final = [str(int(v)) for v in np.asarray(final_num).tolist()]
    if not n_cluster and ncells > 1e4:
        names = sorted(set(final))
        cnt = {k: 0 for k in names}
        for k in final:
            cnt[k] += 1
        small = [k for k in names if cnt[k] < 10]
        if small:
            tgt = str(min(int(k) for k in small))
            ss = set(small)
            final = [tgt if k in ss else k for k in final]
    names = sorted(set(final))
    cnt = {k: 0 for k in names}
    for k in final:
        cnt[k] += 1
    order = np.argsort(-np.array([cnt[k] for k in names]), kind="stable")
    mp = {names[j]: r + 1 for r, j in enumerate(order)}
    return np.array([mp[k] for k in final])


def unlimited2_transcribed(parts, rms, p, K, ng, seed_reinds, minN=2, maxN=None, sil_thre=0.35, height_ntimes=2.0,
                           n_cluster=0, flag=True):
    """SHARP_unlimited2, R/SHARP_unlimited2.R:29-267.  parts: list of dense genes x cells; seed_reinds: per part
    `set.seed(50); sample(n)`.  -> (pred_clusters, E1)"""
    ncells = sum(x.shape[1] for x in parts)
    maxN = max(40, math.ceil(ncells / 5000)) if maxN is None else maxN
    hc40 = orc.hc_params(min_n=minN, max_n=40, sil_thre=sil_thre, height_ntimes=height_ntimes)
    enp = orc.hc_params(min_n=minN, max_n=maxN, sil_thre=sil_thre, height_ntimes=height_ntimes)
    fC, E1s = [], []
    for i, (x, re) in enumerate(zip(parts, seed_reinds), start=1):                   # :150-170
        f, e1, _ = fpart_transcribed(x, rms, p, K, ng, re, hc40, flag, enp_hc=enp)
        fC += [f"{c}s{i}" for c in f.tolist()]                                       # :161 paste(y$fColor, "s", i)
        E1s.append(e1)
    E1 = np.concatenate(E1s, axis=0)
    codes = {c: k + 1 for k, c in enumerate(dict.fromkeys(fC))}                     # sMetaC works on unique(fColor)
    prm = orc.hc_params(n_cluster=n_cluster, min_n=minN, max_n=maxN, sil_thre=sil_thre, height_ntimes=height_ntimes)
    s = orc.smetac(np.array([codes[c] for c in fC]), E1, prm)                        # :183-185
    return table_merge_and_size_relabel(s["finalColor"], ncells, n_cluster), E1


def testlog_transcribed(x, p, cells, colsum=None):
    """testlog, R/SHARP.R:877-924, for a given subsample `cells` (0-based; the reference draws it unseeded, :884):
    project with ranM(E, p, 5) without and with log2, getrowColor(., 'ward.D', , 2, 40, sil.thre = 0, 2) each (:906-914),
    flag = msil[1] < 0.75 && msil[1] >= 0.95 * msil[2] (:918-922).  -> (flag, [maxsil no-log, maxsil log])"""
    from sharp_b200.rrng import ranM2
    m, n = x.shape
    R = ranM2(m, p, 5)                                                               # :889
    prm = orc.hc_params(min_n=2, max_n=40, sil_thre=0.0, height_ntimes=2.0)
    msil = []
    for logkind in (0, 2):
        E1 = orc.rp_project(m, n, R, dense=np.asfortranarray(x), cells=cells, colsum=colsum, logkind=logkind)
        msil.append(orc.getrowcolor(E1, prm)[1])
    return bool(msil[0] < 0.75 and msil[0] >= 0.95 * msil[1]), msil
