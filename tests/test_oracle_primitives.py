"""CPU tests that PIN THE ORACLE (oracle/sharp_oracle.cpp, the CPU restatement of the reference): the reference
ships no tests or golden vectors for this path ("parity unpinned"), so every R primitive the oracle
restates is checked against (i) hand-computable known answers, (ii) published R outputs that are common knowledge
(set.seed(42); runif(3) / sample(1:10)), and (iii) independent implementations available in this image (scipy,
scikit-learn).  Runs without a GPU."""
import numpy as np
import pytest
import scipy.cluster.hierarchy as sch
from scipy.spatial.distance import squareform
from sklearn.metrics import adjusted_rand_score, calinski_harabasz_score, silhouette_samples

import orc
import synth
from sharp_b200 import _lib
from sharp_b200.rrng import RRandom, r_sample_perm, ranM2


# ------------------------------------------------------------------------------------------------- A.1  RNG
def test_r_set_seed_runif_known_answers():
    # R: set.seed(42); runif(3)  /  set.seed(1); runif(3)  /  set.seed(123); runif(2)
    assert np.allclose(RRandom(42).unif_rand(3), [0.914806, 0.9370754, 0.2861395], atol=5e-8)
    assert np.allclose(RRandom(1).unif_rand(3), [0.2655087, 0.3721239, 0.5728534], atol=5e-8)
    assert np.allclose(RRandom(123).unif_rand(2), [0.2875775, 0.7883051], atol=5e-8)


def test_r_sample_perm_known_answers():
    # R >= 3.6: set.seed(42); sample(1:10)  /  set.seed(123); sample(5)
    assert list(r_sample_perm(10, 42)) == [1, 5, 10, 8, 2, 4, 6, 9, 7, 3]
    assert list(r_sample_perm(5, 123)) == [3, 2, 5, 4, 1]
    for n in (1, 2, 17, 1063, 5000):
        p = r_sample_perm(n, 50)
        assert sorted(p) == list(range(1, n + 1))


@pytest.mark.parametrize("n", [1, 2, 10, 1063, 4097, 70000])
def test_native_sample_perm_matches_python(n):
    assert np.array_equal(_lib.r_sample_perm_native(n, 50), r_sample_perm(n, 50))


@pytest.mark.parametrize("m,p,seed", [(100, 7, 1), (2000, 60, 2154), (5000, 111, 2155)])
def test_native_ranm_matches_python(m, p, seed):
    a, b = ranM2(m, p, seed), _lib.r_ranm(m, p, seed)
    assert tuple(b["Dim"]) == (m, p)
    for k in ("p", "i", "x"):
        assert np.array_equal(a[k], b[k])


def test_ranm_structure():
    m, p = 4000, 50
    r = ranM2(m, p, 2154)
    s = np.sqrt(m)
    assert set(np.round(np.abs(r["x"]), 12)) == {round(np.sqrt(s), 12)}          # entries +-sqrt(sqrt(m))
    assert abs(len(r["i"]) / (m * p) - 1 / s) < 0.1 / s                              # density 1/sqrt(m)
    assert 0.4 < np.mean(r["x"] > 0) < 0.6                                          # signs balanced
    for j in range(p):                                                              # rows ascending inside a column
        seg = r["i"][r["p"][j]:r["p"][j + 1]]
        assert np.all(np.diff(seg) > 0)
    # R[i, j] = x0[(i-1)*p + j]: the draw sequence is ProbSampleReplace over (0, -sqrt(s), +sqrt(s)) cumulative
    u = RRandom(2154).unif_rand(m * p)
    x0 = np.where(u <= 1 - 1 / s, 0.0, np.where(u <= 1 - 1 / (2 * s), -np.sqrt(s), np.sqrt(s))).reshape(m, p)
    dense = np.zeros((m, p))
    for j in range(p):
        dense[r["i"][r["p"][j]:r["p"][j + 1]], j] = r["x"][r["p"][j]:r["p"][j + 1]]
    assert np.array_equal(dense, x0)


# ------------------------------------------------------------------------------------------------- A.2  projection
def test_rp_project_matches_dense_algebra():
    m, n, p = 500, 40, 23
    x, _ = synth.make_expression(m, n, seed=11)
    rm = ranM2(m, p, 99)
    R = np.zeros((m, p))
    for j in range(p):
        R[rm["i"][rm["p"][j]:rm["p"][j + 1]], j] = rm["x"][rm["p"][j]:rm["p"][j + 1]]
    ref = ((1 / np.sqrt(p)) * R.T @ np.log2(x + 1)).T
    got = orc.rp_project(m, n, rm, dense=x, logkind=2)
    assert np.allclose(got, ref, rtol=1e-12, atol=1e-12)
    got_csc = orc.rp_project(m, n, rm, csc=synth.to_csc(x), logkind=2)
    assert np.array_equal(got, got_csc)
    # CPM normalisation (R/SHARP.R:113) and the cell subset E[, tind]
    cells = np.array([5, 0, 39, 7])
    cs = x.sum(0)
    ref2 = ((1 / np.sqrt(p)) * R.T @ np.log2(x[:, cells] / cs[cells] * 1e6 + 1)).T
    got2 = orc.rp_project(m, n, rm, dense=x, cells=cells, colsum=cs, logkind=2)
    assert np.allclose(got2, ref2, rtol=1e-12, atol=1e-12)
    # SHARP_fpart: log10 and round(., 1)
    got3 = orc.rp_project(m, n, rm, dense=x, logkind=10, round_digits=1)
    ref3 = np.round(((1 / np.sqrt(p)) * R.T @ np.log10(x + 1)).T, 1)
    assert np.mean(np.abs(got3 - ref3) > 1e-9) < 5e-3


# ------------------------------------------------------------------------------------------------- A.3  distance
def test_zscore_corrdist():
    X = np.random.default_rng(0).normal(size=(30, 12))
    z, d = orc.zscore_corrdist(X)
    zr = (X - X.mean(1, keepdims=True)) / X.std(1, ddof=1, keepdims=True)  # scale(): sd with p-1
    assert np.allclose(z, zr, atol=1e-13)
    assert np.allclose(d, 1 - np.corrcoef(X), atol=1e-13)
    assert np.all(np.diag(d) == 0)


# ------------------------------------------------------------------------------------------------- A.4  hclust
def test_hclust_ward_hand_example():
    # 4 points on a line at 0, 1, 3, 7 -> d = |xi - xj|.  ward.D (Lance-Williams on the UNSQUARED d):
    # step 1: (1,2) at 1; d({12},3) = (2*3 + 2*2 - 1*1)/3 = 3, d({12},4) = (2*7 + 2*6 - 1)/3 = 25/3, d(3,4) = 4
    # step 2: ({12},3) at 3; d({123},4) = ((2+1)*25/3 + (1+1)*4 - 1*3)/4 = 7.5
    x = np.array([0.0, 1.0, 3.0, 7.0])
    d = np.abs(x[:, None] - x[None, :])
    ia, ib, h = orc.hclust(d, orc.WARD_D)
    assert list(ia) == [1, 1, 1] and list(ib) == [2, 3, 4]
    assert np.allclose(h, [1.0, 3.0, 7.5])
    ia, ib, h = orc.hclust(d, orc.SINGLE)
    assert np.allclose(h, [1.0, 2.0, 4.0])
    ia, ib, h = orc.hclust(d, orc.COMPLETE)
    assert np.allclose(h, [1.0, 3.0, 7.0])
    ia, ib, h = orc.hclust(d, orc.AVERAGE)
    assert np.allclose(h, [1.0, 2.5, (7 + 6 + 4) / 3])


def test_hclust_tie_break_first_minimum():
    # all distances equal: hclust.f takes the first strict minimum over i, NN to the right ->
    # (1,2), then ({12},3) ... with ward.D heights 1, 1, 1 for the all-ones matrix? LW: ((1+1)*1+(1+1)*1-1*1)/3 = 1
    n = 5
    d = np.ones((n, n)) - np.eye(n)
    ia, ib, h = orc.hclust(d, orc.WARD_D)
    assert list(ia) == [1, 1, 1, 1] and list(ib) == [2, 3, 4, 5]
    # block structure with exact ties (what wMetaC's S matrices look like): two groups of identical clusters
    S = np.kron(np.eye(2), np.ones((3, 3)))
    ia, ib, h = orc.hclust(1 - S, orc.WARD_D)
    assert np.allclose(h[:4], 0.0) and h[4] > 0
    lab = orc.cutree_k(ia, ib, 2)
    assert list(lab) == [1, 1, 1, 2, 2, 2]


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_hclust_ward_d_vs_scipy(seed):
    # R's ward.D on d == scipy 'ward' on sqrt(d) with heights squared (continuous data: no ties)
    X = np.random.default_rng(seed).normal(size=(60, 9))
    _, d = orc.zscore_corrdist(X)
    ia, ib, h = orc.hclust(d, orc.WARD_D)
    Z = sch.linkage(squareform(np.sqrt(d), checks=False), method="ward")
    assert np.allclose(np.sort(h), np.sort(Z[:, 2] ** 2), rtol=1e-10)
    for k in (2, 3, 5, 8):
        a = orc.cutree_k(ia, ib, k)
        b = sch.fcluster(Z, k, criterion="maxclust")
        assert adjusted_rand_score(a, b) == 1.0
    # ward.D2 == scipy ward on d itself
    ia2, ib2, h2 = orc.hclust(d, orc.WARD_D2)
    Z2 = sch.linkage(squareform(d, checks=False), method="ward")
    assert np.allclose(np.sort(h2), np.sort(Z2[:, 2]), rtol=1e-10)


@pytest.mark.parametrize("method,name", [(orc.SINGLE, "single"), (orc.COMPLETE, "complete"), (orc.AVERAGE, "average"),
                                         (orc.MCQUITTY, "weighted")])
def test_hclust_other_methods_vs_scipy(method, name):
    X = np.random.default_rng(5).normal(size=(40, 6))
    _, d = orc.zscore_corrdist(X)
    _, _, h = orc.hclust(d, method)
    Z = sch.linkage(squareform(d, checks=False), method=name)
    assert np.allclose(np.sort(h), np.sort(Z[:, 2]), rtol=1e-10)


# ------------------------------------------------------------------------------------------------- A.5  cutree
def test_cutree_first_appearance_numbering():
    # points 0,10,0.1,10.1,5 -> merges (1,3), (2,4), then 5 joins; cutree ids follow first appearance (obs 1 -> 1)
    x = np.array([0.0, 10.0, 0.1, 10.1, 5.0])
    d = np.abs(x[:, None] - x[None, :])
    ia, ib, _ = orc.hclust(d, orc.WARD_D)
    assert list(orc.cutree_k(ia, ib, 5)) == [1, 2, 3, 4, 5]
    assert list(orc.cutree_k(ia, ib, 3)) == [1, 2, 1, 2, 3]
    assert list(orc.cutree_k(ia, ib, 1)) == [1, 1, 1, 1, 1]


# ------------------------------------------------------------------------------------------------- A.6  silhouette
def test_silhouette_vs_sklearn_and_median():
    X = np.random.default_rng(3).normal(size=(51, 5)) + np.repeat(np.arange(3), 17)[:, None] * 2
    _, d = orc.zscore_corrdist(X)
    ia, ib, _ = orc.hclust(d, orc.WARD_D)
    for k in (2, 3, 6):
        lab = orc.cutree_k(ia, ib, k)
        sil, med = orc.silhouette_median(d, lab, k)
        ref = silhouette_samples(d, lab, metric="precomputed")
        assert np.allclose(sil, ref, atol=1e-12)
        assert np.isclose(med, np.median(ref), atol=1e-12)
    # even n: mean of the two middle order statistics; singleton cluster -> width 0
    d4 = np.array([[0, 1, 5, 9], [1, 0, 5, 9], [5, 5, 0, 9], [9, 9, 9, 0.0]])
    sil, med = orc.silhouette_median(d4, np.array([1, 1, 2, 3], dtype=np.int32), 3)
    assert sil[2] == 0 and sil[3] == 0
    assert np.allclose(sil[:2], [(5 - 1) / 5, (5 - 1) / 5])
    assert np.isclose(med, 0.4)


# ------------------------------------------------------------------------------------------------- A.7  CH
def test_ch_index_vs_sklearn_on_standardised_rows():
    X = np.random.default_rng(4).normal(size=(40, 7)) + np.repeat(np.arange(4), 10)[:, None]
    lab = np.repeat(np.arange(1, 5), 10).astype(np.int32)
    z = (X - X.mean(1, keepdims=True)) / X.std(1, ddof=1, keepdims=True)  # "1-corr": rows standardised first
    assert np.isclose(orc.get_ch(X, lab, 4), calinski_harabasz_score(z, lab), rtol=1e-10)


# ------------------------------------------------------------------------------------------------- get_opt_hclust
def planted(n_per, k, p, seed, sep=6.0):
    rng = np.random.default_rng(seed)
    cen = rng.normal(size=(k, p)) * sep
    return np.concatenate([cen[c] + rng.normal(size=(n_per, p)) for c in range(k)]), np.repeat(np.arange(k), n_per)


def test_opt_hclust_picks_planted_k_by_silhouette():
    X, truth = planted(25, 4, 30, 0)
    r = orc.opt_hclust(X, 0, orc.hc_params())
    assert r["optN.cluster"] == 4 and adjusted_rand_score(r["f"], truth) == 1.0
    assert r["maxsil"] == r["msil"].max() > 0.35
    assert r["v"].shape == (100, 39) and np.array_equal(r["v"][:, r["oind"] - 1], r["f"])
    assert np.all(np.diff(r["height"]) >= 0)


def test_opt_hclust_fixed_k_and_errors():
    X, _ = planted(10, 3, 8, 1)
    r = orc.opt_hclust(X, 0, orc.hc_params(n_cluster=5))
    assert r["optN.cluster"] == 5 and len(np.unique(r["f"])) == 5 and r["v"].shape[1] == 1
    with pytest.raises(orc.OracleError):
        orc.opt_hclust(X, 0, orc.hc_params(n_cluster=1))  # "The given N.cluster is less than 2 ..."


def test_opt_hclust_ch_and_height_fallback():
    # pure noise: max median silhouette <= sil.thre -> CH index decides (R/get_opt_hclust.R:194-217)
    X = np.random.default_rng(7).normal(size=(80, 20))
    r = orc.opt_hclust(X, 0, orc.hc_params())
    assert r["maxsil"] <= 0.35
    k_ch = 2 + int(np.argmax(r["CHind"]))
    assert r["optN.cluster"] == k_ch or np.argmax(r["CHind"]) == 0  # height-gap rule only when which.max(CH) == 1


def test_opt_hclust_symmetric_branch_and_middle_tie():
    # three groups of four identical clusters: S is 0/1.  k = 3 is the "right" answer, but the median silhouette is
    # exactly 1 for k = 3..6 (splitting singletons off ONE group zeroes only 4 of the 12 widths), and the reference
    # takes the MIDDLE index among the maxima, tmp[ceiling(length(tmp)/2)] (R/get_opt_hclust.R:162-168) -> k = 4
    S = np.kron(np.eye(3), np.ones((4, 4)))
    r = orc.opt_hclust(S, 1, orc.hc_params(max_n=8))
    msil = r["msil"]
    assert np.allclose(msil, [3 / 7, 1, 1, 1, 1, 0, 0])
    ties = np.flatnonzero(msil == msil.max())
    assert list(ties) == [1, 2, 3, 4]
    assert r["oind"] - 1 == ties[(len(ties) + 1) // 2 - 1] == 2
    assert r["optN.cluster"] == 4
    assert list(r["f"]) == [1] * 4 + [2] * 4 + [3] * 3 + [4]


# ------------------------------------------------------------------------------------------------- wMetaC / sMetaC
def test_wmetac_consensus_of_identical_solutions():
    truth = np.repeat(np.arange(1, 5), 30)
    rng = np.random.default_rng(0)
    cols = []
    for c in range(5):  # the same partition under 5 different label permutations
        perm = rng.permutation(4) + 1
        cols.append(perm[truth - 1])
    r = orc.wmetac(np.stack(cols, 1), orc.hc_params())
    assert r["N.cluster"] == 4 and adjusted_rand_score(r["finalC"], truth) == 1.0
    # identical solutions: AA is 0/1 -> nd = 0 -> w0 = 0 -> w1 = 0.01/1.01 for every cell
    assert np.allclose(r["w1"], 0.01 / 1.01)
    # x0: 1 for the chosen cluster, 0 elsewhere when all members agree
    assert r["x0"].shape == (120, 4) and np.array_equal(r["x0"].sum(1), np.ones(120))


def test_wmetac_majority_vote_with_noise():
    rng = np.random.default_rng(1)
    truth = np.repeat(np.arange(1, 4), 50)
    cols = []
    for c in range(7):
        lab = truth.copy()
        flip = rng.random(150) < 0.08
        lab[flip] = rng.integers(1, 4, flip.sum())
        cols.append(lab)
    r = orc.wmetac(np.stack(cols, 1), orc.hc_params())
    assert r["N.cluster"] == 3 and adjusted_rand_score(r["finalC"], truth) > 0.97
    assert np.all((r["w1"] > 0) & (r["w1"] <= 1))
    assert np.all(r["x0"].max(1) == 1.0) and np.all(r["x0"] >= 0)


def test_smetac_merges_blocks_by_centroid_correlation():
    rng = np.random.default_rng(2)
    p, per = 40, 30
    cen = rng.normal(size=(3, p)) * 3
    truth = np.tile(np.repeat(np.arange(3), per), 2)          # two blocks, the same 3 types in each
    E = cen[truth] + rng.normal(size=(len(truth), p))
    block = np.repeat([0, 1], 3 * per)
    labels = block * 10 + truth                                # block-local cluster names
    r = orc.smetac(labels, E, orc.hc_params())
    assert len(r["tf"]) == 6
    assert adjusted_rand_score(r["finalColor"], truth) == 1.0


# ------------------------------------------------------------------------------------------------- whole path
def test_sharp_large_recovers_planted_types():
    x, truth = synth.make_expression(2000, 900, n_types=4, seed=1, sep=2.5, frac=0.5)
    K, p = 5, 120
    rms = [ranM2(2000, p, 50 + 2103 + k) for k in range(1, K + 1)]
    prm = orc.SharpParams(1, 1, K, p, 300, 0, 0, 0, orc.hc_params(), 2, -1)
    r = orc.sharp(2000, 900, rms, prm, dense=x, reind=r_sample_perm(900, 50))
    assert synth.ari(r["pred_clusters"], truth) > 0.9
    assert r["viE"].shape == (900, p) and r["x0"].shape[0] == 900
    # dense and dgCMatrix inputs give the same answer
    r2 = orc.sharp(2000, 900, rms, prm, csc=synth.to_csc(x), reind=r_sample_perm(900, 50))
    assert np.array_equal(r["pred_clusters"], r2["pred_clusters"])


def test_ari_helper_vs_sklearn():
    rng = np.random.default_rng(0)
    a, b = rng.integers(0, 5, 300), rng.integers(0, 4, 300)
    assert np.isclose(synth.ari(a, b), adjusted_rand_score(a, b), atol=1e-12)
