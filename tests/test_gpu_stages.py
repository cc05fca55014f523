"""GPU parity tests, stage by stage: every C-ABI entry point of libsharpb200 against the CPU oracle on the same
seeded inputs.  Integer artefacts (merge sequences, labels) must be identical; floating point within the
tolerance written in each test (the projection's contract is 1e-5 relative, BASELINE.json north_star)."""
import numpy as np
import pytest

import orc
import synth
from sharp_b200 import Context, RStop, _lib, hc_params
from sharp_b200.rrng import ranM2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = Context(0)
    yield c
    c.close()


def orc_prm(p):
    return orc.hc_params(p.hmethod, p.n_cluster, p.min_n, p.max_n, p.sil_thre, p.height_ntimes)


def relerr(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


# --------------------------------------------------------------------------------------------- K1
@pytest.mark.parametrize("fmt", ["dense", "csc"])
@pytest.mark.parametrize("normalize,logkind", [(0, 2), (1, 2), (2, 2), (0, 0), (0, 10)])
def test_rp_project(ctx, fmt, normalize, logkind):
    m, n, K, p = 3000, 257, 3, 61
    kind = "umi" if normalize else "tpm"
    x, _ = synth.make_expression(m, n, seed=3, zero_frac=0.8, kind=kind)
    rms = [ranM2(m, p, 2154 + k) for k in range(K)]
    rm = ctx.upload_rm(rms)
    colsum = x.sum(0)
    kw = dict(dense=x) if fmt == "dense" else dict(csc=synth.to_csc(x))
    got = ctx.rp_project(m, n, rm, normalize=normalize, colsum=colsum if normalize == 1 else None, logkind=logkind, **kw)
    for k in range(K):
        ref = orc.rp_project(m, n, rms[k], colsum=colsum if normalize else None, logkind=logkind, **kw)
        assert relerr(got[k], ref) <= 1e-5, (k, relerr(got[k], ref))
        assert relerr(got[k], ref) <= 1e-11  # what fp64 accumulation should actually deliver


@pytest.mark.parametrize("fmt", ["dense", "csc"])
@pytest.mark.parametrize("logkind", [2, 0])
def test_rp_project_both_variants_agree(ctx, fmt, logkind):
    """the fixed-point atomic kernel (default) and the fp64 gene-order kernel against the oracle and each other"""
    m, n, K, p = 5000, 333, 5, 97
    x, _ = synth.make_expression(m, n, seed=8, zero_frac=0.9, kind="umi")
    x[:, 7] = 0.0
    x[11, 7] = 3.0                      # a cell with a single non-zero gene
    x[5, 9] = 1e5                       # a wide dynamic range inside one cell
    rms = [ranM2(m, p, 300 + k) for k in range(K)]
    rm = ctx.upload_rm(rms)
    kw = dict(dense=x) if fmt == "dense" else dict(csc=synth.to_csc(x))
    out = {}
    for variant in (0, 1, 2, 3):
        ctx.set_rp_variant(variant)
        out[variant] = ctx.rp_project(m, n, rm, normalize=2, logkind=logkind, **kw)
    ctx.set_rp_variant(0)
    again = ctx.rp_project(m, n, rm, normalize=2, logkind=logkind, **kw)
    assert np.array_equal(again, out[0])          # integer atomics: bit-reproducible whatever the order
    # the record-gather kernel (0), its TMA-staged form (3) and the round-1 fixed-point kernel (2) form the same integers
    assert np.array_equal(out[0], out[2]) and np.array_equal(out[0], out[3])
    out = {False: out[0], True: out[1]}
    refs = np.stack([orc.rp_project(m, n, rms[k], colsum=x.sum(0), logkind=logkind, **kw) for k in range(K)])
    # The fp64 kernel is accurate relative to every member's own projection.  The fixed-point kernel's quantum is set
    # per CELL from its largest transformed value (2^-55 of it here), so its error is measured against the cell's
    # largest output over all members (cell 9, whose 1e5 count dwarfs its other genes, is the case in point).
    per_member = np.maximum(np.max(np.abs(refs), axis=2, keepdims=True), 1e-300)
    per_cell = np.max(per_member, axis=0, keepdims=True)
    assert np.max(np.abs(out[True] - refs) / per_member) <= 1e-13
    assert np.max(np.abs(out[False] - refs) / per_cell) <= 1e-14
    assert np.max(np.abs(out[False] - refs) / per_member) <= 1e-11


def test_rp_project_count_classes_exact(ctx):
    """count data: non-zeros equal to 1..4 are counted per output (+/- fields) instead of added, everything else takes
    the two-limb generic path.  Both are the same exact integer sum, so the CSC kernel (class passes + compacted generic
    passes) and the dense kernel (mixed lanes) must agree BITWISE, and both with the oracle and the fp64 kernel to
    rounding.  Explicit zeros stored in the dgCMatrix and an empty cell contribute nothing."""
    m, n, K, p = 4000, 120, 5, 101
    rng = np.random.default_rng(12)
    x = rng.poisson(0.25, size=(m, n)).astype(np.float64)
    x[rng.random((m, n)) < 0.01] = 7.0          # some generic integers
    x[rng.random((m, n)) < 0.005] = 2.5         # and non-integers
    x[:, 3] = 0.0                               # an empty cell
    x[:, 5] = 0.0
    x[17, 5] = 1.0                              # a cell with one class-1 gene
    cp, ri, xv = synth.to_csc(x)
    # explicit zeros: append a stored 0.0 to the first cell's column
    free = int(np.setdiff1d(np.arange(m), ri[cp[0]:cp[1]])[0])
    ri = np.concatenate([ri[:cp[1]], [free], ri[cp[1]:]]).astype(np.int32)
    xv = np.concatenate([xv[:cp[1]], [0.0], xv[cp[1]:]])
    cp = cp.copy()
    cp[1:] += 1
    order = np.argsort(ri[cp[0]:cp[1]], kind="stable")
    ri[cp[0]:cp[1]] = ri[cp[0]:cp[1]][order]
    xv[cp[0]:cp[1]] = xv[cp[0]:cp[1]][order]
    rms = [ranM2(m, p, 900 + k) for k in range(K)]
    rm = ctx.upload_rm(rms)
    colsum = x.sum(0)
    colsum[3] = 1.0                             # the reference divides by colSums; keep the empty cell finite
    for logkind in (2, 0):
        got_csc = ctx.rp_project(m, n, rm, csc=(cp, ri, xv), normalize=1, colsum=colsum, logkind=logkind)
        got_dense = ctx.rp_project(m, n, rm, dense=x, normalize=1, colsum=colsum, logkind=logkind)
        assert np.array_equal(got_csc, got_dense)   # class counting, generic passes and the mixed dense path: same integers
        ctx.set_rp_variant(True)
        legacy = ctx.rp_project(m, n, rm, dense=x, normalize=1, colsum=colsum, logkind=logkind)
        ctx.set_rp_variant(False)
        for k in range(K):
            ref = orc.rp_project(m, n, rms[k], dense=x, colsum=colsum, logkind=logkind)
            assert relerr(got_csc[k], ref) <= 1e-12 and relerr(legacy[k], ref) <= 1e-12
        assert np.all(got_csc[:, 3, :] == 0.0)


@pytest.mark.parametrize("variant", [0, 3])
def test_rp_project_record_overflow_and_staging_limits(ctx, variant):
    """record-gather kernel: genes whose entry list overflows the fixed-size record (served from the CSR lists), cells
    with more non-zeros than one staged buffer holds (2048) and cells starting at every 16-byte misalignment"""
    m, K, p = 1200, 4, 300                      # ~35 entries per gene: 8-vector records, some genes overflow 56
    rng = np.random.default_rng(21)
    n = 37
    x = np.zeros((m, n))
    for c in range(n):
        k = [0, 1, 2, 3, 5, 1199, 1200][c % 7] if c < 14 else int(rng.integers(5, 900))
        rows = rng.choice(m, size=min(k, m), replace=False)
        x[rows, c] = rng.integers(1, 9, size=len(rows)).astype(np.float64)
    x[:, 20] = rng.integers(1, 4, size=m)       # dense cells: more non-zeros than... (m < 2048: see the wide matrix below)
    rms = [ranM2(m, p, 700 + k) for k in range(K)]
    per_gene = np.zeros(m, dtype=int)
    for r in rms:
        per_gene += np.bincount(r["i"], minlength=m)
    rm = ctx.upload_rm(rms)
    ctx.set_rp_variant(variant)
    try:
        got = ctx.rp_project(m, n, rm, csc=synth.to_csc(x), normalize=0, logkind=2)
        # a wide matrix: cells with > 2048 non-zeros
        m2 = 6000
        x2 = np.zeros((m2, 9))
        for c in range(9):
            rows = rng.choice(m2, size=[2047, 2048, 2049, 4096, 5000, 1, 2050, 3000, 6000][c], replace=False)
            x2[rows, c] = rng.integers(1, 30, size=len(rows)).astype(np.float64)
        rms2 = [ranM2(m2, 64, 800 + k) for k in range(3)]
        rm2 = ctx.upload_rm(rms2)
        got2 = ctx.rp_project(m2, 9, rm2, csc=synth.to_csc(x2), normalize=2, logkind=2)
    finally:
        ctx.set_rp_variant(0)
    assert per_gene.max() > 56                  # two genes overflow the 8-vector records: the CSR path is exercised
    for k in range(K):
        assert relerr(got[k], orc.rp_project(m, n, rms[k], dense=x, logkind=2)) <= 1e-12
    for k in range(3):
        assert relerr(got2[k], orc.rp_project(m2, 9, rms2[k], dense=x2, colsum=x2.sum(0), logkind=2)) <= 1e-12


def test_rp_project_long_ranm_columns_use_16bit_fields(ctx):
    """a ranM column with more than 255 entries (m > 65 000 genes) cannot be counted in 8-bit fields: the kernel then
    counts classes 1 and 2 in 16-bit fields; m = 90 000, all counts 1 or 2 in a few dense cells so that single outputs
    really receive > 255 terms of one class"""
    m, n, p = 90000, 6, 12
    rng = np.random.default_rng(4)
    x = rng.integers(1, 3, size=(m, n)).astype(np.float64)   # every gene expressed: 1 or 2
    x[:, 4] = rng.integers(0, 6, size=m)                      # one ordinary cell
    rms = [ranM2(m, p, 41), ranM2(m, p, 42)]
    assert max(int(np.max(np.diff(r["p"]))) for r in rms) > 255
    rm = ctx.upload_rm(rms)
    got = ctx.rp_project(m, n, rm, csc=synth.to_csc(x), normalize=2, logkind=2)
    for k in range(2):
        ref = orc.rp_project(m, n, rms[k], dense=x, colsum=x.sum(0), logkind=2)
        assert relerr(got[k], ref) <= 1e-12


def test_rp_project_nonfinite_value_poisons_the_cell(ctx):
    m, n, p = 800, 20, 16
    x, _ = synth.make_expression(m, n, seed=9)
    x[3, 4] = -5.0                      # log2(-5 + 1) = NaN in the reference too
    rm = ctx.upload_rm([ranM2(m, p, 5)])
    got = ctx.rp_project(m, n, rm, dense=x, logkind=2)[0]
    assert np.all(np.isnan(got[4])) and np.all(np.isfinite(np.delete(got, 4, axis=0)))


def test_rp_project_cells_and_round(ctx):
    m, n, K, p = 1500, 100, 2, 40
    x, _ = synth.make_expression(m, n, seed=5)
    rms = [ranM2(m, p, 77 + k) for k in range(K)]
    rm = ctx.upload_rm(rms)
    cells = np.random.default_rng(0).permutation(n)[:63]
    got = ctx.rp_project(m, n, rm, csc=synth.to_csc(x), cells=cells)
    for k in range(K):
        ref = orc.rp_project(m, n, rms[k], csc=synth.to_csc(x), cells=cells)
        assert relerr(got[k], ref) <= 1e-11
    got = ctx.rp_project(m, n, rm, dense=x, logkind=10, round_digits=1)
    ref = orc.rp_project(m, n, rms[0], dense=x, logkind=10, round_digits=1)
    # rounding to one decimal can flip on a last-bit difference of the unrounded value: allow a few
    assert np.mean(np.abs(got[0] - ref) > 1e-9) < 1e-3


def test_rp_project_rejects_non_ternary(ctx):
    rm = ranM2(100, 10, 1)
    rm["x"] = rm["x"].copy()
    rm["x"][0] *= 2
    with pytest.raises(_lib.SharpError):
        ctx.upload_rm([rm])


# --------------------------------------------------------------------------------------------- K2
@pytest.mark.parametrize("n,p", [(50, 7), (300, 61), (517, 333)])
def test_corrdist(ctx, n, p):
    X = np.random.default_rng(n).normal(size=(n, p)) + np.arange(p) * 0.01
    d = ctx.corrdist(X)
    _, ref = orc.zscore_corrdist(X)
    assert np.all(np.diag(d) == 0.0)
    assert np.array_equal(d, d.T)
    assert np.max(np.abs(d - ref)) <= 1e-12


# --------------------------------------------------------------------------------------------- K3
@pytest.mark.parametrize("method", [1, 2, 3, 4, 5, 6, 7, 8])
def test_hclust_exact_order(ctx, method):
    rng = np.random.default_rng(method)
    X = rng.normal(size=(150, 9))
    _, d = orc.zscore_corrdist(X)
    ia, ib, h = ctx.hclust(d, method)
    ria, rib, rh = orc.hclust(d, method)
    assert np.array_equal(ia, ria) and np.array_equal(ib, rib)
    assert np.array_equal(h, rh)  # same operations in the same order: bit-identical heights


def test_hclust_ties(ctx):
    """similarity-like matrix full of exact ties (0 / 0.5 / 1): the merge order must be hclust.f's."""
    rng = np.random.default_rng(9)
    n = 90
    d = rng.choice([0.0, 0.5, 1.0], size=(n, n), p=[0.1, 0.2, 0.7])
    d = np.triu(d, 1)
    d = d + d.T
    for method in (1, 3, 4):
        ia, ib, h = ctx.hclust(d, method)
        ria, rib, rh = orc.hclust(d, method)
        assert np.array_equal(ia, ria) and np.array_equal(ib, rib) and np.array_equal(h, rh)


def test_hclust_large(ctx):
    rng = np.random.default_rng(11)
    X = rng.normal(size=(1203, 40))
    X[:400] += 1.5
    _, d = orc.zscore_corrdist(X)
    ia, ib, h = ctx.hclust(d, 1)
    ria, rib, rh = orc.hclust(d, 1)
    assert np.array_equal(ia, ria) and np.array_equal(ib, rib) and np.array_equal(h, rh)


# --------------------------------------------------------------------------------------------- K4-K6
def _blobs(n, p, g, seed, sep=2.0):
    rng = np.random.default_rng(seed)
    lab = rng.integers(0, g, n)
    cen = rng.normal(size=(g, p)) * sep
    return cen[lab] + rng.normal(size=(n, p)), lab


@pytest.mark.parametrize("exact", [False, True])
@pytest.mark.parametrize("n,p,g,sep", [(400, 50, 5, 2.0), (333, 31, 3, 0.3), (45, 20, 2, 1.0)])
def test_opt_hclust_feature(ctx, exact, n, p, g, sep):
    X, _ = _blobs(n, p, g, seed=n, sep=sep)
    prm = hc_params()
    got = ctx.opt_hclust(X, False, prm, exact=exact)
    ref = orc.opt_hclust(X, 0, orc_prm(prm))
    assert got["v"].shape == ref["v"].shape
    assert np.array_equal(got["v"], ref["v"])
    assert np.allclose(got["height"], ref["height"], rtol=1e-10, atol=1e-13)
    assert np.allclose(got["msil"], ref["msil"], rtol=0, atol=1e-12)
    assert np.allclose(got["CHind"], ref["CHind"], rtol=1e-8)
    assert got["oind"] == ref["oind"] and got["optN.cluster"] == ref["optN.cluster"]
    assert np.array_equal(got["f"], ref["f"])
    assert abs(got["maxsil"] - ref["maxsil"]) <= 1e-12


@pytest.mark.parametrize("n,p,g,sep,method", [(1203, 64, 6, 1.5, "ward.D"), (2000, 120, 9, 1.0, "ward.D"), (900, 40, 4, 2.0, "ward.D2"),
                                              (700, 30, 3, 1.0, "average"), (650, 30, 3, 1.0, "complete")])
def test_opt_hclust_round_parallel(ctx, n, p, g, sep, method):
    """n > 384 feature problems run the round-parallel (reciprocal nearest neighbour) agglomeration: identical
    dendrogram / cuts / selection, heights equal to rounding"""
    X, _ = _blobs(n, p, g, seed=n, sep=sep)
    prm = hc_params(hmethod=method)
    got = ctx.opt_hclust(X, False, prm)
    ref = orc.opt_hclust(X, 0, orc_prm(prm))
    assert np.array_equal(got["v"], ref["v"])
    assert np.allclose(got["height"], ref["height"], rtol=1e-10, atol=1e-13)
    assert np.allclose(got["msil"], ref["msil"], rtol=0, atol=1e-12)
    assert got["oind"] == ref["oind"] and got["optN.cluster"] == ref["optN.cluster"]
    assert np.array_equal(got["f"], ref["f"])


@pytest.mark.parametrize("n,max_n", [(1100, 60), (901, 64), (500, 41)])
def test_opt_hclust_wide_k_range(ctx, n, max_n):
    """maxN.cluster above 40 (R/SHARP.R:216-223 for parts over 200 000 cells): the nested sweep runs its 256-thread
    layout (the cached means of 60+ clusters do not fit beside 512 threads); every level against the oracle"""
    X, _ = _blobs(n, 36, 7, seed=n + max_n, sep=1.2)
    prm = hc_params(max_n=max_n)
    got = ctx.opt_hclust(X, False, prm)
    ref = orc.opt_hclust(X, 0, orc_prm(prm))
    assert got["v"].shape == ref["v"].shape == (n, max_n - 1)
    assert np.array_equal(got["v"], ref["v"])
    assert np.allclose(got["msil"], ref["msil"], rtol=0, atol=1e-12)
    assert np.allclose(got["CHind"], ref["CHind"], rtol=1e-8)
    assert got["oind"] == ref["oind"] and got["optN.cluster"] == ref["optN.cluster"]
    assert np.array_equal(got["f"], ref["f"])


def test_opt_hclust_round_parallel_ties_fall_back(ctx):
    """duplicated cells give exact ties (d = 0, equal rows): the round-parallel kernel must hand the problem to the
    exact kernel, and the result is still hclust.f's"""
    X, _ = _blobs(300, 25, 4, seed=5, sep=2.0)
    X = np.concatenate([X, X[:150], X[:40]], 0)  # n = 490 > 384 with duplicate and triplicate rows
    prm = hc_params()
    got = ctx.opt_hclust(X, False, prm)
    ref = orc.opt_hclust(X, 0, orc_prm(prm))
    assert np.array_equal(got["v"], ref["v"]) and np.array_equal(got["f"], ref["f"])
    assert np.allclose(got["height"], ref["height"], rtol=1e-10, atol=1e-13)


def test_opt_hclust_ch_and_height_paths(ctx):
    """weak structure -> max(msil) <= sil.thre -> CH index, possibly the height-gap rule"""
    for seed, sil in [(1, 0.9), (2, 0.9), (3, 2.0)]:
        X, _ = _blobs(260, 40, 4, seed=seed, sep=0.6)
        prm = hc_params(sil_thre=sil, height_ntimes=1.05)
        try:
            ref = orc.opt_hclust(X, 0, orc_prm(prm))
        except orc.OracleError as e:
            with pytest.raises(RStop):
                ctx.opt_hclust(X, False, prm)
            continue
        got = ctx.opt_hclust(X, False, prm)
        assert got["oind"] == ref["oind"]
        assert np.array_equal(got["f"], ref["f"])


def test_opt_hclust_fixed_k(ctx):
    X, _ = _blobs(300, 30, 4, seed=8)
    prm = hc_params(n_cluster=6)
    got = ctx.opt_hclust(X, False, prm)
    ref = orc.opt_hclust(X, 0, orc_prm(prm))
    assert np.array_equal(got["f"], ref["f"]) and got["optN.cluster"] == 6
    assert abs(got["msil"][0] - ref["msil"][0]) <= 1e-12


def test_opt_hclust_symmetric_bit_exact(ctx):
    """similarity branch: the sweep follows the reference's summation order -> msil identical to the last bit"""
    rng = np.random.default_rng(4)
    lab = np.stack([rng.integers(1, 6, 300) for _ in range(4)], 1)
    # a similarity matrix with plenty of exact structure: correlation of one-hot cluster indicators
    H = np.concatenate([(lab[:, [c]] == np.arange(1, 6)[None, :]).astype(float) for c in range(4)], 1)
    S = np.corrcoef(H.T)
    S = (S + S.T) / 2
    np.fill_diagonal(S, 1.0)
    prm = hc_params(max_n=15)
    got = ctx.opt_hclust(S, True, prm)
    ref = orc.opt_hclust(S, 1, orc_prm(prm))
    assert np.array_equal(got["v"], ref["v"])
    assert np.array_equal(got["height"], ref["height"])
    assert np.array_equal(got["msil"], ref["msil"])
    assert np.allclose(got["CHind"], ref["CHind"], rtol=1e-9)
    assert np.array_equal(got["f"], ref["f"])


def test_opt_hclust_too_few_points(ctx):
    X = np.random.default_rng(0).normal(size=(2, 10))
    with pytest.raises(RStop):
        ctx.opt_hclust(X, False, hc_params())


def test_getrowcolor(ctx):
    X, _ = _blobs(500, 40, 6, seed=21)
    prm = hc_params()
    col, ms = ctx.getrowcolor(X, prm)
    rcol, rms = orc.getrowcolor(X, orc_prm(prm))
    assert np.array_equal(col, rcol) and abs(ms - rms) <= 1e-12


# --------------------------------------------------------------------------------------------- K7-K9
def _ensemble_labels(N, C, g, noise, seed):
    rng = np.random.default_rng(seed)
    truth = rng.integers(0, g, N)
    cols = []
    for c in range(C):
        perm = rng.permutation(g + 2)
        l = perm[truth]
        flip = rng.random(N) < noise
        l = np.where(flip, rng.integers(0, g + 2, N), l)
        # relabel by first appearance, 1-based (what getrowColor produces)
        _, first = np.unique(l, return_index=True)
        order = np.argsort(first)
        m = {int(np.unique(l)[o]): i + 1 for i, o in enumerate(order)}
        cols.append(np.array([m[int(v)] for v in l]))
    return np.stack(cols, 1)


@pytest.mark.parametrize("N,C,g,noise", [(600, 5, 4, 0.1), (350, 15, 6, 0.3), (200, 3, 2, 0.0), (501, 5, 9, 0.5)])
def test_wmetac(ctx, N, C, g, noise):
    lab = _ensemble_labels(N, C, g, noise, seed=N + C)
    prm = hc_params()
    try:
        ref = orc.wmetac(lab, orc_prm(prm))
    except orc.OracleError:
        with pytest.raises(_lib.SharpError):
            ctx.wmetac(lab, prm)
        return
    got = ctx.wmetac(lab, prm)
    assert np.array_equal(got["w1"], ref["w1"])          # same sums in the same order
    assert np.array_equal(got["finalC"], ref["finalC"])
    assert got["N.cluster"] == ref["N.cluster"]
    assert np.array_equal(got["x0"], ref["x0"])


def test_wmetac_fixed_k_and_arbitrary_codes(ctx):
    lab = _ensemble_labels(400, 5, 5, 0.2, seed=2) * 7 + 100  # any integer coding is valid
    prm = hc_params(n_cluster=3)
    got = ctx.wmetac(lab, prm)
    ref = orc.wmetac(lab, orc_prm(prm))
    assert np.array_equal(got["finalC"], ref["finalC"])


# --------------------------------------------------------------------------------------------- K10
@pytest.mark.parametrize("ncells,p,nclu", [(3000, 50, 23), (12000, 64, 60)])
def test_smetac(ctx, ncells, p, nclu):
    rng = np.random.default_rng(ncells)
    groups = rng.integers(0, 6, nclu)
    cen = rng.normal(size=(6, p)) * 3
    lab = rng.integers(0, nclu, ncells)
    E = cen[groups[lab]] + rng.normal(size=(ncells, p))
    codes = lab * 13 + 5
    prm = hc_params()
    got = ctx.smetac(codes, E, prm)
    ref = orc.smetac(codes, E, orc_prm(prm))
    assert np.array_equal(got["tf"], ref["tf"])
    assert np.array_equal(got["finalColor"], ref["finalColor"])
