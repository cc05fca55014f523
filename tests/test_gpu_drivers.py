"""GPU parity of the exported DRIVERS and of the stages at the benchmark's own shapes: everything
goes through the C ABI (sharp_b200.api -> libsharpb200.so) and is compared with the CPU oracle / the literal R
transcriptions of tests/rtrans.py on the same seeded inputs.  Integer artefacts must be identical, projections within
1e-5 relative (the contract; measured ~1e-15)."""
import math
import os

import numpy as np
import pytest

import orc
import rtrans
import synth
from sharp_b200 import Context, RunParams, api, hc_params
from sharp_b200.rrng import r_sample_perm, ranM2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = Context(0)
    yield c
    c.close()


def relerr(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def first_appearance(y):
    _, idx, inv = np.unique(y, return_index=True, return_inverse=True)
    rank = np.empty(len(idx), dtype=np.int64)
    rank[np.argsort(idx)] = np.arange(1, len(idx) + 1)
    return rank[inv]


def glue(labels, n):
    """R/SHARP.R:816-832 on the raw device labels: merge of clusters below 10 cells (n > 1e4), match(y, unique(y))"""
    gl = np.asarray(labels).copy()
    if n > 10000:
        vals, cnt = np.unique(gl, return_counts=True)
        small = vals[cnt < 10]
        if len(small):
            gl[np.isin(gl, small)] = small.min()
    return first_appearance(gl)


# ---------------------------------------------------------------------------------------------------------
# BASELINE config 2 in full: 20 000 genes x 10 000 cells dense TPM, SHARP() -> SHARP_large, K = 5, p = 333
# ---------------------------------------------------------------------------------------------------------
def test_cfg2_full_dense_tpm_sharp_large(ctx):
    m, n, K = 20000, 10000, 5
    x, truth = synth.make_expression(m, n, n_types=8, seed=2103, kind="tpm", zero_frac=0.7, sep=1.5, frac=0.3)
    p = math.ceil(math.log2(n) / 0.04)
    assert p == 333
    seed = 2103
    got = api.SHARP(x, exp_type="TPM", rN_seed=seed, logflag=False, ctx=ctx)       # ncells >= 5000 -> SHARP_large, K = 5
    rms = [ranM2(m, p, 50 + seed + k) for k in range(1, K + 1)]
    reind = r_sample_perm(n, 50)
    oprm = orc.SharpParams(1, 1, K, p, 2000, 0, 0, 0, orc.hc_params(max_n=40), 2, -1)
    ref = orc.sharp(m, n, rms, oprm, dense=x, reind=reind)
    assert got["paras"]["reduced.ndim"] == 333 and got["ensize.K"] == 5
    assert np.array_equal(got["pred_clusters"], ref["pred_clusters"])
    assert synth.ari(got["pred_clusters"], ref["pred_clusters"]) == 1.0
    assert relerr(got["viE"], ref["viE"]) <= 1e-5
    assert np.allclose(got["viE"], ref["viE"], rtol=1e-9, atol=1e-11)
    assert got["x0"].shape == ref["x0"].shape and np.allclose(got["x0"], ref["x0"], rtol=0, atol=1e-12)
    assert synth.ari(got["pred_clusters"], truth) > 0.5


# ---------------------------------------------------------------------------------------------------------
# BASELINE config 3, a slice: UMI -> CPM fused, 15 members, p = 416 (K*p = 6240 accumulators per cell), 4 blocks
# ---------------------------------------------------------------------------------------------------------
def test_cfg3_slice_umi_cpm_15_members(ctx):
    m, n, K, p = 20000, 8000, 15, 416
    x, _ = synth.make_expression(m, n, n_types=10, seed=3, kind="umi", zero_frac=0.93, sep=1.5, frac=0.3)
    csc = synth.to_csc(x)
    colsum = x.sum(0)
    seed = 2103
    rms = [ranM2(m, p, 50 + seed + k) for k in range(1, K + 1)]
    reind = r_sample_perm(n, 50)
    rm = ctx.upload_rm(rms)
    proj = ctx.rp_project(m, n, rm, csc=csc, normalize=2, logkind=2)
    worst = 0.0
    for k in range(K):
        worst = max(worst, relerr(proj[k], orc.rp_project(m, n, rms[k], csc=csc, colsum=colsum, logkind=2)))
    assert worst <= 1e-5 and worst <= 1e-11
    hc = hc_params(max_n=40)
    got = ctx.run(rm, RunParams(1, 1, 2, -1, 2000, 0, 0, 0, hc, 2, 1e6), m=m, n=n, csc=csc, reind=reind)
    rm.close()
    oprm = orc.SharpParams(1, 1, K, p, 2000, 0, 0, 0, orc.hc_params(max_n=40), 2, -1)
    ref = orc.sharp(m, n, rms, oprm, csc=csc, colsum=colsum, reind=reind)
    assert np.array_equal(glue(got["labels"], n), ref["pred_clusters"])
    assert np.allclose(got["viE"], ref["viE"], rtol=1e-9, atol=1e-11)


# ---------------------------------------------------------------------------------------------------------
# BASELINE config 4's own block shape: m = 27 998, p = 508, K = 5, 2 blocks of 2000 (the projection and the fused run)
# ---------------------------------------------------------------------------------------------------------
def test_cfg4_block_shape(ctx):
    m, n, K, p = 27998, 4000, 5, 508
    x, _ = synth.make_expression(m, n, n_types=6, seed=4, kind="umi", zero_frac=0.93, sep=1.5, frac=0.3)
    csc = synth.to_csc(x)
    colsum = x.sum(0)
    rms = [ranM2(m, p, 50 + 2103 + k) for k in range(1, K + 1)]
    reind = r_sample_perm(n, 50)
    rm = ctx.upload_rm(rms)
    got = ctx.run(rm, RunParams(1, 1, 2, -1, 2000, 0, 0, 0, hc_params(max_n=40), 2, 1e6), m=m, n=n, csc=csc, reind=reind)
    rm.close()
    oprm = orc.SharpParams(1, 1, K, p, 2000, 0, 0, 0, orc.hc_params(max_n=40), 2, -1)
    ref = orc.sharp(m, n, rms, oprm, csc=csc, colsum=colsum, reind=reind)
    assert np.array_equal(glue(got["labels"], n), ref["pred_clusters"])
    assert relerr(got["viE"], ref["viE"]) <= 1e-11


# ---------------------------------------------------------------------------------------------------------
# sMetaC at >= 1e6 cells: the k-range branch the benchmark runs (R/sMetaC.R:110-119)
# ---------------------------------------------------------------------------------------------------------
def test_smetac_million_cell_k_range(ctx):
    rng = np.random.default_rng(7)
    ncells, p, nclu, g = 1_000_000, 24, 250, 30
    centres = rng.normal(size=(g, p)) * 2.0
    which = np.arange(nclu) % g
    # centroid values on a 2^-20 grid and clusters of at most 8192 rows: every colMeans() is exact, so the centroid
    # entry point (what SHARP_unlimited's global step calls) and the rows entry point see identical numbers
    cen = np.round((centres[which] + 0.25 * rng.normal(size=(nclu, p))) * 2 ** 20) / 2 ** 20
    sizes = np.full(nclu, ncells // nclu)
    labels = np.repeat(np.arange(1, nclu + 1), sizes).astype(np.int32)
    se1 = np.ascontiguousarray(cen[labels - 1])
    prm = orc.hc_params(max_n=max(40, ncells // 5000))
    ref = orc.smetac(labels, se1, prm)
    hp = hc_params(max_n=max(40, ncells // 5000))
    tf = ctx.smetac_centroids(cen, ncells, hp)
    assert np.array_equal(tf, ref["tf"])
    assert len(np.unique(tf)) >= ncells // 50000        # minN.cluster was raised to floor(ncells / 50000) = 20
    got = ctx.smetac(labels, se1, hp)
    assert np.array_equal(got["tf"], ref["tf"]) and np.array_equal(got["finalColor"], ref["finalColor"])
    # just below the threshold the other branch applies (:103-109)
    tf_small = ctx.smetac_centroids(cen, 999_999, hp)
    ref_small = orc.smetac(labels[:-1], se1[:-1], prm)
    assert np.array_equal(tf_small, ref_small["tf"])


# ---------------------------------------------------------------------------------------------------------
# getrowColor beyond 40 clusters: the colour index wraps (quirk B1, R/getrowColor.R:59-68)
# ---------------------------------------------------------------------------------------------------------
def test_getrowcolor_colour_wrap(ctx):
    rng = np.random.default_rng(5)
    E = rng.normal(size=(260, 30))
    for ncl in (45, 80):
        got, _ = ctx.getrowcolor(E, hc_params(n_cluster=ncl))
        ref, _ = orc.getrowcolor(E, orc.hc_params(n_cluster=ncl))
        assert np.array_equal(got, ref)
        assert got.max() == 40 and len(np.unique(got)) == 40
    r = api.getrowColor(E, indN_cluster=45, ctx=ctx)
    assert set(r["rowColor"]) <= set(api.colorL) and len(set(r["rowColor"])) == 40


# ---------------------------------------------------------------------------------------------------------
# testlog (R/SHARP.R:877-924)
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind,normalize", [("tpm", False), ("umi", True)])
def test_testlog(ctx, kind, normalize):
    m, n, p = 1500, 400, 120
    x, _ = synth.make_expression(m, n, n_types=4, seed=8, kind=kind, zero_frac=0.8, sep=2.0, frac=0.4)
    e = api.Expression.wrap(x)
    if normalize:
        e = api.Expression(e.m, e.n, e.dense, normalize=True)
    for seed in (1, 2, 3):
        cells = r_sample_perm(n, seed)[:100] - 1
        flag_t, msil_t = rtrans.testlog_transcribed(np.asarray(x), p, cells, colsum=x.sum(0) if normalize else None)
        assert api.testlog(e, n, p, 100, ctx=ctx, _seed=seed) == flag_t
        # the two maxsil values behind the rule, through the same calls api.testlog makes
        rm = ctx.upload_rm([ranM2(m, p, 5)])
        for logkind, ms_t in zip((0, 2), msil_t):
            proj = ctx.rp_project(m, n, rm, dense=x, cells=cells, normalize=2 if normalize else 0, logkind=logkind)
            _, ms = ctx.getrowcolor(proj[0], hc_params("ward.D", None, 2, 40, 0.0, 2.0))
            assert ms == pytest.approx(ms_t, rel=1e-9)
        rm.close()


# ---------------------------------------------------------------------------------------------------------
# SHARP_fpart / SHARP_unlimited2 (R/SHARP_unlimited2.R): log10, round(., 1), block maxN = 40, two levels
# ---------------------------------------------------------------------------------------------------------
def test_fpart_and_unlimited2(ctx):
    m, K, seed, ng = 900, 3, 17, 500
    sizes = [2300, 1900, 2150]
    x, truth = synth.make_expression(m, sum(sizes), n_types=5, seed=12, kind="tpm", zero_frac=0.7, sep=2.5, frac=0.5)
    parts, o = [], 0
    for n in sizes:
        parts.append(np.asfortranarray(x[:, o:o + n]))
        o += n
    ncells = sum(sizes)
    p = math.ceil(math.log2(ncells) / 0.04)
    rms = [ranM2(m, p, 50 + seed + k) for k in range(1, K + 1)]
    reinds = [np.asarray(r_sample_perm(n, 50)) for n in sizes]
    # one part through SHARP_fpart: block-level ids, E1, folds
    f = api.SHARP_fpart(parts[0], K, p, ng, rN_seed=seed, ctx=ctx)
    fc_t, e1_t, folds_t = rtrans.fpart_transcribed(np.asarray(parts[0]), rms, p, K, ng, reinds[0], orc.hc_params(max_n=40))
    assert np.array_equal(first_appearance(f["fColor"]), first_appearance(np.unique(fc_t, return_inverse=True)[1]))
    assert np.array_equal(f["folds"], folds_t) and f["nmcluster"] == len(set(fc_t.tolist()))
    assert np.allclose(f["E1"], e1_t, rtol=0, atol=1e-12)    # rounded to one decimal, then averaged over K
    # the driver, dense and CSC parts
    pred_t, E1_t = rtrans.unlimited2_transcribed([np.asarray(a) for a in parts], rms, p, K, ng, reinds)
    for plist in (parts, [synth.to_csc(a) + (a.shape,) for a in parts]):
        r = api.SHARP_unlimited2(plist, ensize_K=K, partition_ncells=ng, rN_seed=seed, logflag=False, ctx=ctx)
        assert np.array_equal(r["pred_clusters"], pred_t)
        assert np.allclose(r["viE"], E1_t, rtol=0, atol=1e-12) and r["reduced.ndim"] == p
        assert r["N.pred_clusters"] == len(np.unique(pred_t))
    assert synth.ari(pred_t, truth) > 0.5


# ---------------------------------------------------------------------------------------------------------
# SHARP_unlimited3 (R/SHARP_unlimited3.R): parts read from a directory in the order of the first integer in their names
# ---------------------------------------------------------------------------------------------------------
def test_unlimited3_from_directory(ctx, tmp_path):
    m, K, seed = 900, 3, 31
    sizes = [10000, 10650, 11200]
    x, truth = synth.make_expression(m, sum(sizes), n_types=5, seed=21, kind="umi", zero_frac=0.8, sep=2.0, frac=0.4)
    parts, o = [], 0
    for n in sizes:
        parts.append(np.asfortranarray(x[:, o:o + n]))
        o += n
    for name, a in zip(["cells_2.npz", "cells_10.npz", "cells_1.npz"], [parts[1], parts[2], parts[0]]):
        cp, ri, v = synth.to_csc(a)
        np.savez(tmp_path / name, p=cp, i=ri, x=v, Dim=np.array(a.shape))
    ncells = sum(sizes)
    got = api.SHARP_unlimited3({"dir": str(tmp_path), "ncells": ncells, "ngenes": m}, viewflag=False, rN_seed=seed,
                               ensize_K=K, exp_type="UMI", ctx=ctx)
    # the oracle, driven like R/SHARP_unlimited3.R:103-183: the global sMetaC takes part 1's maxN.cluster
    p = math.ceil(math.log2(ncells) / 0.04)
    rms = [ranM2(m, p, 50 + seed + k) for k in range(1, K + 1)]
    preds, vies, part_of = [], [], []
    for i, a in enumerate(parts):
        n = a.shape[1]
        prm = orc.SharpParams(1, 1, K, p, 2000, 0, 0, 0, orc.hc_params(max_n=max(40, -(-n // 5000))), 2, -1)
        r = orc.sharp(m, n, rms, prm, csc=synth.to_csc(a), colsum=a.sum(0), reind=r_sample_perm(n, 50))
        preds.append(r["pred_clusters"])
        vies.append(r["viE"])
        part_of.append(np.full(n, i + 1))
    hc = orc.hc_params(max_n=max(40, -(-sizes[0] // 5000)))
    final, nf = orc.unlimited_combine(np.concatenate(part_of), np.concatenate(preds), np.concatenate(vies), hc)
    assert np.array_equal(got["pred_clusters"], final) and got["N.pred_clusters"] == nf
    assert synth.ari(final, truth) > 0.8


# ---------------------------------------------------------------------------------------------------------
# RPmat (R/RPmat.R:14-47) and the 50-dimensional re-projection of viE for viewflag (R/SHARP_unlimited.R:216-228)
# ---------------------------------------------------------------------------------------------------------
def test_rpmat(ctx):
    m, n, p = 800, 150, 40
    x, _ = synth.make_expression(m, n, n_types=3, seed=9)
    for data in (x, synth.to_csc(x) + ((m, n),)):
        r = api.RPmat(data, p, 77, ctx=ctx)
        R = ranM2(m, p, 77)
        assert np.array_equal(r["R"]["i"], R["i"]) and np.array_equal(r["R"]["x"], R["x"])
        ref = orc.rp_project(m, n, R, dense=x, logkind=0)            # t(projmat)
        assert r["projmat"].shape == (p, n)
        assert relerr(r["projmat"], ref.T) <= 1e-5 and relerr(r["projmat"], ref.T) <= 1e-12
    # the dense product the R line writes: 1/sqrt(p) * t(x) %*% scdata
    Rd = np.zeros((m, p))
    for j in range(p):
        Rd[R["i"][R["p"][j]:R["p"][j + 1]], j] = R["x"][R["p"][j]:R["p"][j + 1]]
    assert relerr(r["projmat"], (Rd.T / math.sqrt(p)) @ np.asarray(x)) <= 1e-12


def test_view_project(ctx):
    rng = np.random.default_rng(3)
    ncells, p = 3000, 333
    E1 = rng.normal(size=(ncells, p)) * 3.0
    z0 = ranM2(p, 50, 50 + 2103 + 5)
    got = api._view_project(ctx, E1, z0)
    Z = np.zeros((p, 50))
    for j in range(50):
        Z[z0["i"][z0["p"][j]:z0["p"][j + 1]], j] = z0["x"][z0["p"][j]:z0["p"][j + 1]]
    ref = (E1 @ Z) / math.sqrt(50)                                   # as.matrix(1/sqrt(kdim) * E1 %*% z0)
    assert got.shape == (ncells, 50) and relerr(got, ref) <= 1e-12


def test_unlimited_viewflag_above_1e5_cells_uses_the_reprojection(ctx):
    """viewflag = TRUE with more than 1e5 cells: viE is the 50-dimensional re-projection (quirk B3: k = ensize.K)"""
    m, K, seed = 300, 2, 5
    sizes = [50500, 50600]
    x, _ = synth.make_expression(m, sum(sizes), n_types=4, seed=6, kind="umi", zero_frac=0.6, sep=2.5, frac=0.5)
    parts = [synth.to_csc(np.asfortranarray(x[:, :sizes[0]])) + ((m, sizes[0]),),
             synth.to_csc(np.asfortranarray(x[:, sizes[0]:])) + ((m, sizes[1]),)]
    r = api.SHARP_unlimited(parts, viewflag=True, rN_seed=seed, ensize_K=K, exp_type="UMI", ctx=ctx)
    ncells = sum(sizes)
    assert r["viE"].shape == (ncells, 50) and r["x0"].shape == (ncells, r["N.pred_clusters"])
    r2 = api.SHARP_unlimited(parts, viewflag=False, rN_seed=seed, ensize_K=K, exp_type="UMI", ctx=ctx)
    assert np.array_equal(r["pred_clusters"], r2["pred_clusters"])


def test_unlimited3_streamed_from_shcsc_files(ctx, tmp_path):
    """SHARP_unlimited3 over a directory of SHCSC001 files: batches of parts through the fused loop over parts while the
    native reader (pread into pinned buffers) loads the next batch -- same labels as the in-memory driver with part 1's
    k-range, for several batch sizes"""
    from sharp_b200 import io as sio
    m, K, seed = 900, 3, 31
    sizes = [10000, 10650, 11200, 10100, 10300]
    x, truth = synth.make_expression(m, sum(sizes), n_types=5, seed=22, kind="umi", zero_frac=0.8, sep=2.0, frac=0.4)
    parts, o = [], 0
    for i, n in enumerate(sizes):
        a = np.asfortranarray(x[:, o:o + n])
        parts.append(synth.to_csc(a) + (a.shape,))
        sio.write_csc(tmp_path / f"part_{i + 1}.csc", m, n, *parts[-1][:3])
        o += n
    ref = api.SHARP_unlimited(parts, viewflag=False, rN_seed=seed, ensize_K=K, exp_type="UMI", ctx=ctx, _krange_from_part1=True)
    for batch in (2, 5, 1):
        got = api.SHARP_unlimited3({"dir": str(tmp_path), "ncells": sum(sizes), "ngenes": m}, viewflag=False, rN_seed=seed,
                                   ensize_K=K, exp_type="UMI", ctx=ctx, _batch=batch)
        assert np.array_equal(got["pred_clusters"], ref["pred_clusters"]), batch
        assert got["N.pred_clusters"] == ref["N.pred_clusters"]
    assert synth.ari(ref["pred_clusters"], truth) > 0.6
