"""`-m "not gpu"` tests of the HOST side: the Python mirror of the reference's R closures (sharp_b200/api.py) with its
compute calls answered by the CPU oracle (tests/fakectx.py).  What is checked is the glue the reference keeps in R:
defaults, seeds, dispatch small/large, block layout, merge / relabel rules, SHARP_unlimited's combine."""
import math
import os

import numpy as np
import pytest

import orc
import synth
from fakectx import FakeContext
from sharp_b200 import api
from sharp_b200.rrng import r_sample_perm, ranM2


def test_first_appearance_merge_and_relabel_rules():
    codes, uniq = api._first_appearance_codes(np.array([7, 7, 3, 9, 3, 7]))
    assert list(codes) == [1, 1, 2, 3, 2, 1] and list(uniq) == [7, 3, 9]
    # clusters with < 10 cells are merged into the SMALLEST such id (R/SHARP.R:816-825)
    lab = np.array([1] * 30 + [2] * 3 + [5] * 20 + [4] * 9 + [6] * 10)
    out = api._merge_small(lab)
    assert set(out[30:33]) == {2} and set(out[53:62]) == {2} and set(out[62:]) == {6}
    # relabel by decreasing size; ties keep the STRING order of the ids (table() names), R/SHARP_unlimited.R:180-183
    lab = np.array([10] * 5 + [2] * 5 + [3] * 8 + [1] * 2)
    out = api._relabel_by_size(lab)
    assert list(out[:5]) == [2] * 5          # "10" < "2" as strings -> id 10 gets rank 2, id 2 rank 3
    assert list(out[5:10]) == [3] * 5 and list(out[10:18]) == [1] * 8 and list(out[18:]) == [4] * 2


def test_seed_validation_matches_the_reference():
    with pytest.raises(ValueError, match="should be a numeric"):
        api._check_seed("a")
    with pytest.raises(ValueError, match="should be an integer"):
        api._check_seed(1.5)
    assert api._check_seed(None) == 0.5 and api._check_seed(0.5) == 0.5 and api._check_seed(2103) == 2103
    with pytest.raises(ValueError):
        api._check_seed(0.5, allow_half=False)      # SHARP_unlimited does not accept the sentinel
    assert api._member_seed(2103, 1) == 2154 and api._member_seed(0.5, 3) == 0.5


def test_rm_list_uses_the_reference_seeds():
    rms = api._rm_list(400, 9, 3, 2103)
    for k, r in enumerate(rms, 1):
        ref = ranM2(400, 9, 50 + 2103 + k)
        assert np.array_equal(r["i"], ref["i"]) and np.array_equal(r["x"], ref["x"]) and np.array_equal(r["p"], ref["p"])
    assert np.array_equal(api._reind(777, 2103), r_sample_perm(777, 50))


def test_expression_wrapping_and_prep():
    x, _ = synth.make_expression(60, 25, seed=0)
    x[5, :] = 0
    x[7, 3] = -2.0
    e = api.Expression.wrap(x)
    assert (e.m, e.n) == (60, 25) and e.any_negative()
    e2 = e.clamp_negative()
    keep = e2.row_sums() != 0
    e3 = e2.keep_rows(keep)
    assert e3.m == int(keep.sum()) and not e3.any_negative()
    # the same through dgCMatrix slots
    cp, ri, v = synth.to_csc(x)
    s = api.Expression.wrap((cp, ri, v, x.shape))
    s3 = s.clamp_negative()
    s3 = s3.keep_rows(s3.row_sums() != 0)
    assert s3.m == e3.m
    dense = np.zeros((s3.m, s3.n))
    for c in range(s3.n):
        dense[s3.csc[1][s3.csc[0][c]:s3.csc[0][c + 1]], c] = s3.csc[2][s3.csc[0][c]:s3.csc[0][c + 1]]
    assert np.array_equal(dense, e3.dense)
    import scipy.sparse as sp
    w = api.Expression.wrap(sp.csr_matrix(x))
    assert np.array_equal(w.csc[0], cp) and np.array_equal(w.csc[1], ri)


def test_sharp_dispatch_defaults_large():
    x, truth = synth.make_expression(800, 5200, n_types=3, seed=2, sep=2.5, frac=0.5)
    ctx = FakeContext()
    r = api.SHARP(x, rN_seed=2103, reduced_ndim=40, partition_ncells=2600, logflag=False, ctx=ctx)
    # ncells >= base.ncells -> SHARP_large, K = 5, flag forced TRUE, no normalisation without exp.type
    assert ctx.calls == [("run", 5200, 1, 1, 2600, 0, 5)]
    assert r["N.cells"] == 5200 and r["N.genes"] == 800 and r["reduced.dim"] == 40 and r["ensize.K"] == 5
    assert r["paras"]["maxN.cluster"] == 40 and r["paras"]["base.ncells"] == 5000 and r["paras"]["logmark"] is True
    assert synth.ari(r["pred_clusters"], truth) > 0.9
    # identical to the oracle driven directly with the reference's seeds
    rms = [ranM2(800, 40, 50 + 2103 + k) for k in range(1, 6)]
    prm = orc.SharpParams(1, 1, 5, 40, 2600, 0, 0, 0, orc.hc_params(), 2, -1)
    ref = orc.sharp(800, 5200, rms, prm, dense=x, reind=r_sample_perm(5200, 50))
    assert np.array_equal(r["pred_clusters"], ref["pred_clusters"])
    assert r["N.pred_cluster"] == ref["N.pred_cluster"]
    assert list(r["unique_pred_clusters"]) == list(range(1, r["N.pred_cluster"] + 1))
    assert sum(r["distr_pred_clusters"].values()) == 5200


def test_sharp_dispatch_small_with_normalisation_and_ncluster():
    x, truth = synth.make_expression(500, 400, n_types=3, seed=3, kind="umi", sep=2.5, frac=0.5)
    ctx = FakeContext()
    r = api.SHARP(x, exp_type="UMI", rN_seed=7, logflag=False, forview=False, ctx=ctx)
    p = math.ceil(math.log2(400) / 0.04)
    assert ctx.calls == [("run", 400, 0, 1, 0, 2, 15)] and r["reduced.dim"] == p
    assert synth.ari(r["pred_clusters"], truth) > 0.9
    # N.cluster given and ncells < base.ncells: two blocks of ceil(n/2), indN.cluster = N.cluster, K = 15
    ctx2 = FakeContext()
    r2 = api.SHARP(x, exp_type="UMI", N_cluster=3, rN_seed=7, logflag=False, forview=False, ctx=ctx2)
    assert ctx2.calls == [("run", 400, 1, 1, 200, 2, 15)]
    assert r2["N.pred_cluster"] == 3
    # dotted R argument names are accepted
    ctx3 = FakeContext()
    api.SHARP(x, ctx=ctx3, forview=False, logflag=False, **{"rN.seed": 7, "ensize.K": 3, "exp.type": "CPM"})
    assert ctx3.calls == [("run", 400, 0, 1, 0, 0, 3)]


def test_sharp_argument_errors():
    with pytest.raises(ValueError, match="No expression data"):
        api.SHARP(None)
    x = np.ones((5, 8))
    with pytest.raises(ValueError, match="should be an integer"):
        api.SHARP(x, rN_seed=2.5, ctx=FakeContext())


def test_testlog_rule():
    x, _ = synth.make_expression(400, 150, n_types=3, seed=4, sep=2.0, frac=0.5)
    ctx = FakeContext()
    flag = api.testlog(x, 150, 60, ctx=ctx, _seed=1)
    # recompute with the oracle: ranM(E, p, 5), cells reind[1:100], getrowColor(., "ward.D", NULL, 2, 40, 0, 2)
    cells = r_sample_perm(150, 1)[:100] - 1
    rm = ranM2(400, 60, 5)
    ms = []
    for lk in (0, 2):
        pr = orc.rp_project(400, 150, rm, dense=x, cells=cells, logkind=lk)
        ms.append(orc.getrowcolor(pr, orc.hc_params(sil_thre=0.0))[1])
    assert flag == bool(ms[0] < 0.75 and ms[0] >= 0.95 * ms[1])


def make_parts(nparts, n_each, m=700, seed=5):
    x, truth = synth.make_expression(m, nparts * n_each, n_types=3, seed=seed, sep=2.5, frac=0.5)
    return [np.asfortranarray(x[:, i * n_each:(i + 1) * n_each]) for i in range(nparts)], truth


def unlimited_reference(parts, K, seed):
    """the oracle driven like R/SHARP_unlimited.R:97-183"""
    ncells = sum(x.shape[1] for x in parts)
    p = math.ceil(math.log2(ncells) / 0.04)
    m = parts[0].shape[0]
    rms = [ranM2(m, p, 50 + seed + k) for k in range(1, K + 1)]
    preds, vies, part_of = [], [], []
    for i, x in enumerate(parts):
        n = x.shape[1]
        large = int(n >= 5000)
        prm = orc.SharpParams(large, 1, K if large else 15, p, 2000, 0, 0, 0, orc.hc_params(), 2, -1)
        rm_i = rms if large else [ranM2(m, p, 50 + seed + k) for k in range(1, 16)]
        r = orc.sharp(m, n, rm_i, prm, dense=x, reind=r_sample_perm(n, 50) if large else None)
        preds.append(r["pred_clusters"])
        vies.append(r["viE"])
        part_of.append(np.full(n, i + 1))
    hc = orc.hc_params(max_n=max(40, -(-ncells // 5000)))
    final, nf = orc.unlimited_combine(np.concatenate(part_of), np.concatenate(preds), np.concatenate(vies), hc)
    return final, nf


def test_sharp_unlimited_matches_the_oracle_driver():
    parts, truth = make_parts(3, 300)
    ref, nf = unlimited_reference(parts, 5, 11)
    for streams in (1, 2):
        r = api.SHARP_unlimited(parts, viewflag=False, rN_seed=11, ctx=FakeContext(), n_streams=streams)
        assert np.array_equal(r["pred_clusters"], ref) and r["N.pred_clusters"] == nf
        assert r["N.cells"] == 900 and synth.ari(r["pred_clusters"], truth) > 0.9
    # viewflag: stacked viE and the 0/1 indicator x0
    r = api.SHARP_unlimited(parts, viewflag=True, rN_seed=11, ctx=FakeContext(), n_streams=1)
    assert r["viE"].shape == (900, math.ceil(math.log2(900) / 0.04))
    assert r["x0"].shape == (900, nf) and np.array_equal(r["x0"].argmax(1) + 1, r["pred_clusters"])


def test_sharp_unlimited_argument_errors():
    with pytest.raises(ValueError, match="LIST"):
        api.SHARP_unlimited("x", ctx=FakeContext())
    parts, _ = make_parts(2, 100)
    with pytest.raises(ValueError, match="integer"):
        api.SHARP_unlimited(parts, rN_seed=0.5, ctx=FakeContext())


def test_ari_five_indices():
    # hand-checkable contingency table
    r = api.ARI([1, 1, 1, 2, 2, 2], [1, 1, 2, 2, 3, 3])
    from sklearn.metrics import adjusted_rand_score, rand_score, fowlkes_mallows_score
    assert np.isclose(r["HA"], adjusted_rand_score([1, 1, 1, 2, 2, 2], [1, 1, 2, 2, 3, 3]))
    assert np.isclose(r["Rand"], rand_score([1, 1, 1, 2, 2, 2], [1, 1, 2, 2, 3, 3]))
    assert np.isclose(r["FM"], fowlkes_mallows_score([1, 1, 1, 2, 2, 2], [1, 1, 2, 2, 3, 3]))
    assert np.isclose(r["Jaccard"], 2 / (6 + 3 - 2))
    same = api.ARI([1, 2, 3, 1], [5, 6, 7, 5])
    assert same["HA"] == 1.0 and same["Rand"] == 1.0 and same["Jaccard"] == 1.0


def test_geta_and_getss():
    A = api.getA(["a", "b", "a", "c"]).toarray()
    assert np.array_equal(A, [[1, 0, 1, 0], [0, 1, 0, 0], [1, 0, 1, 0], [0, 0, 0, 1]])
    # two solutions over 4 cells, x = c(paste(col1,"_",1), paste(col2,"_",2))
    x = np.array(["a_1", "a_1", "b_1", "b_1", "u_2", "v_2", "v_2", "v_2"])
    R = ["a_1", "b_1", "u_2", "v_2"]
    w1 = np.array([0.1, 0.2, 0.3, 0.4])
    assert np.isclose(api.getss((1, 4), R, x, w1), 0.2 / (0.1 + 0.2 + 0.3 + 0.4))   # a & v = {2}; a | v = {1,2,3,4}
    assert api.getss((1, 2), R, x, w1) == 0.0                                       # disjoint


def test_plan_groups_covers_every_part_once():
    """the split of SHARP_unlimited's parts into groups (sharp_run_parts / sharp_parts_prefetch): contiguous, complete, no
    group larger than the group size, host data starts with half a group, few parts use small groups"""
    from sharp_b200 import _lib
    for nparts in list(range(1, 30)) + [64, 101]:
        for host in (False, True):
            for group, lanes in ((0, 0), (1, 1), (2, 2), (4, 2), (3, 3), (8, 2)):
                pl = _lib.plan_groups(nparts, host, group, lanes)
                gs, g = pl["gstart"], pl["group"]
                assert gs[0] == 0 and gs[-1] == nparts and all(b > a for a, b in zip(gs, gs[1:]))
                sizes = [b - a for a, b in zip(gs, gs[1:])]
                assert max(sizes) <= g and 1 <= pl["lanes"] <= len(sizes)
                if group:
                    assert g == min(group, nparts)
                else:
                    assert g == min(nparts, (1 if host else 2) if nparts <= 8 else (2 if host else 4))
                if host and g >= 2 and nparts > g:
                    assert sizes[0] == g // 2
                rest = sizes[1:] if (host and g >= 2 and nparts > g) else sizes
                assert max(rest) - min(rest) <= 1                       # evenly split
    assert _lib.plan_groups(26, False)["gstart"] == [0, 3, 7, 11, 15, 19, 23, 26]      # the benchmark, parts in HBM
    assert _lib.plan_groups(26, True)["gstart"] == [0, 1] + list(range(3, 26, 2)) + [26]   # the benchmark, host buffers: 3 lanes x 2 parts
    assert _lib.plan_groups(13, True)["gstart"] == [0, 1, 3, 5, 7, 9, 11, 13]          # one of two ranks


def test_sharp_unlimited2_is_two_level_like_the_reference():
    """R/SHARP_unlimited2.R: SHARP_fpart stops after the per-block wMetaC (log10, round(., 1), block maxN = 40) and ONE
    global sMetaC merges the block-level clusters of all parts -- against the literal transcription in tests/rtrans.py"""
    import math

    import rtrans
    from sharp_b200.rrng import r_sample_perm, ranM2
    x, truth = synth.make_expression(500, 1500, n_types=4, seed=12, kind="tpm", zero_frac=0.7, sep=2.5, frac=0.5)
    sizes = [700, 800]
    parts, o = [], 0
    for n in sizes:
        parts.append(np.asfortranarray(x[:, o:o + n]))
        o += n
    K, seed, ng = 3, 17, 300
    ctx = FakeContext()
    r = api.SHARP_unlimited2(parts, ensize_K=K, partition_ncells=ng, rN_seed=seed, logflag=False, ctx=ctx)
    p = math.ceil(math.log2(1500) / 0.04)
    rms = [ranM2(500, p, 50 + seed + k) for k in range(1, K + 1)]
    pred_t, E1_t = rtrans.unlimited2_transcribed([np.asarray(a) for a in parts], rms, p, K, ng,
                                                 [np.asarray(r_sample_perm(n, 50)) for n in sizes])
    assert np.array_equal(r["pred_clusters"], pred_t)
    assert np.array_equal(r["viE"], E1_t) and r["reduced.ndim"] == p and r["N.pred_clusters"] == len(np.unique(pred_t))
    assert r["x0"].shape == (1500, r["N.pred_clusters"]) and np.all(r["x0"].sum(1) == 1)
    assert all(c[0] == "run" and c[2] == 1 for c in ctx.calls) and len(ctx.calls) == 2
    # SHARP_fpart alone: block-level ids, un-shuffled folds
    f = api.SHARP_fpart(parts[0], K, p, ng, rN_seed=seed, ctx=FakeContext())
    fc_t, e1_t, folds_t = rtrans.fpart_transcribed(np.asarray(parts[0]), rms, p, K, ng, np.asarray(r_sample_perm(700, 50)),
                                                   orc.hc_params(max_n=40))
    assert synth.ari(f["fColor"], np.unique(fc_t, return_inverse=True)[1]) == 1.0
    assert np.array_equal(f["folds"], folds_t) and f["nmcluster"] == len(set(fc_t.tolist())) and f["ncells"] == 700
    assert synth.ari(r["pred_clusters"], truth) > 0.5   # round(., 1) of log10 values costs accuracy; parity is what is tested


def test_sharp_unlimited3_reads_parts_lazily_and_uses_part_one_k_range(tmp_path):
    """R/SHARP_unlimited3.R:59-61 (files ordered by the first integer in their names), :105/:124 (one part in memory at a
    time), :165-166 (the global sMetaC takes y[[1]]$paras$minN.cluster / maxN.cluster)"""
    x, _ = synth.make_expression(400, 3 * 260, n_types=3, seed=4, sep=2.5, frac=0.5)
    names = ["part10.npz", "part2.npz", "part1.npz"]   # numeric order: 1, 2, 10
    order = [2, 1, 0]
    plist = [np.asfortranarray(x[:, i * 260:(i + 1) * 260]) for i in range(3)]
    for nm, i in zip(names, order):
        cp, ri, v = synth.to_csc(plist[i])
        np.savez(tmp_path / nm, p=cp, i=ri, x=v, Dim=np.array(plist[i].shape))
    seen = []

    def reader(path):
        seen.append(os.path.basename(path))
        z = np.load(path)
        return {"p": z["p"], "i": z["i"], "x": z["x"], "Dim": tuple(int(q) for q in z["Dim"])}

    r = api.SHARP_unlimited3({"dir": str(tmp_path), "ncells": 780, "ngenes": 400, "ncells_each": [260, 260, 260]},
                             viewflag=False, rN_seed=3, ensize_K=2, reader=reader, ctx=FakeContext(), n_streams=1,
                             logflag=False, exp_type="UMI")
    ref = api.SHARP_unlimited(plist, viewflag=False, rN_seed=3, ensize_K=2, ctx=FakeContext(), n_streams=1, exp_type="UMI")
    assert seen == ["part1.npz", "part2.npz", "part10.npz"]
    assert np.array_equal(r["pred_clusters"], ref["pred_clusters"])


def test_shcsc_files_and_the_streamed_unlimited3(tmp_path):
    """SHCSC001 container: writer (python) <-> native reader (sharp_csc_file_*), the double-buffered batch reader, and
    SHARP_unlimited3 over a directory of .csc files (batches through the fused loop over parts) against the in-memory
    driver with part 1's k-range"""
    from sharp_b200 import io as sio
    m, sizes = 300, [10050, 10000, 10200]
    x, _ = synth.make_expression(m, sum(sizes), n_types=3, seed=14, kind="umi", zero_frac=0.6, sep=2.5, frac=0.5)
    plist, o = [], 0
    for n in sizes:
        plist.append(np.asfortranarray(x[:, o:o + n]))
        o += n
    for name, a in zip(["p3.csc", "p1.csc", "p2.csc"], [plist[2], plist[0], plist[1]]):
        cp, ri, v = synth.to_csc(a)
        sio.write_csc(tmp_path / name, m, a.shape[1], cp, ri, v)
    assert sio.file_info(tmp_path / "p1.csc")[:2] == (m, sizes[0])
    slot = sio.PartSlot(max(sizes), int((x != 0).sum()), pinned=False)
    mm, n, (cp, ri, v) = slot.load(str(tmp_path / "p2.csc"))
    cp0, ri0, v0 = synth.to_csc(plist[1])
    assert (mm, n) == (m, sizes[1]) and np.array_equal(cp, cp0) and np.array_equal(ri, ri0) and np.array_equal(v, v0)
    with pytest.raises(api.SharpError):
        (tmp_path / "bad.csc").write_bytes(b"not a matrix" * 10)
        sio.file_info(tmp_path / "bad.csc")
    os.remove(tmp_path / "bad.csc")
    paths = [str(tmp_path / f"p{i}.csc") for i in (1, 2, 3)]
    rd = sio.BatchReader(paths, [sio.file_info(p) for p in paths], batch=2, pinned=False)
    seen = [(b, [q[1] for q in loaded]) for b, loaded in rd]
    rd.close()
    assert seen == [(0, sizes[:2]), (1, sizes[2:])]
    ctx = FakeContext()
    r = api.SHARP_unlimited3({"dir": str(tmp_path), "ncells": sum(sizes), "ngenes": m}, viewflag=False, rN_seed=3, ensize_K=2,
                             exp_type="UMI", ctx=ctx, _batch=2)
    ref = api.SHARP_unlimited([synth.to_csc(a) + (a.shape,) for a in plist], viewflag=False, rN_seed=3, ensize_K=2, exp_type="UMI",
                              ctx=FakeContext(), n_streams=1, _krange_from_part1=True)
    assert np.array_equal(r["pred_clusters"], ref["pred_clusters"]) and r["N.pred_clusters"] == ref["N.pred_clusters"]


def test_labels_combine_native_equals_numpy_tail():
    """sharp_labels_combine (host C++) == tf[fColor] + _merge_small + _relabel_by_size, including the string order of
    equal-sized ids (R/SHARP_unlimited.R:166-183)"""
    from sharp_b200 import _lib
    rng = np.random.default_rng(5)
    for trial in range(6):
        nparts = int(rng.integers(1, 6))
        counts = [int(rng.integers(2, 9)) for _ in range(nparts)]
        ntf = sum(counts)
        tf = rng.integers(1, 14 if trial % 2 else 120, size=ntf).astype(np.int32)
        preds = []
        for t in range(nparts):
            n = int(rng.integers(40, 3000))
            pr = rng.integers(1, counts[t] + 1, size=n).astype(np.int32)
            if trial >= 3:                       # a few tiny clusters, and equal sizes (ties resolved in string order)
                pr[:n // 2] = 1
                pr[n // 2:n // 2 + 3] = min(2, counts[t])
            preds.append(pr)
        ncells = sum(len(p) for p in preds)
        for merge in (0, 10):
            want = api._combine_labels_py(tf, counts, preds, ncells, bool(merge))
            got, sizes = _lib.labels_combine(preds, counts, tf, merge)
            assert np.array_equal(got, want)
            assert np.array_equal(sizes, np.bincount(want)[1:])
    # ties between ids 2, 10, 3: string order "10" < "2" < "3"
    tf = np.array([2, 10, 3], dtype=np.int32)
    preds = [np.array([1, 2, 3, 1, 2, 3], dtype=np.int32)]
    got, _ = _lib.labels_combine(preds, [3], tf, 0)
    assert np.array_equal(got, api._combine_labels_py(tf, [3], preds, 6, False))
    assert got.tolist() == [2, 1, 3, 2, 1, 3]
    with pytest.raises(Exception):
        _lib.labels_combine([np.array([4], dtype=np.int32)], [3], tf, 0)


def test_first_appearance_codes_native():
    """match(y, unique(y)) (R/SHARP.R:429-432): the native pass against a direct transcription"""
    from sharp_b200 import _lib
    rng = np.random.default_rng(11)
    for n, hi in ((1, 1), (17, 3), (5000, 40), (30000, 700)):
        y = rng.integers(0, hi + 1, size=n).astype(np.int32)
        codes, uniq = _lib.first_appearance_codes(y, hi + 1)
        seen = list(dict.fromkeys(y.tolist()))
        assert uniq.tolist() == seen
        assert codes.tolist() == [seen.index(v) + 1 for v in y.tolist()]
        c2, u2 = api._first_appearance_codes(y)                      # the API helper routes small ids through it
        assert np.array_equal(c2, codes) and np.array_equal(u2, uniq)
    with pytest.raises(Exception):
        _lib.first_appearance_codes(np.array([0, 5], dtype=np.int32), 5)
    s = np.array(["b", "a", "b", "c"])                                # non-integer ids keep the sort-based path
    c3, u3 = api._first_appearance_codes(s)
    assert c3.tolist() == [1, 2, 1, 3] and u3.tolist() == ["b", "a", "c"]


def test_read_slots_are_pooled_between_runs(tmp_path):
    """page-locking memory costs more than reading a part: a finished BatchReader hands its slots to the next one"""
    from sharp_b200 import io as sio
    rng = np.random.default_rng(3)
    paths = []
    for i in range(3):
        x = (rng.random((30, 20 + i)) < 0.3) * rng.integers(1, 9, size=(30, 20 + i))
        cp, ri, v = synth.to_csc(x.astype(np.float64))
        p = tmp_path / f"part{i + 1}.csc"
        sio.write_csc(p, 30, x.shape[1], cp, ri, v)
        paths.append(str(p))
    infos = [sio.file_info(p) for p in paths]
    sio.release_pool()
    rd = sio.BatchReader(paths, infos, batch=2, pinned=False)
    first = {id(s) for st in rd.sets for s in st}
    got = [(m, n, int(cp[-1])) for _, batch in rd for (m, n, (cp, ri, v)) in batch]
    rd.close()
    assert got == [(30, inf[1], inf[2]) for inf in infos]
    rd2 = sio.BatchReader(paths, infos, batch=2, pinned=False)
    assert {id(s) for st in rd2.sets for s in st} == first           # the same slot objects came back from the pool
    rd2.close()
    sio.release_pool()
