"""Golden fixtures (tests/golden/*.npz, written by tests/golden/make_golden.py): the oracle must keep reproducing
them on CPU, and the CUDA path must reproduce them on the GPU box through the C ABI."""
import os
import sys

import numpy as np
import pytest

import orc
import synth

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_golden import case_inputs  # noqa: E402


def gold(name):
    return np.load(os.path.join(HERE, "golden", name + ".npz"))


# ------------------------------------------------------------------------------------------- oracle (CPU)
def test_oracle_reproduces_golden_projection():
    c, g = case_inputs("project"), gold("project")
    assert c["x"].sum() == g["checksum_x"][0]  # the seeded input itself is stable
    cs = c["x"].sum(0)
    for k, r in enumerate(c["rms"]):
        assert np.array_equal(orc.rp_project(600, 48, r, csc=synth.to_csc(c["x"]), colsum=cs, logkind=2), g["proj"][k])


def test_oracle_reproduces_golden_clustering():
    c, g = case_inputs("opt_hclust"), gold("opt_hclust")
    _, d = orc.zscore_corrdist(c["X"])
    ia, ib, h = orc.hclust(d, orc.WARD_D)
    assert np.array_equal(ia, g["ia"]) and np.array_equal(ib, g["ib"]) and np.allclose(h, g["height"], rtol=1e-13)
    r = orc.opt_hclust(c["X"], 0, orc.hc_params())
    assert np.array_equal(r["f"], g["f"]) and r["optN.cluster"] == g["optn"][0] == 3
    assert np.allclose(r["msil"], g["msil"], atol=1e-13)


def test_oracle_reproduces_golden_meta_clustering():
    c, g = case_inputs("wmetac"), gold("wmetac")
    r = orc.wmetac(c["labels"], orc.hc_params())
    assert np.array_equal(r["finalC"], g["finalc"]) and np.array_equal(r["x0"], g["x0"]) and np.array_equal(r["w1"], g["w1"])
    c, g = case_inputs("smetac"), gold("smetac")
    r = orc.smetac(c["labels"], c["E"], orc.hc_params())
    assert np.array_equal(r["finalColor"], g["finalcolor"]) and np.array_equal(r["tf"], g["tf"])


def test_oracle_reproduces_golden_pipeline():
    c, g = case_inputs("pipeline"), gold("pipeline")
    prm = orc.SharpParams(1, 1, 3, 50, 250, 0, 0, 0, orc.hc_params(), 2, -1)
    r = orc.sharp(900, 700, c["rms"], prm, csc=synth.to_csc(c["x"]), colsum=c["x"].sum(0), reind=c["reind"])
    assert np.array_equal(r["pred_clusters"], g["pred"])
    assert synth.ari(g["pred"], c["truth"]) > 0.9


# ------------------------------------------------------------------------------------------- CUDA path (GPU box)
@pytest.fixture(scope="module")
def ctx():
    from sharp_b200 import Context
    c = Context(0)
    yield c
    c.close()


@pytest.mark.gpu
def test_gpu_reproduces_golden_projection(ctx):
    c, g = case_inputs("project"), gold("project")
    rm = ctx.upload_rm(c["rms"])
    got = ctx.rp_project(600, 48, rm, csc=synth.to_csc(c["x"]), normalize=2, logkind=2)
    err = np.max(np.abs(got - g["proj"])) / np.max(np.abs(g["proj"]))
    assert err <= 1e-5          # the contract (BASELINE.json north_star)
    assert err <= 1e-12         # what fp64 accumulation in gene order delivers


@pytest.mark.gpu
def test_gpu_reproduces_golden_clustering(ctx):
    from sharp_b200 import hc_params
    c, g = case_inputs("opt_hclust"), gold("opt_hclust")
    _, d = orc.zscore_corrdist(c["X"])
    ia, ib, h = ctx.hclust(d)
    assert np.array_equal(ia, g["ia"]) and np.array_equal(ib, g["ib"]) and np.allclose(h, g["height"], rtol=1e-12)
    for exact in (False, True):
        r = ctx.opt_hclust(c["X"], False, hc_params(), exact=exact)
        assert np.array_equal(r["f"], g["f"]) and r["optN.cluster"] == g["optn"][0] and r["oind"] == g["oind"][0]
        assert np.allclose(r["msil"], g["msil"], atol=1e-10)


@pytest.mark.gpu
def test_gpu_reproduces_golden_meta_clustering(ctx):
    from sharp_b200 import hc_params
    c, g = case_inputs("wmetac"), gold("wmetac")
    r = ctx.wmetac(c["labels"], hc_params())
    assert np.array_equal(r["finalC"], g["finalc"]) and np.allclose(r["x0"], g["x0"], atol=1e-15)
    assert np.allclose(r["w1"], g["w1"], rtol=1e-14)
    c, g = case_inputs("smetac"), gold("smetac")
    r = ctx.smetac(c["labels"], c["E"], hc_params())
    assert np.array_equal(r["finalColor"], g["finalcolor"]) and np.array_equal(r["tf"], g["tf"])


@pytest.mark.gpu
def test_gpu_reproduces_golden_pipeline_through_the_r_api_mirror(ctx):
    import sharp_b200
    c, g = case_inputs("pipeline"), gold("pipeline")
    r = sharp_b200.SHARP_large(synth.to_csc(c["x"]) + (c["x"].shape,), ensize_K=3, reduced_dim=50, partition_ncells=250,
                               rM=c["rms"], rN_seed=2103, ctx=ctx,
                               **{})  # no exp.type at this level: normalise like SHARP() does, through Expression
    # SHARP_large itself does not normalise (SHARP() does): compare with the un-normalised oracle run instead
    prm = orc.SharpParams(1, 1, 3, 50, 250, 0, 0, 0, orc.hc_params(), 2, -1)
    ref = orc.sharp(900, 700, c["rms"], prm, csc=synth.to_csc(c["x"]), reind=c["reind"])
    assert np.array_equal(r["pred_clusters"], ref["pred_clusters"])
    # and the golden (CPM-normalised) result through SHARP()
    r2 = sharp_b200.SHARP(synth.to_csc(c["x"]) + (c["x"].shape,), exp_type="UMI", ensize_K=3, reduced_ndim=50,
                          partition_ncells=250, base_ncells=100, logflag=False, prep=False, rN_seed=2103, ctx=ctx)
    assert np.array_equal(r2["pred_clusters"], g["pred"])
    assert np.allclose(r2["viE"], g["vie"], rtol=1e-9, atol=1e-11)
    assert np.allclose(r2["x0"], g["x0"], atol=1e-12)
