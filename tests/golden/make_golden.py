"""Generates tests/golden/*.npz.

The reference is pure R and there is no R interpreter in this image (and the reference ships no golden vectors of its
own), so these fixtures are outputs of the CPU ORACLE (oracle/sharp_oracle.cpp) on small seeded inputs.
They (i) pin the oracle against regressions (tests/test_golden.py, CPU) and (ii) give the CUDA path fixed vectors to
reproduce on the GPU box, where neither /root/reference nor a rebuilt oracle is needed to read them.
If a machine with R + the SHARP package ever becomes available, the `case_*` inputs below are what to feed the real
reference (SHARP_large / get_opt_hclust / wMetaC / sMetaC with the same ranM matrices) to replace these files.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import orc  # noqa: E402
import synth  # noqa: E402
from sharp_b200.rrng import r_sample_perm, ranM2  # noqa: E402


def case_inputs(name):
    """seeded inputs, regenerated identically by the tests (numpy Generator streams are stable across versions)"""
    if name == "project":
        x, _ = synth.make_expression(600, 48, seed=21, kind="umi", zero_frac=0.8)
        return dict(x=x, rms=[ranM2(600, 19, 2154 + k) for k in range(2)])
    if name == "opt_hclust":
        rng = np.random.default_rng(22)
        cen = rng.normal(size=(3, 24)) * 4
        X = np.concatenate([cen[c] + rng.normal(size=(20, 24)) for c in range(3)])
        return dict(X=X)
    if name == "wmetac":
        rng = np.random.default_rng(23)
        truth = np.repeat(np.arange(1, 4), 40)
        cols = []
        for c in range(5):
            lab = truth.copy()
            flip = rng.random(120) < 0.1
            lab[flip] = rng.integers(1, 5, flip.sum())
            cols.append(lab)
        return dict(labels=np.stack(cols, 1).astype(np.int32))
    if name == "smetac":
        rng = np.random.default_rng(24)
        cen = rng.normal(size=(4, 30)) * 3
        truth = np.tile(np.repeat(np.arange(4), 15), 3)
        E = cen[truth] + rng.normal(size=(len(truth), 30))
        labels = (np.repeat(np.arange(3), 60) * 10 + truth).astype(np.int32)
        return dict(labels=labels, E=E)
    if name == "pipeline":
        x, truth = synth.make_expression(900, 700, n_types=4, seed=25, kind="umi", zero_frac=0.8, sep=2.5, frac=0.5)
        return dict(x=x, truth=truth, rms=[ranM2(900, 50, 50 + 2103 + k) for k in range(1, 4)], reind=r_sample_perm(700, 50))
    raise KeyError(name)


def main():
    c = case_inputs("project")
    cs = c["x"].sum(0)
    proj = np.stack([orc.rp_project(600, 48, r, csc=synth.to_csc(c["x"]), colsum=cs, logkind=2) for r in c["rms"]])
    np.savez_compressed(os.path.join(HERE, "project.npz"), proj=proj, checksum_x=np.array([c["x"].sum()]))

    c = case_inputs("opt_hclust")
    _, d = orc.zscore_corrdist(c["X"])
    ia, ib, h = orc.hclust(d, orc.WARD_D)
    r = orc.opt_hclust(c["X"], 0, orc.hc_params())
    np.savez_compressed(os.path.join(HERE, "opt_hclust.npz"), ia=ia, ib=ib, height=h, f=r["f"], msil=r["msil"],
                        chind=r["CHind"], optn=np.array([r["optN.cluster"]]), oind=np.array([r["oind"]]))

    c = case_inputs("wmetac")
    r = orc.wmetac(c["labels"], orc.hc_params())
    np.savez_compressed(os.path.join(HERE, "wmetac.npz"), finalc=r["finalC"], x0=r["x0"], w1=r["w1"])

    c = case_inputs("smetac")
    r = orc.smetac(c["labels"], c["E"], orc.hc_params())
    np.savez_compressed(os.path.join(HERE, "smetac.npz"), finalcolor=r["finalColor"], tf=r["tf"])

    c = case_inputs("pipeline")
    prm = orc.SharpParams(1, 1, 3, 50, 250, 0, 0, 0, orc.hc_params(), 2, -1)
    r = orc.sharp(900, 700, c["rms"], prm, csc=synth.to_csc(c["x"]), colsum=c["x"].sum(0), reind=c["reind"])
    np.savez_compressed(os.path.join(HERE, "pipeline.npz"), pred=r["pred_clusters"], vie=r["viE"], x0=r["x0"])
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
