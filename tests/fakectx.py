"""A stand-in for sharp_b200.Context whose compute calls are answered by the CPU ORACLE -- test infrastructure for
the `-m "not gpu"` tests of the HOST glue (argument defaults, seeds, block layout, gather / relabel / merge rules,
sharding over ranks).  It lets the Python mirror of SHARP() / SHARP_unlimited() run end to end on a machine without
a GPU; it is never importable from the product package."""
import numpy as np

import orc


def _densify(m, n, csc):
    cp, ri, v = csc
    x = np.zeros((m, n), order="F")
    for c in range(n):
        x[ri[cp[c]:cp[c + 1]], c] = v[cp[c]:cp[c + 1]]
    return x


class FakeRm:
    def __init__(self, rms):
        self.rms = list(rms)
        self.K = len(rms)
        self.m, self.p = int(rms[0]["Dim"][0]), int(rms[0]["Dim"][1])
        self._h = True

    def close(self):
        pass


def _orc_hc(prm):
    return orc.hc_params(prm.hmethod, prm.n_cluster, prm.min_n, prm.max_n, prm.sil_thre, prm.height_ntimes)


class FakeContext:
    """answers the subset of Context used by sharp_b200.api with oracle calls"""

    def __init__(self, device=0):
        self.device = device
        self._h = True
        self.calls = []
        self._last = None

    def upload_rm(self, rms):
        return FakeRm(rms)

    def rp_project(self, m, n, rm, dense=None, csc=None, cells=None, normalize=0, colsum=None, norm_mul=1e6, logkind=2,
                   round_digits=-1):
        if normalize == 2:
            colsum = dense.sum(0) if dense is not None else np.add.reduceat(csc[2], csc[0][:-1])
        kw = dict(dense=dense) if dense is not None else dict(csc=csc)
        return np.stack([orc.rp_project(m, n, r, cells=cells, colsum=colsum if normalize else None, norm_mul=norm_mul,
                                        logkind=logkind, round_digits=round_digits, **kw) for r in rm.rms])

    def getrowcolor(self, emat, prm):
        return orc.getrowcolor(emat, _orc_hc(prm))

    def opt_hclust(self, mat, symmetric, prm, exact=False, want_v=True):
        return orc.opt_hclust(mat, int(bool(symmetric)), _orc_hc(prm))

    def wmetac(self, labels, prm, want_x0=True):
        return orc.wmetac(labels, _orc_hc(prm))

    def smetac(self, labels, se1, prm):
        return orc.smetac(labels, se1, _orc_hc(prm))

    def smetac_centroids(self, cen, ncells_total, prm):
        return orc.smetac_centroids(cen, ncells_total, _orc_hc(prm))

    def run(self, rm, prm, m=None, n=None, dense=None, csc=None, expr=None, colsum=None, reind=None, want_vie=True,
            want_x0=True, max_x0_cols=None):
        self.calls.append(("run", n, prm.large, prm.logflag, prm.partition_ncells, prm.normalize, rm.K))
        if prm.normalize == 2:
            colsum = dense.sum(0) if dense is not None else np.add.reduceat(csc[2], csc[0][:-1])
        oprm = orc.SharpParams(prm.large, prm.logflag, rm.K, rm.p, prm.partition_ncells, prm.n_cluster, prm.enp_n_cluster,
                               prm.ind_n_cluster, _orc_hc(prm.hc), prm.logkind or 2, prm.round_digits)
        kw = dict(dense=dense) if dense is not None else dict(csc=csc)
        if getattr(prm, "skip_smetac", 0):  # SHARP_fpart: the block stage only (tests/rtrans.py transcribes it)
            import rtrans
            x = dense if dense is not None else _densify(m, n, csc)
            hcb = _orc_hc(prm.hc)
            hcb.max_n = prm.block_max_n or hcb.max_n
            re = np.asarray(reind) if reind is not None else np.arange(1, n + 1)
            fcol, e1, _ = rtrans.fpart_transcribed(np.asarray(x), rm.rms, rm.p, rm.K, prm.partition_ncells, re, hcb,
                                                   bool(prm.logflag), colsum if prm.normalize else None,
                                                   enp_hc=orc.hc_params(prm.hc.hmethod, prm.enp_n_cluster, prm.hc.min_n,
                                                                        prm.hc.max_n, prm.hc.sil_thre, prm.hc.height_ntimes))
            assert prm.logkind == 10 and prm.round_digits == 1
            # any injective integer code of the strings will do (the product returns block-order positions)
            codes = {c: i + 1 for i, c in enumerate(sorted(set(fcol.tolist())))}
            self._last = {"viE": e1}
            return {"labels": np.array([codes[c] for c in fcol.tolist()], dtype=np.int32), "viE": e1 if want_vie else None,
                    "x0": None, "x0_cols": len(codes)}
        r = orc.sharp(m, n, rm.rms, oprm, colsum=colsum if prm.normalize else None, reind=reind, **kw)
        # the oracle returns pred_clusters AFTER the host glue (merge + relabel); relabelling is idempotent, and the
        # merge only applies above 1e4 cells, so feeding it back as the "raw" device labels exercises the same glue
        self._last = r
        return {"labels": r["pred_clusters"].copy(), "viE": r["viE"] if want_vie else None,
                "x0": r.get("x0") if want_x0 else None, "x0_cols": 0}

    def parts_prefetch(self, m, parts, group=0, lanes=0, sharded=None):
        return None

    def run_parts(self, rm, prm, m, parts, reinds, small_thre=10, cen_cap=64, group=0, lanes=0, sharded=None):
        """sharp_run_parts: per part SHARP_large + merge of small clusters (n > 1e4) + first-appearance relabel + centroids"""
        from sharp_b200.api import _first_appearance_codes, _merge_small
        out = []
        for pt, re in zip(parts, reinds):
            n = int(pt["n"])
            r = self.run(rm, prm, m=m, n=n, dense=pt.get("dense"), csc=pt.get("csc"), reind=re, want_x0=False)
            lab = r["labels"]
            if prm.n_cluster == 0 and n > 10000:
                lab = _merge_small(lab, small_thre)
            cid, _ = _first_appearance_codes(lab)
            nc = int(cid.max())
            cen, cnt = self.centroids(cid, nc, rm.p)
            out.append({"pred_clusters": cid, "N.pred_cluster": nc, "cen": cen, "counts": cnt})
        return out

    def centroids(self, labels, nclust, p):
        vie = self._last["viE"]
        cen = np.zeros((nclust, p))
        cnt = np.zeros(nclust, dtype=np.int64)
        for c in range(1, nclust + 1):
            rows = np.flatnonzero(labels == c)
            s = np.zeros(p)
            for i in rows:  # ascending row order, like colMeans
                s = s + vie[i]
            cen[c - 1] = s / len(rows)
            cnt[c - 1] = len(rows)
        return cen, cnt

    def last_vie(self, n, p):
        return self._last["viE"]

    def last_member(self, k, n, p):
        raise NotImplementedError
