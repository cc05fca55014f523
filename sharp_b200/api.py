"""Host-side mirror of the reference's R operator API for the ensemble random-projection clustering path.

R is not available in this image, so the host glue that stays in R where R exists (argument defaults, prep,
seeds / ranM / sample(), the small-cluster merge, the final relabel, the result list) is written here in
Python with the same names, argument meaning and error behaviour as the R closures it mirrors
(``NAMESPACE:3-26``); dots in R argument names become underscores (``rN.seed`` -> ``rN_seed``; the dotted
spelling is accepted through ``**kwargs`` too).  Every compute step goes through the C ABI of
``libsharpb200.so`` (include/sharp_b200.h) -- there is no CPU fallback: without the library or without a
CUDA device the calls raise.

Mirrored closures (reference file:line):
  SHARP                R/SHARP.R:44-318          SHARP_small        R/SHARP.R:339-454
  SHARP_large          R/SHARP.R:478-851         testlog            R/SHARP.R:877-924
  SHARP_unlimited      R/SHARP_unlimited.R:29-242
  SHARP_unlimited2     R/SHARP_unlimited2.R:29-267     SHARP_fpart   R/SHARP_unlimited2.R:297-544
  SHARP_unlimited3     R/SHARP_unlimited3.R:29-235
  RPmat / ranM / ranM2 R/RPmat.R:14-47, R/ranM.R:11-33, R/ranM2.R:11-35
  get_opt_hclust       R/get_opt_hclust.R:33-244   getrowColor      R/getrowColor.R:17-121
  wMetaC / getA / getss  R/wMetaC.R:15-320         sMetaC           R/sMetaC.R:17-210
  ARI                  R/ARI.R:20-42               run_Mtimes_SHARP R/run_Mtimes_SHARP.R:22-60

Result lists are dicts keyed by the R element names (``"pred_clusters"``, ``"N.pred_cluster"``, ...).
Matrices: expression is genes x cells (numpy 2-D, scipy.sparse, or dgCMatrix slots ``(p, i, x, (m, n))``);
cell-by-feature results (``viE``, ``x0``, projections) are numpy ``ncells x ncol``.
"""
from __future__ import annotations

import math
import os
import re
import sys
import time

import numpy as np

from . import _lib
from ._lib import Context, ExprDev, RmDev, RStop, RunParams, SharpError, hc_params
from .rrng import RRandom, r_sample_perm, ranM, ranM2  # noqa: F401  (ranM / ranM2 are part of the API)

__all__ = ["SHARP", "SHARP_small", "SHARP_large", "SHARP_unlimited", "SHARP_unlimited2", "SHARP_fpart",
           "SHARP_unlimited3", "RPmat", "ranM", "ranM2", "get_opt_hclust", "getrowColor", "wMetaC", "sMetaC",
           "getA", "getss", "testlog", "ARI", "run_Mtimes_SHARP", "Expression", "get_context", "set_devices",
           "set_verbose", "colorL", "stream_contexts"]

# R/getrowColor.R:52-58
colorL = ["red", "purple", "blue", "yellow", "green", "orange", "brown", "gray", "black", "coral", "beige", "cyan",
          "turquoise", "pink", "khaki", "magenta", "violet", "salmon", "goldenrod", "orchid", "seagreen", "slategray",
          "darkred", "darkblue", "darkcyan", "darkgreen", "darkgray", "darkkhaki", "darkorange", "darkmagenta",
          "darkviolet", "darkturquoise", "darksalmon", "darkgoldenrod", "darkorchid", "darkseagreen", "darkslategray",
          "deeppink", "lightcoral", "lightcyan"]

_verbose = False
_TRACE = bool(os.environ.get("SHARP_B200_TRACE"))
_contexts: dict[int, Context] = {}
_current_device = 0


def set_verbose(flag: bool) -> None:
    """The reference reports progress with cat(); the mirror is silent unless this is switched on."""
    global _verbose
    _verbose = bool(flag)


def _cat(*a):
    if _verbose:
        print(*a)


def set_devices(device: int) -> None:
    """Device used by the module-level context (the analogue of ``n.cores``: the unit of parallelism is the GPU)."""
    global _current_device
    _current_device = int(device)


def get_context(device: int | None = None) -> Context:
    d = _current_device if device is None else int(device)
    c = _contexts.get(d)
    if c is None or not c._h:
        c = Context(d)
        _contexts[d] = c
    return c


_stream_ctxs: dict[int, list] = {}
_default_streams = 8


def stream_contexts(n: int, device: int | None = None) -> list:
    """``n`` contexts (CUDA stream + workspace each) on one device; the first is the module-level context.
    Parts of a SHARP_unlimited call are clustered concurrently on them: the Ward agglomeration is latency-bound
    (one CTA per 2000-cell problem, 125 problems per 50 000-cell part), so several parts in flight are what fills
    the 148 SMs, and one part's H2D copy overlaps the others' kernels."""
    d = _current_device if device is None else int(device)
    lst = _stream_ctxs.setdefault(d, [])
    if not lst or not lst[0]._h:
        lst[:] = [get_context(d)]
    while len(lst) < n:
        lst.append(Context(d))
    return lst[:n]


def _kw(kwargs: dict) -> dict:
    """accept R's dotted argument names (``**{"rN.seed": 2103}``)"""
    return {k.replace(".", "_"): v for k, v in kwargs.items()}


# =====================================================================================================
# expression matrices
# =====================================================================================================
class Expression:
    """A genes x cells matrix the way R holds it: dense column-major, or dgCMatrix slots (``p``, ``i``, ``x``).

    ``normalize``: the CPM scaling ``t(t(x)/colSums(x))*1e6`` (R/SHARP.R:113) still has to be applied -- it is
    fused into the projection kernel's load instead of rewriting the matrix.  ``dev``: the matrix already lives
    on the device (sharp_expr_upload); host slots may then be absent."""

    def __init__(self, m, n, dense=None, csc=None, dev: ExprDev | None = None, normalize=False):
        self.m, self.n = int(m), int(n)
        self.dense, self.csc, self.dev = dense, csc, dev
        self.normalize = bool(normalize)

    @staticmethod
    def wrap(x) -> "Expression":
        if isinstance(x, Expression):
            return x
        if isinstance(x, ExprDev):
            return Expression(x.m, x.n, dev=x)
        if isinstance(x, dict) and "i" in x and "p" in x:
            m, n = x["Dim"]
            return Expression(m, n, csc=(np.asarray(x["p"], dtype=np.int64), np.asarray(x["i"], dtype=np.int32),
                                         np.asarray(x["x"], dtype=np.float64)))
        if isinstance(x, tuple) and len(x) == 4:
            cp, ri, v, (m, n) = x
            return Expression(m, n, csc=(np.asarray(cp, dtype=np.int64), np.asarray(ri, dtype=np.int32),
                                         np.asarray(v, dtype=np.float64)))
        if hasattr(x, "tocsc") and hasattr(x, "indptr") or hasattr(x, "tocsc"):
            s = x.tocsc()
            s.sort_indices()
            m, n = s.shape
            return Expression(m, n, csc=(s.indptr.astype(np.int64), s.indices.astype(np.int32),
                                         s.data.astype(np.float64)))
        a = np.asarray(x)
        if a.ndim != 2:
            raise ValueError("The expression matrix must be 2-dimensional (genes x cells)")
        return Expression(a.shape[0], a.shape[1], dense=np.asfortranarray(a, dtype=np.float64))

    # -- the R-side preprocessing of SHARP() (host, only used for ncells < 1e4 by default) ---------------
    def any_negative(self) -> bool:
        v = self.dense if self.dense is not None else self.csc[2]
        return bool((v < 0).any())

    def clamp_negative(self) -> "Expression":
        if self.dense is not None:
            return Expression(self.m, self.n, dense=np.asfortranarray(np.where(self.dense < 0, 0.0, self.dense)),
                              normalize=self.normalize)
        cp, ri, v = self.csc
        return Expression(self.m, self.n, csc=(cp, ri, np.where(v < 0, 0.0, v)), normalize=self.normalize)

    def row_sums(self) -> np.ndarray:
        if self.dense is not None:
            return self.dense.sum(axis=1)
        cp, ri, v = self.csc
        return np.bincount(ri, weights=v, minlength=self.m)

    def keep_rows(self, keep: np.ndarray) -> "Expression":
        keep = np.asarray(keep, dtype=bool)
        if keep.all():
            return self
        if self.dense is not None:
            d = np.asfortranarray(self.dense[keep, :])
            return Expression(d.shape[0], self.n, dense=d, normalize=self.normalize)
        cp, ri, v = self.csc
        newidx = np.cumsum(keep) - 1
        sel = keep[ri]
        col = np.repeat(np.arange(self.n), np.diff(cp))[sel]
        ncp = np.zeros(self.n + 1, dtype=np.int64)
        np.add.at(ncp, col + 1, 1)
        return Expression(int(keep.sum()), self.n, csc=(np.cumsum(ncp), newidx[ri[sel]].astype(np.int32), v[sel]),
                          normalize=self.normalize)

    def with_normalize(self) -> "Expression":
        """the same matrix with the CPM scaling pending (a shallow copy: subclasses that load lazily stay lazy)"""
        import copy
        c = copy.copy(self)
        c.normalize = True
        return c

    def run_kwargs(self) -> dict:
        if self.dev is not None:
            return {"expr": self.dev}
        if self.dense is not None:
            return {"m": self.m, "n": self.n, "dense": self.dense}
        return {"m": self.m, "n": self.n, "csc": self.csc}

    def project_kwargs(self) -> dict:
        if self.dense is not None:
            return {"dense": self.dense}
        if self.csc is not None:
            return {"csc": self.csc}
        raise ValueError("this operation needs the host copy of the expression matrix")


# =====================================================================================================
# small host helpers (R semantics)
# =====================================================================================================
def _first_appearance_codes(y) -> tuple[np.ndarray, np.ndarray]:
    """match(y, unique(y)) -> (codes 1.., unique values in first-appearance order)"""
    y = np.asarray(y)
    n = len(y)
    if n and y.dtype.kind in "iu" and 0 <= int(y.min()) and int(y.max()) <= min(4 * n + 1024, 2 ** 31 - 2):
        # small non-negative integers (cluster ids): one native pass with a lookup table instead of a sort
        codes, uniq = _lib.first_appearance_codes(y, int(y.max()) + 1)
        return codes, uniq.astype(y.dtype)
    vals, first, inv = np.unique(y, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")
    rank = np.empty(len(vals), dtype=np.int64)
    rank[order] = np.arange(1, len(vals) + 1)
    return rank[inv].astype(np.int32), vals[order]


def _counts(ids: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """(sorted distinct values, counts) -- table(); bincount for small non-negative integers"""
    ids = np.asarray(ids)
    if len(ids) and ids.dtype.kind in "iu" and 0 <= int(ids.min()) and int(ids.max()) <= 4 * len(ids) + 1024:
        c = np.bincount(ids)
        v = np.flatnonzero(c)
        return v.astype(ids.dtype), c[v]
    return np.unique(ids, return_counts=True)


def _merge_small(labels: np.ndarray, thre: int = 10) -> np.ndarray:
    """xt = table(x); s = names(which(xt < 10)); x[x in s] = min(as.numeric(s))   (R/SHARP.R:418-427, 816-825)"""
    labels = np.asarray(labels).copy()
    vals, cnt = _counts(labels)
    small = vals[cnt < thre]
    if len(small):
        labels[np.isin(labels, small)] = small.min()
    return labels


def _table_sorted(ids: np.ndarray) -> dict:
    vals, cnt = _counts(ids)
    return {int(v): int(c) for v, c in zip(vals, cnt)}


def _finish(labels: np.ndarray) -> dict:
    """R/SHARP.R:429-443 / 828-843: clusterID = match(y, unique(y)) and the summary fields"""
    cid, _ = _first_appearance_codes(labels)
    vals, cnt = _counts(cid)
    return {"pred_clusters": cid, "unique_pred_clusters": vals.astype(np.int64),
            "distr_pred_clusters": {int(v): int(c) for v, c in zip(vals, cnt)}, "N.pred_cluster": int(len(vals))}


def _check_seed(rN_seed, allow_half=True):
    """R/SHARP.R:169-179 (SHARP_unlimited: R/SHARP_unlimited.R:81-91 does not accept the 0.5 sentinel)"""
    if rN_seed is None:
        return 0.5
    if isinstance(rN_seed, bool) or not isinstance(rN_seed, (int, float, np.integer, np.floating)):
        raise ValueError("The rN.seed should be a numeric!")
    if float(rN_seed) % 1 != 0 and not (allow_half and float(rN_seed) == 0.5):
        raise ValueError("The rN.seed should be an integer!")
    return rN_seed


def _member_seed(rN_seed, k):
    return 0.5 if rN_seed == 0.5 else 50 + int(rN_seed) + k


def _hc(hmethod, N_cluster, minN, maxN, sil_thre, height_Ntimes, flashmark=False):
    hmethod = "ward.D" if hmethod is None else hmethod
    if flashmark and hmethod not in ("ward.D", "ward"):
        # R/get_opt_hclust.R:80: `hmethod == "ward.D" || "ward.D2"` is an error for any other method (quirk B8)
        raise RStop(_lib.E_RSTOP, "invalid 'y' type in 'x || y'")
    if hmethod not in _lib.HMETHODS:
        raise RStop(_lib.E_RSTOP, "invalid clustering method " + repr(hmethod))
    if N_cluster is not None:
        if isinstance(N_cluster, bool) or not isinstance(N_cluster, (int, float, np.integer, np.floating)):
            N_cluster = None  # is.numeric() FALSE -> automatic
        elif float(N_cluster) % 1 != 0:
            raise RStop(_lib.E_RSTOP, "The given N.cluster is not an integer!")
        elif N_cluster < 2:
            raise RStop(_lib.E_RSTOP, "The given N.cluster is less than 2, which is not suitable for clustering!")
    return hc_params(hmethod, int(N_cluster) if N_cluster is not None else None, 2 if minN is None else int(minN),
                     40 if maxN is None else int(maxN), 0.35 if sil_thre is None else float(sil_thre),
                     2.0 if height_Ntimes is None else float(height_Ntimes))


def _ncl(x) -> int:
    """N.cluster-like argument -> C ABI code (0 = NULL)"""
    if x is None or isinstance(x, bool) or not isinstance(x, (int, float, np.integer, np.floating)):
        return 0
    if float(x) % 1 != 0:
        raise RStop(_lib.E_RSTOP, "The given N.cluster is not an integer!")
    if x < 2:
        raise RStop(_lib.E_RSTOP, "The given N.cluster is less than 2, which is not suitable for clustering!")
    return int(x)


def _entropy_seed() -> int:
    return int(np.random.SeedSequence().entropy % (2 ** 31))


def _reind(n: int, rN_seed) -> np.ndarray:
    """R/SHARP.R:493-498: unseeded sample(ncells), or set.seed(50); sample(ncells)"""
    if rN_seed == 0.5:
        return _lib.r_sample_perm_native(n, _entropy_seed())
    return _lib.r_sample_perm_native(n, 50)


def _reinds_for(sizes: list, rN_seed) -> list:
    """the shuffles of several parts of ONE call (None for parts of 1e5 cells or more, R/SHARP.R:493-498).  A seeded run
    draws `set.seed(50); sample(n)` for every part: a pure function of n, so parts of equal size share one draw (the
    result is read-only); distinct sizes are drawn on separate host threads (the native generator releases the GIL)"""
    need = [n for n in sizes if n < 1e5]
    if not need:
        return [None] * len(sizes)
    if rN_seed == 0.5:       # unseeded: every part has its own stream
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(8, len(need))) as ex:
            drawn = iter(list(ex.map(lambda n: _reind(n, rN_seed), need)))
        return [next(drawn) if n < 1e5 else None for n in sizes]
    distinct = sorted(set(need))
    if len(distinct) == 1:
        table = {distinct[0]: _reind(distinct[0], rN_seed)}
    else:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(8, len(distinct))) as ex:
            table = dict(zip(distinct, ex.map(lambda n: _reind(n, rN_seed), distinct)))
    return [table[n] if n < 1e5 else None for n in sizes]


def _member_seeds(K, rN_seed, comm=None):
    """the integer seeds of the K ranM draws; an unseeded run (rN.seed = 0.5) takes them from the OS entropy -- ONCE:
    with several ranks, rank 0 draws and everybody uses its seeds (the reference builds rM once for all partitions,
    "consistent random matrices across different partitions", R/SHARP_unlimited.R:96-104)"""
    seeds = [_entropy_seed() if rN_seed == 0.5 else _member_seed(rN_seed, k) for k in range(1, K + 1)]
    if comm is not None and rN_seed == 0.5:
        seeds = comm.bcast_obj(seeds, 0)
    return seeds


def _rm_list(m, p, K, rN_seed, comm=None):
    """K x ranM(E, p, 50 + rN.seed + k) (R/SHARP.R:539-549) with R's RNG stream, one host thread per member (the
    native generator releases the GIL).  With several ranks on one host the members are dealt over the ranks and the
    slots exchanged (a draw is one sequential MT19937 stream of m * p numbers: K threads per rank on every rank would
    only fight for the cores -- 20 ms instead of 8 at 8 ranks)."""
    from concurrent.futures import ThreadPoolExecutor
    seeds = _member_seeds(K, rN_seed, comm)
    world = comm.world if comm is not None else 1
    mine = [k for k in range(K) if k % world == (comm.rank if comm is not None else 0)]

    def draw(k):
        return _lib.r_ranm(m, p, seeds[k])

    if len(mine) <= 1:
        got = {k: draw(k) for k in mine}
    else:
        with ThreadPoolExecutor(max_workers=min(len(mine), 16)) as ex:
            got = dict(zip(mine, ex.map(draw, mine)))
    if world == 1:
        return [got[k] for k in range(K)]
    # the slots travel as (p, i, sign bytes, magnitude): a ranM matrix is ternary, x = +-sqrt(sqrt(m)) (R/ranM.R:20-29)
    payload = {}
    for k, r in got.items():
        x = np.asarray(r["x"], dtype=np.float64)
        mag = float(np.abs(x[0])) if len(x) else 0.0
        if len(x) and not np.array_equal(np.abs(x), np.full(len(x), mag)):
            raise ValueError("ranM produced a matrix that is not ternary")
        payload[4 * k] = np.ascontiguousarray(r["p"], dtype=np.int32)
        payload[4 * k + 1] = np.ascontiguousarray(r["i"], dtype=np.int32)
        payload[4 * k + 2] = (x < 0).astype(np.uint8)
        payload[4 * k + 3] = np.array([mag], dtype=np.float64)
    allv = comm.allgather_parts(payload, 4 * K)
    out = []
    for k in range(K):
        mag = float(allv[4 * k + 3][0])
        x = np.where(allv[4 * k + 2] != 0, -mag, mag)
        out.append({"Dim": (m, p), "p": allv[4 * k], "i": allv[4 * k + 1], "x": x})
    return out


def _as_rmdev(ctx: Context, rM, m, p, K, rN_seed) -> tuple[RmDev, bool]:
    """rM: an RmDev (already on the device), a list of dgCMatrix dicts, or anything else (= TRUE: generate)."""
    if isinstance(rM, RmDev):
        return rM, False
    if isinstance(rM, (list, tuple)) and len(rM) > 0 and isinstance(rM[0], dict):
        return ctx.upload_rm(list(rM)), True
    return ctx.upload_rm(_rm_list(m, p, K, rN_seed)), True


# =====================================================================================================
# stage-level operators
# =====================================================================================================
def RPmat(scdata, p, seedn, ctx: Context | None = None) -> dict:
    """R/RPmat.R:14-47 -> ``{"R": dgCMatrix slots (m x p), "projmat": p x n}``"""
    ctx = ctx or get_context()
    e = Expression.wrap(scdata)
    R = ranM2(e.m, int(p), seedn)
    rm = ctx.upload_rm([R])
    try:
        proj = ctx.rp_project(e.m, e.n, rm, logkind=0, **e.project_kwargs())
    finally:
        rm.close()
    return {"R": R, "projmat": np.ascontiguousarray(proj[0].T)}


def get_opt_hclust(mat, hmethod=None, N_cluster=None, minN_cluster=None, maxN_cluster=None, sil_thre=None,
                   height_Ntimes=None, flashmark=None, ctx: Context | None = None, exact=True, **kwargs) -> dict:
    """R/get_opt_hclust.R:33-244 -> ``{"f", "v", "maxsil", "msil", "CHind", "height", "optN.cluster"}``.

    ``isSymmetric(mat)`` (R/get_opt_hclust.R:66) is evaluated here: square and equal to its transpose within
    100 eps (all.equal's tolerance)."""
    k = _kw(kwargs)
    N_cluster = k.get("N_cluster", N_cluster)
    ctx = ctx or get_context()
    mat = np.ascontiguousarray(mat, dtype=np.float64)
    prm = _hc(hmethod, N_cluster, k.get("minN_cluster", minN_cluster), k.get("maxN_cluster", maxN_cluster),
              k.get("sil_thre", sil_thre), k.get("height_Ntimes", height_Ntimes), bool(flashmark))
    sym = mat.shape[0] == mat.shape[1] and _is_symmetric(mat)
    r = ctx.opt_hclust(mat, sym, prm, exact=exact)
    return {"f": r["f"], "v": r["v"], "maxsil": r["maxsil"], "msil": r["msil"], "CHind": r["CHind"],
            "height": r["height"], "optN.cluster": r["optN.cluster"]}


def _is_symmetric(a: np.ndarray) -> bool:
    """isSymmetric.matrix: all.equal(a, t(a), tolerance = 100 * .Machine$double.eps) (mean relative difference)"""
    t = a.T
    if np.array_equal(a, t):
        return True
    if not (np.isfinite(a).all()):
        return False
    diff = np.abs(a - t).sum()
    scale = np.abs(a).sum()
    xy = diff / scale if (np.isfinite(scale) and scale > 100 * np.finfo(float).eps) else diff
    return bool(xy < 100 * np.finfo(float).eps)


def getrowColor(Emat, hmethod=None, indN_cluster=None, minN_cluster=None, maxN_cluster=None, sil_thre=None,
                height_Ntimes=None, flashmark=False, ctx: Context | None = None, **kwargs) -> dict:
    """R/getrowColor.R:17-121 -> ``{"rowColor": colour names, "maxsil", "mat"}``"""
    k = _kw(kwargs)
    ctx = ctx or get_context()
    Emat = np.ascontiguousarray(Emat, dtype=np.float64)
    prm = _hc(hmethod, k.get("indN_cluster", indN_cluster), k.get("minN_cluster", minN_cluster),
              k.get("maxN_cluster", maxN_cluster), k.get("sil_thre", sil_thre), k.get("height_Ntimes", height_Ntimes),
              bool(flashmark))
    color, maxsil = ctx.getrowcolor(Emat, prm)
    return {"rowColor": np.array(colorL, dtype=object)[color - 1], "maxsil": maxsil, "mat": Emat}


def wMetaC(nC, hmethod=None, enN_cluster=None, minN_cluster=None, maxN_cluster=None, sil_thre=None,
           height_Ntimes=None, ctx: Context | None = None, **kwargs) -> dict:
    """R/wMetaC.R:15-226.  ``nC``: N x C matrix of cluster labels (any hashable values, e.g. colour names).
    -> ``{"finalC": N strings (meta-cluster ids, like R), "x0": N x N.cluster}``"""
    k = _kw(kwargs)
    ctx = ctx or get_context()
    nC = np.asarray(nC)
    if nC.ndim != 2:
        raise ValueError("nC must be an N x C matrix")
    codes = np.empty(nC.shape, dtype=np.int32, order="F")
    for c in range(nC.shape[1]):
        codes[:, c], _ = _first_appearance_codes(nC[:, c])
    prm = _hc(hmethod, k.get("enN_cluster", enN_cluster), k.get("minN_cluster", minN_cluster),
              k.get("maxN_cluster", maxN_cluster), k.get("sil_thre", sil_thre), k.get("height_Ntimes", height_Ntimes))
    r = ctx.wmetac(codes, prm)
    return {"finalC": r["finalC"].astype(str), "x0": r["x0"]}


def sMetaC(rerowColor, sE1, folds=None, hmethod=None, finalN_cluster=None, minN_cluster=None, maxN_cluster=None,
           sil_thre=None, height_Ntimes=None, ctx: Context | None = None, **kwargs) -> dict:
    """R/sMetaC.R:17-210 (``folds`` is accepted and unused, like in the reference).
    -> ``{"finalColor": ncells ints, "tf": nC ints}``"""
    k = _kw(kwargs)
    ctx = ctx or get_context()
    codes, _ = _first_appearance_codes(np.asarray(rerowColor))
    prm = _hc(hmethod, k.get("finalN_cluster", finalN_cluster), k.get("minN_cluster", minN_cluster),
              k.get("maxN_cluster", maxN_cluster), k.get("sil_thre", sil_thre), k.get("height_Ntimes", height_Ntimes))
    r = ctx.smetac(codes, np.ascontiguousarray(sE1, dtype=np.float64), prm)
    return {"finalColor": r["finalColor"], "tf": r["tf"]}


def getA(rowColor):
    """R/wMetaC.R:242-283: the N x N 0/1 co-membership matrix of one clustering, as a scipy CSC matrix.
    (The GPU path never materialises it -- kept for API completeness; needs scipy.)"""
    import scipy.sparse as sp
    codes, _ = _first_appearance_codes(np.asarray(rowColor))
    N = len(codes)
    H = sp.csr_matrix((np.ones(N), (np.arange(N), codes - 1)))
    return (H @ H.T).tocsc()


def getss(pind, R, x, w1) -> float:
    """R/wMetaC.R:299-320: weighted Jaccard similarity of clusters ``R[pind[0]]`` and ``R[pind[1]]`` where ``x`` is
    the flattened (column-major) N x C matrix of "label_member" strings.  1-based ``pind`` like R."""
    x = np.asarray(x)
    w1 = np.asarray(w1, dtype=np.float64)
    N = len(w1)
    a = np.flatnonzero(x == R[pind[0] - 1]) % N
    b = np.flatnonzero(x == R[pind[1] - 1]) % N
    inter = np.intersect1d(a, b)
    if len(inter) == 0:
        return 0.0
    union = np.concatenate([a, b[~np.isin(b, a)]])
    si = 0.0
    for i in inter:
        si += w1[i]
    su = 0.0
    for i in union:
        su += w1[i]
    return si / su


def ARI(gt, pred) -> dict:
    """R/ARI.R:20-42 (clues::adjustedRand): Rand, HA, MA, FM, Jaccard"""
    a = np.unique(np.asarray(gt), return_inverse=True)[1]
    b = np.unique(np.asarray(pred), return_inverse=True)[1]
    n = len(a)
    ct = np.zeros((a.max() + 1, b.max() + 1), dtype=np.float64)
    np.add.at(ct, (a, b), 1)
    c2 = lambda v: v * (v - 1) / 2.0
    nij2, ni2, nj2, tot = c2(ct).sum(), c2(ct.sum(1)).sum(), c2(ct.sum(0)).sum(), c2(n)
    rand = 1.0 + (2 * nij2 - ni2 - nj2) / tot
    exp = ni2 * nj2 / tot
    ha = (nij2 - exp) / (0.5 * (ni2 + nj2) - exp) if 0.5 * (ni2 + nj2) != exp else 1.0
    # Morey & Agresti: expectation under the multinomial model
    sq_i, sq_j = (ct.sum(1) ** 2).sum(), (ct.sum(0) ** 2).sum()
    nc = (n * (n * n + 1) - (n + 1) * sq_i - (n + 1) * sq_j + 2.0 * sq_i * sq_j / n) / (2.0 * (n - 1))
    a_ = tot + (ct ** 2).sum() - 0.5 * (sq_i + sq_j)
    ma = (a_ - nc) / (tot - nc) if tot != nc else 1.0
    fm = nij2 / math.sqrt(ni2 * nj2) if ni2 > 0 and nj2 > 0 else 0.0
    jac = nij2 / (ni2 + nj2 - nij2) if (ni2 + nj2 - nij2) > 0 else 0.0
    return {"Rand": rand, "HA": ha, "MA": ma, "FM": fm, "Jaccard": jac}


# =====================================================================================================
# testlog
# =====================================================================================================
def testlog(scExp, ncells=None, p=None, sncells=None, n_cores=None, ctx: Context | None = None, _seed=None) -> bool:
    """R/SHARP.R:877-924: on <= ``sncells`` randomly chosen cells (unseeded in the reference) and the fixed matrix
    ``ranM(E, p, 5)``, cluster with and without log2; log is used iff ``msil[1] < 0.75 && msil[1] >= 0.95*msil[2]``."""
    ctx = ctx or get_context()
    e = Expression.wrap(scExp)
    ncells = e.n if ncells is None else int(ncells)
    sncells = min(100 if sncells is None else int(sncells), ncells)
    reind = r_sample_perm(ncells, _seed)
    cells = reind[:sncells] - 1
    rm = ctx.upload_rm([ranM2(e.m, int(p), 5)])
    prm = hc_params("ward.D", None, 2, 40, 0.0, 2.0)
    msil = []
    try:
        for logkind in (0, 2):
            proj = ctx.rp_project(e.m, e.n, rm, cells=cells, normalize=2 if e.normalize else 0, logkind=logkind,
                                  **e.project_kwargs())
            _, ms = ctx.getrowcolor(proj[0], prm)
            msil.append(ms)
    finally:
        rm.close()
    return bool(msil[0] < 0.75 and msil[0] >= 0.95 * msil[1])


# =====================================================================================================
# SHARP_small / SHARP_large
# =====================================================================================================
def _run(ctx, e: Expression, rm: RmDev, large, flag, ng, N_cluster, enpN, indN, hc, forview, reind, logkind=2,
         round_digits=-1, skip_smetac=False, block_max_n=0, shard=False):
    prm = RunParams(int(large), int(bool(flag)), int(logkind), int(round_digits), int(ng), _ncl(N_cluster), _ncl(enpN),
                    _ncl(indN), hc, 2 if e.normalize else 0, 1e6, int(bool(skip_smetac)), int(block_max_n), int(bool(shard)), 0)
    return ctx.run(rm, prm, reind=reind, want_vie=bool(forview), want_x0=bool(forview) and not skip_smetac and not shard,
                   max_x0_cols=max(64, hc.max_n + 1, _ncl(N_cluster) + 1), **e.run_kwargs())


def SHARP_small(scExp, ncells=None, ensize_K=15, reduced_ndim=None, hmethod="ward.D", N_cluster=None,
                indN_cluster=None, minN_cluster=2, maxN_cluster=40, sil_thre=0.35, height_Ntimes=2, flashmark=False,
                flag=True, n_cores=None, forview=True, rN_seed=0.5, ctx: Context | None = None, **kwargs) -> dict:
    """R/SHARP.R:339-454.  K members: RPmat with seed 50+rN.seed+k -> getrowColor; wMetaC over the K solutions;
    merge of tiny clusters (only when ncells > 1e4); relabel by first appearance."""
    k = _kw(kwargs)
    rN_seed = k.get("rN_seed", rN_seed)
    ctx = ctx or get_context()
    e = Expression.wrap(scExp)
    ncells = e.n if ncells is None else int(ncells)
    p = int(reduced_ndim if reduced_ndim is not None else math.ceil(math.log2(ncells) / 0.2 ** 2))
    K = int(ensize_K)
    hc = _hc(hmethod, None, minN_cluster, maxN_cluster, sil_thre, height_Ntimes, flashmark)
    rm = ctx.upload_rm(_rm_list(e.m, p, K, rN_seed))  # quirk B9: a supplied rM is ignored, same seeds
    try:
        r = _run(ctx, e, rm, 0, flag, 0, N_cluster, None, indN_cluster, hc, forview, None)
        labels = r["labels"]
        if _ncl(N_cluster) == 0 and ncells > 1e4:
            labels = _merge_small(labels)
        res = _finish(labels)
        if forview:
            allrp = []
            for j in range(K):
                col, inde = ctx.last_member(j, ncells, p)
                rc = np.array(colorL, dtype=object)[col - 1]
                allrp.append({"tag": f"_RP{p}_{j + 1}", "rowColor": rc, "N.cluster": int(len(np.unique(col))), "indE": inde})
            res["allrpinfo"] = allrp
            res["x0"] = r["x0"]
            res["viE"] = r["viE"]
    finally:
        rm.close()
    return res


def SHARP_large(scExp, ncells=None, ensize_K=5, reduced_dim=None, partition_ncells=2000, hmethod="ward.D",
                N_cluster=None, enpN_cluster=None, indN_cluster=None, minN_cluster=2, maxN_cluster=40, sil_thre=0.35,
                height_Ntimes=2, flashmark=False, flag=True, n_cores=None, forview=True, rM=True, rN_seed=0.5,
                ctx: Context | None = None, _logkind=2, _round_digits=-1, comm=None, **kwargs) -> dict:
    """R/SHARP.R:478-851.  Shuffle (iff ncells < 1e5), blocks of ``partition.ncells`` cells, K x T projections and
    block clusterings, per-block wMetaC, sMetaC across blocks, un-shuffle -- all in ONE device call (sharp_run);
    then the host glue: merge of clusters with < 10 cells (ncells > 1e4, no N.cluster), relabel by first appearance."""
    k = _kw(kwargs)
    rN_seed = k.get("rN_seed", rN_seed)
    ctx = ctx or get_context()
    e = Expression.wrap(scExp)
    ncells = e.n if ncells is None else int(ncells)
    p = int(reduced_dim if reduced_dim is not None else k.get("reduced_ndim", math.ceil(math.log2(ncells) / 0.2 ** 2)))
    K = int(ensize_K)
    hc = _hc(hmethod, None, minN_cluster, maxN_cluster, sil_thre, height_Ntimes, flashmark)
    # ``comm`` (sharp_b200.comm.NcclComm on ``ctx``): every rank makes this same call and the cell blocks -- the K x T nested
    # loop of R/SHARP.R:554-618 and the per-block wMetaC, :692-709 -- are dealt over the ranks; all ranks return the full result
    shard = comm is not None and comm.world > 1 and math.ceil(ncells / partition_ncells) >= comm.world
    if shard and rN_seed == 0.5:
        raise ValueError("a sharded run needs an integer rN.seed (every rank must draw the same ranM matrices and shuffle)")
    reind = _reind(ncells, rN_seed) if ncells < 1e5 else None  # drawn always in R, applied iff ncol(E) < 1e5
    rm, own = _as_rmdev(ctx, rM, e.m, p, K, rN_seed)
    try:
        r = _run(ctx, e, rm, 1, flag, partition_ncells, N_cluster, enpN_cluster, indN_cluster, hc, forview, reind,
                 _logkind, _round_digits, shard=shard)
    finally:
        if own:
            rm.close()
    labels = r["labels"]
    if _ncl(N_cluster) == 0 and ncells > 1e4:
        _cat("Adjust clusters with very small number of cells...")
        labels = _merge_small(labels)
    res = _finish(labels)
    if forview:
        res["viE"] = r["viE"]
        if shard:  # the soft indicator stays on the ranks that clustered the block; the hard one follows from the labels
            x0 = np.zeros((ncells, res["N.pred_cluster"]))
            x0[np.arange(ncells), res["pred_clusters"] - 1] = 1.0
            res["x0"] = x0
        else:
            res["x0"] = r["x0"]
    return res


# =====================================================================================================
# SHARP
# =====================================================================================================
def SHARP(scExp, exp_type=None, ensize_K=None, reduced_ndim=None, base_ncells=None, partition_ncells=None,
          hmethod=None, N_cluster=None, enpN_cluster=None, indN_cluster=None, minN_cluster=None, maxN_cluster=None,
          sil_thre=None, height_Ntimes=None, flashmark=False, logflag=None, sncells=None, n_cores=None, forview=True,
          prep=None, rM=None, rN_seed=None, rownames=None, ctx: Context | None = None, comm=None, **kwargs) -> dict:
    """R/SHARP.R:44-318.  ``n.cores`` is accepted and ignored (the GPU is the unit of parallelism); ``comm``: an optional
    :class:`sharp_b200.comm.NcclComm` -- the cell blocks of the SHARP_large path are then dealt over its ranks."""
    k = _kw(kwargs)
    g = lambda name, cur: k.get(name, cur)
    exp_type, ensize_K, reduced_ndim = g("exp_type", exp_type), g("ensize_K", ensize_K), g("reduced_ndim", reduced_ndim)
    base_ncells, partition_ncells = g("base_ncells", base_ncells), g("partition_ncells", partition_ncells)
    N_cluster, enpN_cluster, indN_cluster = g("N_cluster", N_cluster), g("enpN_cluster", enpN_cluster), g("indN_cluster", indN_cluster)
    minN_cluster, maxN_cluster = g("minN_cluster", minN_cluster), g("maxN_cluster", maxN_cluster)
    sil_thre, height_Ntimes, rN_seed = g("sil_thre", sil_thre), g("height_Ntimes", height_Ntimes), g("rN_seed", rN_seed)
    start = time.time()
    if scExp is None:
        raise ValueError("No expression data is provided!")
    ctx = ctx or get_context()
    e = Expression.wrap(scExp)
    ngenes, ncells = e.m, e.n
    _cat("Number of cells:", ncells, "\nNumber of genes:", ngenes)
    if prep is None:
        prep = ncells < 1e4
    if rownames is not None:  # duplicated gene names are dropped (R/SHARP.R:83-88)
        _, first = np.unique(np.asarray(rownames), return_index=True)
        keep = np.zeros(e.m, dtype=bool)
        keep[first] = True
        e = e.keep_rows(keep)
    if prep and e.dev is None:
        if e.any_negative():
            e = e.clamp_negative()
        e = e.keep_rows(e.row_sums() != 0)
    if exp_type is not None and exp_type not in ("CPM", "TPM"):
        e = e.with_normalize()
    if reduced_ndim is None:
        reduced_ndim = math.ceil(math.log2(ncells) / 0.2 ** 2)
    reduced_ndim = int(reduced_ndim)
    base_ncells = 5000 if base_ncells is None else base_ncells
    partition_ncells = 2000 if partition_ncells is None else partition_ncells
    hmethod = "ward.D" if hmethod is None else hmethod
    minN_cluster = 2 if minN_cluster is None else minN_cluster
    maxN_cluster = max(40, math.ceil(ncells / 5000)) if maxN_cluster is None else maxN_cluster
    sil_thre = 0.35 if sil_thre is None else sil_thre
    height_Ntimes = 2 if height_Ntimes is None else height_Ntimes
    rM = True if rM is None else rM
    rN_seed = _check_seed(rN_seed)
    if N_cluster is not None and ncells < base_ncells:  # R/SHARP.R:181-191
        indN_cluster = N_cluster
        base_ncells = math.ceil(ncells / 2)
        partition_ncells = math.ceil(ncells / 2)
        if ensize_K is None:
            ensize_K = 15
    if logflag is None:
        logflag = ncells < 1e4
    if logflag:
        flag = testlog(e, ncells, reduced_ndim, sncells, n_cores, ctx=ctx)
        _cat("Log-transform is necessary!" if flag else "Log-transform is not necessary!")
    else:
        flag = True
    if ncells < base_ncells:
        if ensize_K is None:
            ensize_K = 15
        enresults = SHARP_small(e, ncells, ensize_K, reduced_ndim, hmethod, N_cluster, indN_cluster, minN_cluster,
                                maxN_cluster, sil_thre, height_Ntimes, flashmark, flag, n_cores, forview, rN_seed, ctx=ctx)
    else:
        if ensize_K is None:
            ensize_K = 5
        enresults = SHARP_large(e, ncells, ensize_K, reduced_ndim, partition_ncells, hmethod, N_cluster, enpN_cluster,
                                indN_cluster, minN_cluster, maxN_cluster, sil_thre, height_Ntimes, flashmark, flag,
                                n_cores, forview, rM, rN_seed, ctx=ctx, comm=comm)
    enresults["N.cells"] = ncells
    enresults["N.genes"] = ngenes
    enresults["reduced.dim"] = reduced_ndim
    enresults["ensize.K"] = int(ensize_K)
    enresults["time"] = (time.time() - start) / 60.0  # minutes, like difftime(units = "mins")
    enresults["paras"] = {"ensize.K": int(ensize_K), "reduced.ndim": reduced_ndim, "base.ncells": base_ncells,
                          "partition.ncells": partition_ncells, "logmark": bool(flag), "hmethod": hmethod,
                          "N.cluster": N_cluster, "minN.cluster": minN_cluster, "maxN.cluster": maxN_cluster,
                          "sil.thre": sil_thre, "height.Ntimes": height_Ntimes, "n.cores": n_cores}
    return enresults


def run_Mtimes_SHARP(scExp, Mtimes=10, ensize_K=(15,), **kwargs) -> dict:
    """R/run_Mtimes_SHARP.R:22-60: repeat SHARP() Mtimes per ensemble size."""
    out = {}
    for K in ensize_K:
        out[int(K)] = [SHARP(scExp, ensize_K=int(K), **kwargs) for _ in range(int(Mtimes))]
    return out


# =====================================================================================================
# SHARP_unlimited family
# =====================================================================================================
def _relabel_by_size(labels: np.ndarray) -> np.ndarray:
    """x = sort(table(f), decreasing = TRUE); map names(x) -> 1..length(x)   (R/SHARP_unlimited.R:180-183).
    table() orders its names as STRINGS and sort(decreasing = TRUE) is order(., decreasing = TRUE) (stable), so
    equal counts keep the string order of the ids."""
    vals, cnt = _counts(labels)
    names = sorted(range(len(vals)), key=lambda q: str(int(vals[q])))
    order = sorted(names, key=lambda q: -cnt[q])  # Python's sort is stable
    lut = np.zeros(int(vals.max()) + 1, dtype=np.int32)
    for i, q in enumerate(order):
        lut[int(vals[q])] = i + 1
    return lut[np.asarray(labels, dtype=np.int64)]


def _exchange_parts(comm, own, y, cens, nnp):
    """ONE allgather for everything the global step needs from the other ranks: the labels and centroids of the parts in
    `own`, and part 1's parameter list from the rank that holds it (the `.combine` of R/SHARP_unlimited3.R:137-147)"""
    import pickle
    payload = {i: y[i]["pred_clusters"] for i in own}
    payload.update({nnp + i: cens[i] for i in own})
    if 0 in own:
        head = pickle.dumps({q: y[0][q] for q in ("reduced.dim", "ensize.K", "paras")})
        payload[2 * nnp] = np.frombuffer(head, dtype=np.uint8)
    got = comm.allgather_parts(payload, 2 * nnp + 1)
    return got[:nnp], got[nnp:2 * nnp], pickle.loads(got[2 * nnp].tobytes())


def _unlimited_combine(ctx, cen, counts_per_part, preds, ncells, hmethod, N_cluster, minN, maxN, sil_thre,
                       height_Ntimes) -> np.ndarray:
    """global sMetaC on the part-level cluster centroids + merge + relabel (R/SHARP_unlimited.R:151-183); the label
    tail (tf[fColor], clusters under 10 cells merged, ids by decreasing size) is one native host pass
    (sharp_labels_combine; `_combine_labels_py` is the same thing in numpy, kept as its test oracle)"""
    hc = _hc(hmethod, N_cluster, minN, maxN, sil_thre, height_Ntimes)
    t0 = time.time()
    tf = ctx.smetac_centroids(cen, ncells, hc)
    t1 = time.time()
    merge = 10 if (_ncl(N_cluster) == 0 and ncells > 1e4) else 0
    final, sizes = _lib.labels_combine(preds, counts_per_part, tf, merge)
    if _TRACE:
        print(f"[sharp trace py] global sMetaC on {cen.shape[0]} centroids {1e3 * (t1 - t0):.2f} ms, label tail "
              f"{1e3 * (time.time() - t1):.2f} ms", file=sys.stderr)
    return _SizedLabels(final, sizes)


class _SizedLabels(np.ndarray):
    """an int32 label vector that carries the sizes of its labels 1..L (saves a counting pass over 1.3 M cells)"""
    def __new__(cls, arr, sizes):
        obj = np.asarray(arr).view(cls)
        obj.sizes = sizes
        return obj

    def __array_finalize__(self, obj):
        self.sizes = getattr(obj, "sizes", None)


def _combine_labels_py(tf, counts_per_part, preds, ncells, merge: bool) -> np.ndarray:
    final = np.empty(ncells, dtype=np.int32)
    pos, off = 0, 0
    for pred, nk in zip(preds, counts_per_part):
        final[pos:pos + len(pred)] = tf[off + pred - 1]
        pos += len(pred)
        off += nk
    if merge:
        final = _merge_small(final)
    return _relabel_by_size(final)


def _unlimited_result(final, ncells, ngenes, y0, start):
    sizes = getattr(final, "sizes", None)
    if sizes is not None:
        final = np.asarray(final)
        uf, ufc = np.arange(1, len(sizes) + 1, dtype=np.int32), sizes
    else:
        uf, ufc = _counts(final)
    return {"pred_clusters": final, "unique_pred_clusters": uf.astype(np.int64),
            "distr_pred_clusters": {int(v): int(c) for v, c in zip(uf, ufc)},
            "N.pred_clusters": int(len(uf)), "N.cells": int(ncells), "N.genes": int(ngenes),
            "reduced.dim": y0["reduced.dim"], "ensize.K": y0["ensize.K"], "time": (time.time() - start) / 60.0,
            "paras": y0["paras"]}


_FUSED_KEYS = {"exp_type", "base_ncells", "partition_ncells", "hmethod", "sil_thre", "height_Ntimes", "flashmark",
               "enpN_cluster", "indN_cluster"}
_fused_parts = True     # module switch: False forces the part-by-part path (tests compare the two)
_fused_group = 0        # parts per group / groups in flight of sharp_run_parts (0 = library default)
_fused_lanes = 0


def _parts_fast_path(parts, mine, k, viewflag, part_logflag, n_streams):
    """The arguments of the per-part SHARP() calls when ALL of them take the SHARP_large path with identical
    parameters (then sharp_run_parts runs them as one fused device call), else None."""
    if not _fused_parts or viewflag or part_logflag or isinstance(n_streams, (list, tuple)) or not mine:
        return None
    if set(k) - _FUSED_KEYS:
        return None
    base = 5000 if k.get("base_ncells") is None else k["base_ncells"]
    ns = [parts[i].n for i in mine]
    if min(ns) < base or min(ns) < 1e4:   # SHARP_small, or the default logflag / prep would switch on
        return None
    if len({max(40, math.ceil(n / 5000)) for n in ns}) != 1:   # maxN.cluster differs between the parts
        return None
    if any(parts[i].dev is None and parts[i].dense is None and parts[i].csc is None for i in mine):
        return None
    return {"exp_type": k.get("exp_type"), "base_ncells": base,
            "partition_ncells": 2000 if k.get("partition_ncells") is None else k["partition_ncells"],
            "hmethod": "ward.D" if k.get("hmethod") is None else k["hmethod"],
            "sil_thre": 0.35 if k.get("sil_thre") is None else k["sil_thre"],
            "height_Ntimes": 2 if k.get("height_Ntimes") is None else k["height_Ntimes"],
            "flashmark": bool(k.get("flashmark", False)), "enpN_cluster": k.get("enpN_cluster"),
            "indN_cluster": k.get("indN_cluster"), "maxN_cluster": max(40, math.ceil(ns[0] / 5000))}


def _fused_inputs(parts, mine):
    """the parts of this rank as sharp_run_parts / sharp_parts_prefetch take them"""
    ins = []
    for i in mine:
        e = parts[i]
        if e.dev is not None:
            ins.append(e.dev)
        elif e.dense is not None:
            ins.append({"n": e.n, "dense": e.dense})
        else:
            ins.append({"n": e.n, "csc": e.csc})
    return ins


def _run_parts_fused(ctx, parts, mine, rM, p, K, rN_seed, a, n_cores, shared=frozenset(), reinds=None):
    """What the loop `y[[i]] = SHARP(scExp[[i]], reduced.ndim = p, prep = FALSE, logflag = FALSE, rM = rM, ...)` returns
    for the parts in ``mine`` (R/SHARP_unlimited.R:125-149), computed by ONE sharp_run_parts call."""
    normalize = a["exp_type"] is not None and a["exp_type"] not in ("CPM", "TPM")
    hc = _hc(a["hmethod"], None, 2, a["maxN_cluster"], a["sil_thre"], a["height_Ntimes"], a["flashmark"])
    prm = RunParams(1, 1, 2, -1, int(a["partition_ncells"]), 0, _ncl(a["enpN_cluster"]), _ncl(a["indN_cluster"]), hc,
                    2 if normalize else 0, 1e6)
    ins = _fused_inputs(parts, mine)
    for i in mine:
        _cat("Processing Partition", i + 1, "of the scRNA-seq data...")
    if reinds is None:
        reinds = _reinds_for([parts[i].n for i in mine], rN_seed)
    elif callable(reinds):   # a future started earlier (drawn while the ranM matrices were)
        reinds = reinds()
    start = time.time()
    outs = ctx.run_parts(rM, prm, parts[mine[0]].m, ins, reinds, small_thre=10, cen_cap=max(64, a["maxN_cluster"] + 1),
                         group=_fused_group, lanes=_fused_lanes, sharded=[i in shared for i in mine])
    res = []
    for i, o in zip(mine, outs):
        cid = o["pred_clusters"]
        vals, cnt = _counts(cid)
        r = {"pred_clusters": cid, "unique_pred_clusters": vals.astype(np.int64),
             "distr_pred_clusters": {int(v): int(c) for v, c in zip(vals, cnt)}, "N.pred_cluster": int(len(vals)),
             "N.cells": parts[i].n, "N.genes": parts[i].m, "reduced.dim": int(p), "ensize.K": int(K),
             "time": (time.time() - start) / 60.0,
             "paras": {"ensize.K": int(K), "reduced.ndim": int(p), "base.ncells": a["base_ncells"],
                       "partition.ncells": a["partition_ncells"], "logmark": True, "hmethod": a["hmethod"],
                       "N.cluster": None, "minN.cluster": 2, "maxN.cluster": a["maxN_cluster"],
                       "sil.thre": a["sil_thre"], "height.Ntimes": a["height_Ntimes"], "n.cores": n_cores},
             "cen": o["cen"]}
        res.append(r)
    return res


def SHARP_unlimited(scExp, viewflag=True, n_cores=None, ensize_K=None, N_cluster=None, minN_cluster=None,
                    maxN_cluster=None, rN_seed=None, ctx: Context | None = None, comm=None, n_streams=None,
                    _part_logflag=False, _krange_from_part1=False, **kwargs) -> dict:
    """R/SHARP_unlimited.R:29-242.  ``scExp``: a LIST of genes x cells matrices (parts).  Every part runs SHARP()
    with the shared ranM matrices; the part-level clusters are merged by one global sMetaC over their centroids
    (computed on the device from the part's viE, which never leaves it unless ``viewflag``).

    ``comm``: optional :class:`sharp_b200.comm.NcclComm` -- the parts are then sharded over the ranks (one process per
    GPU) and only centroids, counts and labels are exchanged (allgather); every rank returns the full result."""
    k = _kw(kwargs)
    rN_seed, ensize_K = k.pop("rN_seed", rN_seed), k.pop("ensize_K", ensize_K)
    N_cluster = k.pop("N_cluster", N_cluster)
    minN_cluster, maxN_cluster = k.pop("minN_cluster", minN_cluster), k.pop("maxN_cluster", maxN_cluster)
    start = time.time()
    if scExp is None:
        raise ValueError("No expression data is provided!")
    if not isinstance(scExp, (list, tuple)):
        if isinstance(scExp, np.ndarray):
            return SHARP(scExp, ctx=ctx)  # "SHARP is used instead of SHARP_unlimited because the input is a matrix!"
        raise ValueError("The input should be a LIST of partitioned scRNA-seq expression matrices!")
    if len(scExp) == 1:
        return SHARP(scExp[0], ctx=ctx)
    ctx = ctx or get_context()
    parts = [Expression.wrap(x) for x in scExp]
    nnp = len(parts)
    nnc = [e.n for e in parts]
    ncells = int(sum(nnc))
    p = math.ceil(math.log2(ncells) / 0.2 ** 2)
    minN_cluster = 2 if minN_cluster is None else minN_cluster
    maxN_cluster = max(40, math.ceil(ncells / 5000)) if maxN_cluster is None else maxN_cluster
    rN_seed = _check_seed(rN_seed, allow_half=False)
    ensize_K = 5 if ensize_K is None else int(ensize_K)
    _t = [time.time()]

    def _mark(what):
        if _TRACE:
            now = time.time()
            print(f"[sharp trace py] {what} {1e3 * (now - _t[0]):.2f} ms", file=sys.stderr)
            _t[0] = now

    if _part_logflag is None and "logflag" in k:   # SHARP_unlimited3 passes no logflag itself (:114), so `...` may
        _part_logflag = k.pop("logflag")
    rank, world = (comm.rank, comm.world) if comm is not None else (0, 1)
    # Several ranks: whole parts are dealt round-robin while they divide evenly; the parts that are left
    # over -- 26 parts on 8 ranks leave 2 -- are BLOCK-sharded: every rank clusters a share of their cell blocks and the
    # library allgathers the block-level results (sharp_part.sharded), so that all ranks carry the same load.  That needs
    # the fused path, an integer seed, the left-over parts' data on every rank and at least `world` blocks per part.
    nwhole = nnp
    shared: list = []
    if world > 1 and nnp % world:
        cand = list(range((nnp // world) * world, nnp))
        ng = 2000 if k.get("partition_ncells") is None else int(k["partition_ncells"])
        ok = (rN_seed != 0.5 and getattr(comm, "backend", "") == "nccl" and
              all(parts[i].dev is not None or parts[i].dense is not None or parts[i].csc is not None for i in cand) and
              all(math.ceil(parts[i].n / ng) >= world for i in cand) and
              _parts_fast_path(parts, cand, k, viewflag, _part_logflag, n_streams) is not None)
        ok = all(b == b"1" for b in comm.allgather_bytes(b"1" if ok else b"0"))   # the same decision on every rank
        if ok:
            nwhole, shared = cand[0], cand
    mine = [i for i in range(nwhole) if i % world == rank] + shared
    fast = _parts_fast_path(parts, mine, k, viewflag, _part_logflag, n_streams)
    if shared and fast is None:
        raise ValueError("the block-sharded parts need the fused path on every rank (same parameters as the rank's own parts)")
    # host buffers: the copy of the first parts starts now and runs while the ranM matrices are drawn on the host
    _pf_keep = None
    if fast is not None and mine and parts[mine[0]].dev is None:
        _pf_keep = ctx.parts_prefetch(parts[mine[0]].m, _fused_inputs(parts, mine), _fused_group, _fused_lanes,
                                      sharded=[i in shared for i in mine])
    _reind_job = None
    if fast is not None and mine:   # the parts' shuffles are drawn on a host thread while the ranM matrices are
        import threading
        _box: dict = {}

        def _draw():
            try:
                _box["v"] = _reinds_for([parts[i].n for i in mine], rN_seed)
            except BaseException as ex:  # surfaced by the consumer
                _box["e"] = ex

        _th = threading.Thread(target=_draw, daemon=True)
        _th.start()

        def _reind_job():
            _th.join()
            if "e" in _box:
                raise _box["e"]
            return _box["v"]
    rms_host = _rm_list(parts[0].m, p, ensize_K, rN_seed, comm)
    _mark("ranM")
    rM = ctx.upload_rm(rms_host)
    _mark("upload_rm")
    y, cens, viEs = {}, {}, {}
    if isinstance(n_streams, (list, tuple)):  # explicit contexts, one per stream
        ctxs = list(n_streams)[:max(1, len(mine))]
    else:
        S = max(1, min(int(n_streams or _default_streams), len(mine)))
        if ctx is get_context(ctx.device) if _contexts.get(ctx.device) is ctx else False:
            ctxs = stream_contexts(S, ctx.device)
        else:  # a caller-owned context: its siblings (same type, same device) are created once and kept on it
            sib = ctx.__dict__.setdefault("_siblings", [])
            while len(sib) < S - 1:
                sib.append(type(ctx)(ctx.device))
            ctxs = [ctx] + sib[:S - 1]

    def one_part(i, c):
        _cat("Processing Partition", i + 1, "of the scRNA-seq data...")
        t0 = time.time()
        yi = SHARP(parts[i], reduced_ndim=p, prep=False, logflag=_part_logflag, n_cores=n_cores, rM=rM, ensize_K=ensize_K,
                   rN_seed=rN_seed, forview=False, ctx=c, **k)
        t1 = time.time()
        cen, _ = c.centroids(yi["pred_clusters"], yi["N.pred_cluster"], p)
        if _TRACE:
            print(f"[sharp trace py] part {i}: SHARP {1e3 * (t1 - t0):.2f} ms, centroids {1e3 * (time.time() - t1):.2f} ms",
                  file=sys.stderr)
        return yi, cen, (c.last_vie(nnc[i], p) if viewflag else None)

    try:
        if fast is not None:  # every part takes the SHARP_large path with the same parameters: one fused device call
            t0 = time.time()
            outs = _run_parts_fused(ctx, parts, mine, rM, p, ensize_K, rN_seed, fast, n_cores, set(shared), reinds=_reind_job)
            for i, o in zip(mine, outs):
                y[i], cens[i], viEs[i] = o, o.pop("cen"), None
            if _TRACE:
                print(f"[sharp trace py] fused parts {1e3 * (time.time() - t0):.2f} ms", file=sys.stderr)
        elif len(ctxs) == 1:
            for i in mine:
                y[i], cens[i], viEs[i] = one_part(i, ctx)
        else:  # one host thread per stream; the C ABI calls release the GIL
            import queue
            from concurrent.futures import ThreadPoolExecutor
            free = queue.SimpleQueue()
            for c in ctxs:
                free.put(c)

            def job(i):
                c = free.get()
                try:
                    return one_part(i, c)
                finally:
                    free.put(c)

            with ThreadPoolExecutor(max_workers=len(ctxs)) as ex:
                for i, r in zip(mine, ex.map(job, mine)):
                    y[i], cens[i], viEs[i] = r
    finally:
        rM.close()
    _mark("parts")
    if comm is not None:
        own = [i for i in mine if i not in shared or rank == 0]   # a block-sharded part is complete on every rank: rank 0 contributes it
        preds_all, cens_all, y0 = _exchange_parts(comm, own, y, cens, nnp)
        if viewflag:
            viEs = comm.allgather_parts({i: viEs[i] for i in own}, nnp)
    else:
        preds_all = [y[i]["pred_clusters"] for i in range(nnp)]
        cens_all = [cens[i] for i in range(nnp)]
        y0 = y[0]
    cen = np.ascontiguousarray(np.concatenate(cens_all, axis=0))
    _cat("Total number of single cells:", ncells, "\nNumber of unique meta-clusters:", cen.shape[0])
    if _krange_from_part1:  # SHARP_unlimited3 hands y[[1]]$paras$minN.cluster / maxN.cluster to sMetaC (R/SHARP_unlimited3.R:165-166)
        minN_cluster, maxN_cluster = y0["paras"]["minN.cluster"], y0["paras"]["maxN.cluster"]
    final = _unlimited_combine(ctx, cen, [c.shape[0] for c in cens_all], preds_all, ncells, y0["paras"]["hmethod"],
                               N_cluster, minN_cluster, maxN_cluster, y0["paras"]["sil.thre"], y0["paras"]["height.Ntimes"])
    _mark("combine")
    res = _unlimited_result(final, ncells, parts[0].m, y0, start)
    # N.pred_clusters / reduced.dim: quirk B5 -- the reference copies the nonexistent y[[1]]$reduced.ndim (NULL)
    if viewflag:
        E1 = np.concatenate([viEs[i] for i in range(nnp)], axis=0)
        if ncells > 1e5:
            # quirk B3: the reference uses an undefined `k` here; foreach leaves none, we use k = ensize.K
            z0 = ranM2(p, 50, _member_seeds(1, rN_seed, comm)[0] if rN_seed == 0.5 else _member_seed(rN_seed, ensize_K))
            res["viE"] = _view_project(ctx, E1, z0)
        else:
            res["viE"] = E1
        x0 = np.zeros((ncells, res["N.pred_clusters"]))
        x0[np.arange(ncells), final - 1] = 1.0
        res["x0"] = x0
    return res


def _view_project(ctx, E1, z0):
    """as.matrix(1/sqrt(kdim) * E1 %*% z0): E1 (ncells x p, row-major) is the column-major image of t(E1), so this is
    the projection kernel applied to a p x ncells "expression" matrix (R/SHARP_unlimited.R:216-226)."""
    n, p = E1.shape
    rm = ctx.upload_rm([z0])
    try:
        out = ctx.rp_project(p, n, rm, dense=np.asfortranarray(E1.T), logkind=0)
    finally:
        rm.close()
    return out[0]


def SHARP_fpart(scExp, ensize_K=5, reduced_ndim=None, partition_ncells=2000, hmethod="ward.D", N_cluster=None,
                enpN_cluster=None, indN_cluster=None, minN_cluster=2, maxN_cluster=40, sil_thre=0.35, height_Ntimes=2,
                flag=True, rM=True, rN_seed=0.5, ctx: Context | None = None, **kwargs) -> dict:
    """R/SHARP_unlimited2.R:297-544: the BLOCK stage of SHARP_unlimited2 for one part.  Like the first half of SHARP_large
    -- shuffle (iff ncells < 1e5), blocks, K x T projections and block clusterings, per-block wMetaC -- but with log10
    instead of log2 (:391), projections rounded to one decimal (:410), per-block maxN.cluster = 40 (:421), and it STOPS
    there: no sMetaC across the blocks.  Returns, un-shuffled (:520-524), ``fColor`` (one id per (block, meta-cluster)
    pair: an integer code of the reference's ``"<finalC>en<t>"`` strings), ``E1`` = enE/K, ``nmcluster``, ``folds``,
    ``ncells``, ``ngenes`` (:529-535).  ``N.cluster`` is accepted and unused, like in the reference."""
    k = _kw(kwargs)
    rN_seed = k.get("rN_seed", rN_seed)
    ctx = ctx or get_context()
    e = Expression.wrap(scExp)
    p = int(reduced_ndim if reduced_ndim is not None else math.ceil(math.log2(e.n) / 0.2 ** 2))
    ng = int(partition_ncells)
    # `maxN.cluster = 40` is assigned INSIDE the (member, block) worker (:421): it bounds the block clusterings only; the
    # per-block wMetaC (:481) runs in the function's own frame with the caller's maxN.cluster
    hc = _hc(hmethod, None, minN_cluster, maxN_cluster, sil_thre, height_Ntimes)
    reind = _reind(e.n, rN_seed) if e.n < 1e5 else None                 # (the reference draws it always, applies it iff ncells < 1e5)
    rm, own = _as_rmdev(ctx, rM, e.m, p, int(ensize_K), rN_seed)
    try:
        r = _run(ctx, e, rm, 1, flag, ng, None, enpN_cluster, indN_cluster, hc, True, reind, 10, 1, skip_smetac=True,
                 block_max_n=40)
    finally:
        if own:
            rm.close()
    folds = _folds(e.n, ng)
    if e.n < 1e5:
        f = np.empty_like(folds)
        f[reind - 1] = folds                                            # folds[reind] = folds (:523)
        folds = f
    fColor = np.asarray(r["labels"])
    return {"fColor": fColor, "E1": r["viE"], "nmcluster": int(len(np.unique(fColor))), "folds": folds,
            "ncells": e.n, "ngenes": e.m}


def _folds(ncells: int, ng: int) -> np.ndarray:
    """R/SHARP.R:513-536 == R/SHARP_unlimited2.R:339-358: cut(seq(1, T*ng), breaks = T) with the last two folds averaged"""
    T = int(math.ceil(ncells / ng))
    if T <= 1:
        return np.ones(ncells, dtype=np.int64)
    folds = np.repeat(np.arange(1, T + 1, dtype=np.int64), ng)
    nt = ncells - (T - 2) * ng
    nind = np.flatnonzero(folds == T - 1)
    folds[nind[nt // 2:]] = T
    return folds[:ncells]


def SHARP_unlimited2(scExp, ensize_K=None, reduced_ndim=None, partition_ncells=None, hmethod=None, N_cluster=None,
                     enpN_cluster=None, indN_cluster=None, minN_cluster=None, maxN_cluster=None, sil_thre=None,
                     height_Ntimes=None, logflag=None, n_cores=None, forview=True, rN_seed=None,
                     ctx: Context | None = None, **kwargs) -> dict:
    """R/SHARP_unlimited2.R:29-267: the TWO-level variant over a LIST of parts: every part goes through SHARP_fpart
    (blocks + per-block wMetaC, log10 / round(., 1) projections, shared ``ranM2(ngenes, p, .)`` matrices, :133-139),
    then ONE global sMetaC merges the block-level clusters of all parts (:183-185), followed by the small-cluster merge
    and the relabel by decreasing size (:189-204).  The centroids the global sMetaC needs (colMeans of E1 per cluster,
    R/sMetaC.R:58-63) are taken on the device right after each part, so E1 only comes back when ``forview``."""
    k = _kw(kwargs)
    rN_seed = k.get("rN_seed", rN_seed)
    start = time.time()
    if scExp is None:
        raise ValueError("No expression data is provided!")
    if not isinstance(scExp, (list, tuple)):
        raise ValueError("The input should be a LIST of partitioned scRNA-seq expression matrices!")
    ctx = ctx or get_context()
    parts = [Expression.wrap(x) for x in scExp]
    ncells = int(sum(e.n for e in parts))
    p = int(reduced_ndim if reduced_ndim is not None else math.ceil(math.log2(ncells) / 0.2 ** 2))
    ensize_K = 5 if ensize_K is None else int(ensize_K)
    partition_ncells = 2000 if partition_ncells is None else int(partition_ncells)
    hmethod = "ward.D" if hmethod is None else hmethod
    minN_cluster = 2 if minN_cluster is None else minN_cluster
    maxN_cluster = max(40, math.ceil(ncells / 5000)) if maxN_cluster is None else maxN_cluster
    sil_thre = 0.35 if sil_thre is None else sil_thre
    height_Ntimes = 2 if height_Ntimes is None else height_Ntimes
    rN_seed = _check_seed(rN_seed)
    if logflag is None:
        logflag = ncells < 1e4
    if logflag:  # :92-103: testlog on the first part, 100 cells
        flag = testlog(parts[0], parts[0].n, p, 100, n_cores, ctx=ctx)
    else:
        flag = True
    rM = ctx.upload_rm(_rm_list(parts[0].m, p, ensize_K, rN_seed))
    preds, cens, E1s = [], [], []
    try:
        for e in parts:
            r = SHARP_fpart(e, ensize_K, p, partition_ncells, hmethod, N_cluster, enpN_cluster, indN_cluster, minN_cluster,
                            maxN_cluster, sil_thre, height_Ntimes, flag, rM, rN_seed, ctx=ctx)
            codes, _ = _first_appearance_codes(r["fColor"])     # unique(paste(fColor, "s", i)) within part i
            cen, _ = ctx.centroids(codes, int(codes.max()), p)
            preds.append(codes)
            cens.append(cen)
            if forview:
                E1s.append(r["E1"])
    finally:
        rM.close()
    cen = np.ascontiguousarray(np.concatenate(cens, axis=0))
    final = _unlimited_combine(ctx, cen, [c.shape[0] for c in cens], preds, ncells, hmethod, N_cluster, minN_cluster,
                               maxN_cluster, sil_thre, height_Ntimes)
    y0 = {"reduced.dim": p, "ensize.K": ensize_K,
          "paras": {"ensize.K": ensize_K, "reduced.ndim": p, "partition.ncells": partition_ncells, "logmark": flag,
                    "hmethod": hmethod, "N.cluster": N_cluster, "minN.cluster": minN_cluster,
                    "maxN.cluster": maxN_cluster, "sil.thre": sil_thre, "height.Ntimes": height_Ntimes,
                    "n.cores": n_cores}}
    res = _unlimited_result(final, ncells, parts[0].m, y0, start)
    res["reduced.ndim"] = p                                   # :233 (this driver does name it reduced.ndim)
    if forview:                                               # :235-243
        res["viE"] = np.concatenate(E1s, axis=0)
        x0 = np.zeros((ncells, res["N.pred_clusters"]))
        x0[np.arange(ncells), final - 1] = 1.0
        res["x0"] = x0
    return res


def _first_int(name: str) -> int:
    mt = re.search(r"\d+", name)
    return int(mt.group()) if mt else 0


def SHARP_unlimited3(ndinfo, viewflag=True, n_cores=None, ensize_K=None, rN_seed=None, N_cluster=None,
                     reader=None, ctx: Context | None = None, comm=None, **kwargs) -> dict:
    """R/SHARP_unlimited3.R:29-235: on-disk variant.  ``ndinfo = {"dir": ..., "ncells": ..., "ngenes": ...}``; the
    files of ``dir`` are processed in the order of the first integer in their names (:59-61), one part at a time
    (read, cluster, free).  ``reader(path)`` must return a genes x cells matrix (default: ``.npz`` files holding
    dgCMatrix slots ``p, i, x, Dim``; R's ``readRDS`` stays in R where R exists)."""
    import os
    k = _kw(kwargs)
    if not isinstance(ndinfo, dict) or "dir" not in ndinfo:
        raise ValueError("The input should be a list containing the directory, the number of cells and genes!")
    files = sorted(os.listdir(ndinfo["dir"]), key=_first_int)
    paths = [os.path.join(ndinfo["dir"], f) for f in files]

    def default_reader(path):
        z = np.load(path)
        return {"p": z["p"], "i": z["i"], "x": z["x"], "Dim": tuple(int(v) for v in z["Dim"])}

    _user_reader = reader
    reader = reader or default_reader

    class _Lazy(Expression):
        """read on first use, dropped after the part is done (R/SHARP_unlimited3.R:105,124-125)"""

        def __init__(self, path, m, n):
            super().__init__(m, n)
            self._path = path

        def _load(self):
            e = Expression.wrap(reader(self._path))
            self.dense, self.csc, self.m, self.n = e.dense, e.csc, e.m, e.n

        def run_kwargs(self):
            if self.dense is None and self.csc is None:
                self._load()
            kw = super().run_kwargs()
            self.dense = self.csc = None          # rm(mat); gc() (:124-125): the run is the part's last use
            return kw

        def project_kwargs(self):                 # testlog (parts below 1e4 cells, :114 passes no logflag) reads it first
            if self.dense is None and self.csc is None:
                self._load()
            return super().project_kwargs()

        def any_negative(self):
            return False

    if _user_reader is None and paths and all(pth.endswith(".csc") for pth in paths):
        # raw dgCMatrix files: the streamed path -- batches of parts go through the fused loop over parts while the
        # native reader loads the next batch into pinned buffers
        return _unlimited3_streamed(paths, ndinfo, viewflag, n_cores, ensize_K, rN_seed, N_cluster, ctx, comm, k)
    ncells_each = ndinfo.get("ncells_each")
    if ncells_each is None:  # one metadata pass, like dim(readRDS(.)) in the reference's first loop
        ncells_each = [Expression.wrap(reader(pth)).n for pth in paths]
    parts = [_Lazy(pth, int(ndinfo["ngenes"]), int(nc)) for pth, nc in zip(paths, ncells_each)]
    return SHARP_unlimited(parts, viewflag=viewflag, n_cores=n_cores, ensize_K=ensize_K, N_cluster=N_cluster,
                           rN_seed=rN_seed, ctx=ctx, comm=comm, _part_logflag=None, _krange_from_part1=True,
                           **k)  # :114 passes no logflag


def _unlimited3_streamed(paths, ndinfo, viewflag, n_cores, ensize_K, rN_seed, N_cluster, ctx, comm, k) -> dict:
    """SHARP_unlimited3 over SHCSC001 files (sharp_b200/io.py): R/SHARP_unlimited3.R:103-131 with the per-part SHARP() calls
    fused (sharp_run_parts) batch by batch, reading batch b + 1 from disk while batch b is uploaded and clustered; then
    the global sMetaC with part 1's k-range (:165-166), merge, relabel by size.  Parts that would not take the SHARP_large
    path (fewer than 1e4 cells) fall back to the generic driver with a reader callback."""
    from . import io as sio
    start = time.time()
    ctx = ctx or get_context()
    infos = [sio.file_info(pth) for pth in paths]               # headers only
    nnp = len(paths)
    nnc = [i[1] for i in infos]
    m = infos[0][0]
    ncells = int(sum(nnc))
    if min(nnc) < 1e4 or viewflag or len({max(40, math.ceil(n / 5000)) for n in nnc}) != 1:
        def reader(pth):
            mm, n, (cp, ri, v) = sio.PartSlot(*sio.file_info(pth)[1:], pinned=False).load(pth)
            return {"p": cp, "i": ri, "x": v, "Dim": (mm, n)}
        return SHARP_unlimited3(dict(ndinfo, ncells_each=nnc), viewflag, n_cores, ensize_K, rN_seed, N_cluster, reader=reader,
                                ctx=ctx, comm=comm, **k)
    p = math.ceil(math.log2(ncells) / 0.2 ** 2)
    rN_seed = _check_seed(rN_seed, allow_half=False)
    ensize_K = 5 if ensize_K is None else int(ensize_K)
    rank, world = (comm.rank, comm.world) if comm is not None else (0, 1)
    mine = [i for i in range(nnp) if i % world == rank]        # files are dealt round-robin: a rank reads only its own
    kk = dict(k)
    kk.pop("logflag", None)
    batch = max(1, int(kk.pop("_batch", 8)))
    parts_meta = [Expression(m, n, csc=()) for n in nnc]        # shapes only (the fast-path test looks at sizes and data presence)
    fast = _parts_fast_path(parts_meta, list(range(nnp)), kk, False, False, None)
    if fast is None:
        raise ValueError("SHARP_unlimited3 (streamed): unsupported argument for the fused path: " + ", ".join(sorted(set(kk) - _FUSED_KEYS)))
    rM = ctx.upload_rm(_rm_list(m, p, ensize_K, rN_seed, comm))
    rd = sio.BatchReader([paths[i] for i in mine], [infos[i] for i in mine], batch, pinned=_lib.device_count() > 0)
    y, cens = {}, {}
    try:
        for b, loaded in rd:
            idx = mine[b * batch:(b + 1) * batch]
            plist = list(parts_meta)
            for i, (mm, n, csc) in zip(idx, loaded):
                plist[i] = Expression(mm, n, csc=csc)
            for i, o in zip(idx, _run_parts_fused(ctx, plist, idx, rM, p, ensize_K, rN_seed, fast, n_cores)):
                y[i], cens[i] = o, o.pop("cen")
    finally:
        rd.close()
        rM.close()
    if comm is not None:
        preds_all, cens_all, y0 = _exchange_parts(comm, list(mine), y, cens, nnp)
    else:
        preds_all, cens_all, y0 = [y[i]["pred_clusters"] for i in range(nnp)], [cens[i] for i in range(nnp)], y[0]
    cen = np.ascontiguousarray(np.concatenate(cens_all, axis=0))
    final = _unlimited_combine(ctx, cen, [c.shape[0] for c in cens_all], preds_all, ncells, y0["paras"]["hmethod"], N_cluster,
                               y0["paras"]["minN.cluster"], y0["paras"]["maxN.cluster"], y0["paras"]["sil.thre"],
                               y0["paras"]["height.Ntimes"])
    return _unlimited_result(final, ncells, m, y0, start)
