"""ctypes driver for the CPU oracle (oracle/libsharp_oracle.so) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(_ROOT, "oracle", "libsharp_oracle.so")

WARD_D, SINGLE, COMPLETE, AVERAGE, MCQUITTY, MEDIAN, CENTROID, WARD_D2 = 1, 2, 3, 4, 5, 6, 7, 8


class HcParams(C.Structure):
    _fields_ = [("hmethod", C.c_int), ("n_cluster", C.c_int), ("min_n", C.c_int), ("max_n", C.c_int),
                ("sil_thre", C.c_double), ("height_ntimes", C.c_double)]


class SharpParams(C.Structure):
    _fields_ = [("large", C.c_int), ("logflag", C.c_int), ("ensize_k", C.c_int), ("p", C.c_int),
                ("partition_ncells", C.c_int), ("n_cluster", C.c_int), ("enp_n_cluster", C.c_int),
                ("ind_n_cluster", C.c_int), ("hc", HcParams), ("logkind", C.c_int), ("round_digits", C.c_int)]


def hc_params(hmethod=WARD_D, n_cluster=0, min_n=2, max_n=40, sil_thre=0.35, height_ntimes=2.0):
    return HcParams(hmethod, n_cluster or 0, min_n, max_n, sil_thre, height_ntimes)


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(_ROOT, "oracle")])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.oracle_last_error.restype = C.c_char_p
    return _lib


class OracleError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"oracle error {code}: {msg}")
        self.code = code


def _chk(rc):
    if rc != 0:
        raise OracleError(rc, lib().oracle_last_error().decode())


def _p(a, ct):
    return None if a is None else a.ctypes.data_as(C.POINTER(ct))


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


def _i64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int64)


def num_threads():
    return lib().oracle_num_threads()


def rp_project(m, n, rm, dense=None, csc=None, cells=None, colsum=None, norm_mul=1e6, logkind=2, round_digits=-1):
    """dense: (m x n) array in Fortran (column-major) order, or csc = (colptr int64[n+1], rowidx int32, val f64).
    rm: dict with dgCMatrix slots ("p", "i", "x", "Dim")."""
    p = int(rm["Dim"][1])
    if dense is not None:
        e = np.asfortranarray(dense, dtype=np.float64)
        ev, er, ec = e, None, None
    else:
        ec, er, ev = _i64(csc[0]), _i32(csc[1]), _f64(csc[2])
    cells = _i64(cells)
    ncell = n if cells is None else len(cells)
    out = np.empty((ncell, p), dtype=np.float64)
    cs = _f64(colsum)
    rp, ri, rx = _i32(rm["p"]), _i32(rm["i"]), _f64(rm["x"])
    _chk(lib().oracle_rp_project(C.c_int(m), C.c_int(n), _p(ev, C.c_double), _p(er, C.c_int32), _p(ec, C.c_int64),
                                 _p(cells, C.c_int64), C.c_int64(ncell), _p(cs, C.c_double), C.c_double(norm_mul),
                                 C.c_int(logkind), C.c_int(round_digits), C.c_int(p), _p(rp, C.c_int32),
                                 _p(ri, C.c_int32), _p(rx, C.c_double), _p(out, C.c_double)))
    return out


def zscore_corrdist(mat):
    mat = _f64(mat)
    n, p = mat.shape
    z = np.empty((n, p))
    d = np.empty((n, n))
    _chk(lib().oracle_zscore_corrdist(n, p, _p(mat, C.c_double), _p(z, C.c_double), _p(d, C.c_double)))
    return z, d


def hclust(dist, method=WARD_D):
    dist = _f64(dist)
    n = dist.shape[0]
    ia = np.empty(n - 1, dtype=np.int32)
    ib = np.empty(n - 1, dtype=np.int32)
    crit = np.empty(n - 1)
    _chk(lib().oracle_hclust(n, _p(dist, C.c_double), method, _p(ia, C.c_int32), _p(ib, C.c_int32),
                             _p(crit, C.c_double)))
    return ia, ib, crit


def cutree_k(ia, ib, k):
    n = len(ia) + 1
    lab = np.empty(n, dtype=np.int32)
    _chk(lib().oracle_cutree_k(n, _p(_i32(ia), C.c_int32), _p(_i32(ib), C.c_int32), k, _p(lab, C.c_int32)))
    return lab


def silhouette_median(dist, labels, k):
    dist = _f64(dist)
    n = dist.shape[0]
    sil = np.empty(n)
    med = C.c_double()
    _chk(lib().oracle_silhouette_median(n, _p(dist, C.c_double), _p(_i32(labels), C.c_int32), k,
                                        _p(sil, C.c_double), C.byref(med)))
    return sil, med.value


def get_ch(y, labels, k):
    y = _f64(y)
    ch = C.c_double()
    _chk(lib().oracle_get_ch(y.shape[0], y.shape[1], _p(y, C.c_double), _p(_i32(labels), C.c_int32), k, C.byref(ch)))
    return ch.value


def opt_hclust(mat, symmetric=-1, prm=None):
    mat = _f64(mat)
    prm = prm or hc_params()
    nrow, ncol = mat.shape
    maxlev = max(1, prm.max_n - prm.min_n + 1)
    f = np.empty(nrow, dtype=np.int32)
    v = np.zeros(nrow * maxlev, dtype=np.int32)
    msil = np.zeros(maxlev)
    ch = np.zeros(maxlev)
    height = np.zeros(max(nrow - 1, 1))
    nlev, optn, oind = C.c_int(), C.c_int(), C.c_int()
    maxsil = C.c_double()
    _chk(lib().oracle_opt_hclust(nrow, ncol, _p(mat, C.c_double), symmetric, C.byref(prm), _p(f, C.c_int32),
                                 _p(v, C.c_int32), C.byref(nlev), _p(msil, C.c_double), _p(ch, C.c_double),
                                 _p(height, C.c_double), C.byref(optn), C.byref(maxsil), C.byref(oind)))
    L = nlev.value
    return {"f": f, "v": v[:nrow * L].reshape(nrow, L), "msil": msil[:L], "CHind": ch[:L], "height": height[:nrow - 1],
            "optN.cluster": optn.value, "maxsil": maxsil.value, "oind": oind.value}


def getrowcolor(emat, prm=None):
    emat = _f64(emat)
    prm = prm or hc_params()
    n, p = emat.shape
    color = np.empty(n, dtype=np.int32)
    ms = C.c_double()
    _chk(lib().oracle_getrowcolor(n, p, _p(emat, C.c_double), C.byref(prm), _p(color, C.c_int32), C.byref(ms)))
    return color, ms.value


def wmetac(labels, prm=None):
    """labels: N x C integer array."""
    lab = np.asfortranarray(labels, dtype=np.int32)
    N, Cc = lab.shape
    prm = prm or hc_params()
    maxc = max(prm.max_n, prm.n_cluster, 2) + 1
    fc = np.empty(N, dtype=np.int32)
    x0 = np.zeros(N * maxc)
    w1 = np.empty(N)
    nc = C.c_int()
    _chk(lib().oracle_wmetac(N, Cc, _p(lab, C.c_int32), C.byref(prm), _p(fc, C.c_int32), C.byref(nc),
                             _p(x0, C.c_double), maxc, _p(w1, C.c_double)))
    return {"finalC": fc, "x0": x0[:N * nc.value].reshape(N, nc.value), "w1": w1, "N.cluster": nc.value}


def smetac(labels, se1, prm=None):
    lab = _i32(labels)
    se1 = _f64(se1)
    ncells, p = se1.shape
    prm = prm or hc_params()
    fc = np.empty(ncells, dtype=np.int32)
    tf = np.empty(len(np.unique(lab)), dtype=np.int32)
    nc = C.c_int()
    _chk(lib().oracle_smetac(C.c_int64(ncells), p, _p(lab, C.c_int32), _p(se1, C.c_double), C.byref(prm),
                             _p(fc, C.c_int32), _p(tf, C.c_int32), C.byref(nc)))
    return {"finalColor": fc, "tf": tf}


def smetac_centroids(cen, ncells_total, prm=None):
    cen = _f64(cen)
    nC, p = cen.shape
    prm = prm or hc_params()
    tf = np.empty(nC, dtype=np.int32)
    _chk(lib().oracle_smetac_centroids(nC, p, _p(cen, C.c_double), C.c_int64(int(ncells_total)), C.byref(prm),
                                       _p(tf, C.c_int32)))
    return tf


def pack_rms(rms):
    """list of dgCMatrix dicts -> (colptr K x (p+1) int32, rowidx, x, nnz_off int64[K+1])"""
    colptr = np.ascontiguousarray(np.stack([r["p"] for r in rms]), dtype=np.int32)
    off = np.zeros(len(rms) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(r["i"]) for r in rms])
    ri = np.ascontiguousarray(np.concatenate([r["i"] for r in rms]), dtype=np.int32)
    rx = np.ascontiguousarray(np.concatenate([r["x"] for r in rms]), dtype=np.float64)
    return colptr, ri, rx, off


def sharp(m, n, rms, prm: SharpParams, dense=None, csc=None, colsum=None, norm_mul=1e6, reind=None,
          want_vie=True, want_x0=True, max_x0_cols=None):
    if dense is not None:
        e = np.asfortranarray(dense, dtype=np.float64)
        ev, er, ec = e, None, None
    else:
        ec, er, ev = _i64(csc[0]), _i32(csc[1]), _f64(csc[2])
    colptr, ri, rx, off = pack_rms(rms)
    p = prm.p
    pred = np.empty(n, dtype=np.int32)
    npred, x0c = C.c_int(), C.c_int()
    vie = np.empty((n, p)) if want_vie else None
    if max_x0_cols is None:
        max_x0_cols = max(64, prm.hc.max_n + 1)
    x0 = np.zeros(n * max_x0_cols) if want_x0 else None
    re = _i64(reind)
    cs = _f64(colsum)
    _chk(lib().oracle_sharp(C.c_int(m), C.c_int64(n), _p(ev, C.c_double), _p(er, C.c_int32), _p(ec, C.c_int64),
                            _p(cs, C.c_double), C.c_double(norm_mul), C.byref(prm), _p(colptr, C.c_int32),
                            _p(ri, C.c_int32), _p(rx, C.c_double), _p(off, C.c_int64), _p(re, C.c_int64),
                            _p(pred, C.c_int32), C.byref(npred), _p(vie, C.c_double), _p(x0, C.c_double),
                            C.byref(x0c), C.c_int(max_x0_cols)))
    res = {"pred_clusters": pred, "N.pred_cluster": npred.value, "viE": vie}
    if want_x0:
        res["x0"] = x0[:n * x0c.value].reshape(n, x0c.value)
    return res


def unlimited_combine(part_of, pred, e1, prm=None, n_cluster=0):
    e1 = _f64(e1)
    ncells, p = e1.shape
    prm = prm or hc_params()
    out = np.empty(ncells, dtype=np.int32)
    nf = C.c_int()
    _chk(lib().oracle_unlimited_combine(C.c_int64(ncells), p, _p(_i32(part_of), C.c_int32), _p(_i32(pred), C.c_int32),
                                        _p(e1, C.c_double), C.byref(prm), n_cluster, _p(out, C.c_int32), C.byref(nf)))
    return out, nf.value
