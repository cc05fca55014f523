"""ctypes binding of libsharpb200.so (the C ABI declared in include/sharp_b200.h).

This is the only door to the compute path: there is no CPU fallback.  Loading fails loudly when the
shared library has not been built (``python -c "import __graft_entry__ as g; g.build()"``), and every
compute call fails with :class:`SharpError` (``SHARP_E_CUDA``) on a machine without a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libsharpb200.so")

WARD_D, SINGLE, COMPLETE, AVERAGE, MCQUITTY, MEDIAN, CENTROID, WARD_D2 = 1, 2, 3, 4, 5, 6, 7, 8
HMETHODS = {"ward.D": WARD_D, "ward": WARD_D, "single": SINGLE, "complete": COMPLETE, "average": AVERAGE,
            "mcquitty": MCQUITTY, "median": MEDIAN, "centroid": CENTROID, "ward.D2": WARD_D2}

E_ARG, E_CUDA, E_NOMEM, E_RSTOP, E_LIMIT = -1, -2, -3, -4, -5


class SharpError(RuntimeError):
    """An error reported by libsharpb200 (``code`` is one of the SHARP_E_* values)."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"[sharp_b200 {code}] {msg}")
        self.code = code
        self.msg = msg


class RStop(SharpError):
    """The reference itself would ``stop()`` here; the message quotes the R error."""


class HcParams(C.Structure):
    _fields_ = [("hmethod", C.c_int), ("n_cluster", C.c_int), ("min_n", C.c_int), ("max_n", C.c_int),
                ("sil_thre", C.c_double), ("height_ntimes", C.c_double)]


class RunParams(C.Structure):
    _fields_ = [("large", C.c_int), ("logflag", C.c_int), ("logkind", C.c_int), ("round_digits", C.c_int),
                ("partition_ncells", C.c_int), ("n_cluster", C.c_int), ("enp_n_cluster", C.c_int),
                ("ind_n_cluster", C.c_int), ("hc", HcParams), ("normalize", C.c_int), ("norm_mul", C.c_double),
                ("skip_smetac", C.c_int), ("block_max_n", C.c_int), ("shard", C.c_int), ("shard_rotate", C.c_int)]


class Part(C.Structure):
    """sharp_part: one part of SHARP_unlimited for sharp_run_parts"""
    _fields_ = [("n", C.c_int64), ("dev", C.c_void_p), ("dense", C.POINTER(C.c_double)), ("colptr", C.POINTER(C.c_int64)),
                ("rowidx", C.POINTER(C.c_int32)), ("val", C.POINTER(C.c_double)), ("reind", C.POINTER(C.c_int64)),
                ("pred", C.POINTER(C.c_int32)), ("nclust", C.c_int), ("cen", C.POINTER(C.c_double)),
                ("counts", C.POINTER(C.c_int64)), ("sharded", C.c_int)]


# every symbol include/sharp_b200.h declares (tests check that the library exports all of them)
EXPORTS = [
    "sharp_abi_version", "sharp_last_error", "sharp_device_count", "sharp_device_info", "sharp_ctx_create",
    "sharp_ctx_destroy", "sharp_ctx_stream", "sharp_ctx_sync", "sharp_timer_start", "sharp_timer_stop_ms",
    "sharp_ctx_launch_count", "sharp_rm_upload", "sharp_rm_free", "sharp_rp_project", "sharp_corrdist",
    "sharp_hclust", "sharp_opt_hclust", "sharp_getrowcolor", "sharp_wmetac", "sharp_smetac", "sharp_run",
    "sharp_expr_upload", "sharp_expr_free", "sharp_run_dev", "sharp_centroids", "sharp_smetac_centroids",
    "sharp_last_member", "sharp_last_vie", "sharp_prof_enable", "sharp_prof_reset", "sharp_prof_kernels", "sharp_prof_name",
    "sharp_prof_get", "sharp_ctx_set_rp_variant", "sharp_r_ranm", "sharp_r_sample_perm", "sharp_labels_combine", "sharp_first_appearance_codes", "sharp_ctx_bind_host", "sharp_run_parts",
    "sharp_ctx_set_block_budget", "sharp_ctx_set_serial", "sharp_parts_prefetch", "sharp_plan_groups",
    "sharp_comm_unique_id", "sharp_comm_init", "sharp_comm_destroy", "sharp_comm_info", "sharp_comm_allgatherv",
    "sharp_comm_bcast", "sharp_comm_barrier", "sharp_host_alloc", "sharp_host_free", "sharp_csc_file_info",
    "sharp_csc_file_read",
]

_lib = None


def load():
    """Load libsharpb200.so; raises ImportError with build instructions when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(
            f"{SO_PATH} is missing: the CUDA extension has not been built. Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc); sharp_b200 has no CPU fallback.")
    lib = C.CDLL(SO_PATH)
    lib.sharp_last_error.restype = C.c_char_p
    lib.sharp_ctx_stream.restype = C.c_void_p
    lib.sharp_prof_name.restype = C.c_char_p
    lib.sharp_ctx_launch_count.restype = C.c_int64
    lib.sharp_ctx_destroy.restype = None
    lib.sharp_rm_free.restype = None
    lib.sharp_expr_free.restype = None
    for name in ("sharp_ctx_destroy", "sharp_ctx_stream", "sharp_ctx_sync", "sharp_timer_start",
                 "sharp_ctx_launch_count", "sharp_rm_free", "sharp_expr_free"):
        getattr(lib, name).argtypes = [C.c_void_p]
    lib.sharp_timer_stop_ms.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    lib.sharp_host_free.restype = None
    lib.sharp_host_free.argtypes = [C.c_void_p]
    _lib = lib
    return lib


def _check(rc: int):
    if rc != 0:
        msg = load().sharp_last_error().decode(errors="replace")
        raise (RStop if rc == E_RSTOP else SharpError)(rc, msg)


def _ptr(a, ct):
    return None if a is None else a.ctypes.data_as(C.POINTER(ct))


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


def _i64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int64)


def hc_params(hmethod="ward.D", n_cluster=None, min_n=2, max_n=40, sil_thre=0.35, height_ntimes=2.0) -> HcParams:
    if isinstance(hmethod, str):
        if hmethod not in HMETHODS:
            raise ValueError("invalid clustering method " + repr(hmethod))
        hmethod = HMETHODS[hmethod]
    return HcParams(int(hmethod), int(n_cluster or 0), int(min_n), int(max_n), float(sil_thre), float(height_ntimes))


def device_count() -> int:
    return int(load().sharp_device_count())


def device_info(device=0) -> dict:
    name = C.create_string_buffer(256)
    sm, maj, mnr = C.c_int(), C.c_int(), C.c_int()
    mem = C.c_size_t()
    _check(load().sharp_device_info(int(device), name, 256, C.byref(sm), C.byref(maj), C.byref(mnr), C.byref(mem)))
    return {"name": name.value.decode(), "sm_count": sm.value, "cc": (maj.value, mnr.value), "total_mem": mem.value}


def _expr_args(m, n, dense, csc):
    """-> (keepalive, dense_ptr, colptr_ptr, rowidx_ptr, val_ptr)"""
    if dense is not None:
        e = np.asfortranarray(dense, dtype=np.float64)
        if e.shape != (m, n):
            raise ValueError("dense expression matrix must be genes x cells")
        return (e,), _ptr(e, C.c_double), None, None, None
    cp, ri, v = _i64(csc[0]), _i32(csc[1]), _f64(csc[2])
    if cp.shape[0] != n + 1:
        raise ValueError("CSC colptr must have ncells + 1 entries")
    return (cp, ri, v), None, _ptr(cp, C.c_int64), _ptr(ri, C.c_int32), _ptr(v, C.c_double)


def r_ranm(m: int, p: int, seed: int) -> dict:
    """ranM2(m, p, seed) through the native generator (sharp_r_ranm) -> dgCMatrix slots"""
    lib = load()
    cap = int(p * (m ** 0.5) * 1.15) + 4096
    colptr = np.empty(p + 1, dtype=np.int32)
    while True:
        ri = np.empty(cap, dtype=np.int32)
        x = np.empty(cap, dtype=np.float64)
        nnz = C.c_int64()
        rc = lib.sharp_r_ranm(int(m), int(p), C.c_int64(int(seed)), _ptr(colptr, C.c_int32), _ptr(ri, C.c_int32),
                              _ptr(x, C.c_double), C.c_int64(cap), C.byref(nnz))
        if rc == E_NOMEM:
            cap = nnz.value
            continue
        if rc != 0:
            raise SharpError(rc, "sharp_r_ranm failed")
        return {"Dim": (int(m), int(p)), "p": colptr, "i": ri[:nnz.value], "x": x[:nnz.value]}


def r_sample_perm_native(n: int, seed: int) -> np.ndarray:
    """set.seed(seed); sample(n) through the native generator (sharp_r_sample_perm)"""
    out = np.empty(int(n), dtype=np.int64)
    rc = load().sharp_r_sample_perm(C.c_int64(int(seed)), C.c_int64(int(n)), _ptr(out, C.c_int64))
    if rc != 0:
        raise SharpError(rc, "sharp_r_sample_perm failed")
    return out


def first_appearance_codes(y: np.ndarray, nvals: int) -> tuple[np.ndarray, np.ndarray]:
    """match(y, unique(y)) for int32 ids in 0..nvals-1 (sharp_first_appearance_codes, pure host) -> (codes, unique(y))"""
    y = np.ascontiguousarray(y, dtype=np.int32)
    codes = np.empty(len(y), dtype=np.int32)
    uniq = np.empty(int(nvals), dtype=np.int32)
    k = load().sharp_first_appearance_codes(_ptr(y, C.c_int32), C.c_int64(len(y)), int(nvals), _ptr(codes, C.c_int32),
                                            _ptr(uniq, C.c_int32))
    if k < 0:
        raise SharpError(E_ARG, "sharp_first_appearance_codes: id out of range")
    return codes, uniq[:k]


def labels_combine(preds: list, counts_per_part: list, tf: np.ndarray, merge_thre: int) -> tuple[np.ndarray, np.ndarray]:
    """sharp_labels_combine: (final labels 1.. by decreasing size, their sizes) from the per-part cluster ids and the
    global sMetaC's tf (R/SHARP_unlimited.R:166-183) -- pure host code"""
    nparts = len(preds)
    start = np.zeros(nparts + 1, dtype=np.int64)
    start[1:] = np.cumsum([len(p) for p in preds])
    off = np.zeros(nparts, dtype=np.int32)
    off[1:] = np.cumsum(counts_per_part[:-1])
    pred = np.ascontiguousarray(np.concatenate(preds), dtype=np.int32)
    tf = np.ascontiguousarray(tf, dtype=np.int32)
    out = np.empty(int(start[-1]), dtype=np.int32)
    counts = np.zeros(len(tf), dtype=np.int64)
    nl = C.c_int()
    rc = load().sharp_labels_combine(nparts, _ptr(start, C.c_int64), _ptr(off, C.c_int32), _ptr(pred, C.c_int32),
                                     _ptr(tf, C.c_int32), len(tf), int(merge_thre), _ptr(out, C.c_int32),
                                     _ptr(counts, C.c_int64), C.byref(nl))
    if rc != 0:
        raise SharpError(rc, "sharp_labels_combine: inconsistent cluster ids")
    return out, counts[:nl.value]


def plan_groups(nparts: int, host_data: bool, group=0, lanes=0) -> dict:
    """sharp_plan_groups: how the loop over parts splits ``nparts`` parts -> {"gstart", "group", "lanes"} (host logic)"""
    g = np.zeros(nparts + 2, dtype=np.int32)
    ng, gu, lu = C.c_int(), C.c_int(), C.c_int()
    _check(load().sharp_plan_groups(int(nparts), int(bool(host_data)), int(group), int(lanes), _ptr(g, C.c_int), len(g),
                                    C.byref(ng), C.byref(gu), C.byref(lu)))
    return {"gstart": g[:ng.value + 1].tolist(), "group": gu.value, "lanes": lu.value}


class RmDev:
    """K ranM matrices prepared on the device (sharp_rm_upload)."""

    def __init__(self, ctx: "Context", rms: list):
        self.K = len(rms)
        self.m, self.p = int(rms[0]["Dim"][0]), int(rms[0]["Dim"][1])
        colptr = np.ascontiguousarray(np.stack([np.asarray(r["p"], dtype=np.int32) for r in rms]))
        off = np.zeros(self.K + 1, dtype=np.int64)
        off[1:] = np.cumsum([len(r["i"]) for r in rms])
        ri = _i32(np.concatenate([r["i"] for r in rms]))
        rx = _f64(np.concatenate([r["x"] for r in rms]))
        self._h = C.c_void_p()
        self._ctx = ctx
        _check(load().sharp_rm_upload(ctx._h, self.m, self.p, self.K, _ptr(colptr, C.c_int32), _ptr(ri, C.c_int32),
                                      _ptr(rx, C.c_double), _ptr(off, C.c_int64), C.byref(self._h)))

    def close(self):
        if self._h:
            load().sharp_rm_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ExprDev:
    """An expression matrix (or one part) resident on the device (sharp_expr_upload)."""

    def __init__(self, ctx: "Context", m, n, dense=None, csc=None):
        self.m, self.n = int(m), int(n)
        self._keep, dp, cp, ri, v = _expr_args(self.m, self.n, dense, csc)
        self._h = C.c_void_p()
        _check(load().sharp_expr_upload(ctx._h, self.m, C.c_int64(self.n), dp, cp, ri, v, C.byref(self._h)))

    def close(self):
        if self._h:
            load().sharp_expr_free(self._h)
            self._h = C.c_void_p()
        self._keep = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Context:
    """One CUDA device + stream + workspace (sharp_ctx)."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        _check(load().sharp_ctx_create(int(device), C.byref(self._h)))
        self.device = int(device)

    def close(self):
        if self._h:
            load().sharp_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- plumbing ------------------------------------------------------------------------------
    def sync(self):
        _check(load().sharp_ctx_sync(self._h))

    def stream(self) -> int:
        return int(load().sharp_ctx_stream(self._h) or 0)

    def timer_start(self):
        _check(load().sharp_timer_start(self._h))

    def timer_stop_ms(self) -> float:
        ms = C.c_double()
        _check(load().sharp_timer_stop_ms(self._h, C.byref(ms)))
        return ms.value

    def launch_count(self) -> int:
        return int(load().sharp_ctx_launch_count(self._h))

    def set_rp_variant(self, variant):
        """projection kernel: 0 / False (default) record-gather fixed point (CSC input; dense input takes 2); 1 / True the
        fp64 gene-order read-modify-write kernel; 2 the round-1 fixed-point kernel; 3 record-gather with the cell's
        (rowidx, val) segments staged in shared memory by cp.async.bulk"""
        _check(load().sharp_ctx_set_rp_variant(self._h, int(variant)))

    def prof_enable(self, on=True):
        _check(load().sharp_prof_enable(self._h, int(bool(on))))

    def prof_reset(self):
        _check(load().sharp_prof_reset(self._h))

    def prof_get(self) -> dict:
        """{kernel class: (total device ms, launches)} accumulated while profiling was enabled"""
        lib = load()
        out = {}
        for kid in range(lib.sharp_prof_kernels()):
            ms, n = C.c_double(), C.c_int64()
            _check(lib.sharp_prof_get(self._h, kid, C.byref(ms), C.byref(n)))
            if n.value:
                out[lib.sharp_prof_name(kid).decode()] = (ms.value, n.value)
        return out

    def upload_rm(self, rms: list) -> RmDev:
        return RmDev(self, rms)

    def upload_expr(self, m, n, dense=None, csc=None) -> ExprDev:
        return ExprDev(self, m, n, dense, csc)

    # -- stages ----------------------------------------------------------------------------------
    def rp_project(self, m, n, rm: RmDev, dense=None, csc=None, cells=None, normalize=0, colsum=None, norm_mul=1e6,
                   logkind=2, round_digits=-1) -> np.ndarray:
        """-> array (K, ncell, p)"""
        keep, dp, cp, ri, v = _expr_args(m, n, dense, csc)
        cells = _i64(cells)
        ncell = n if cells is None else len(cells)
        out = np.empty((rm.K, ncell, rm.p), dtype=np.float64)
        cs = _f64(colsum)
        _check(load().sharp_rp_project(self._h, int(m), C.c_int64(n), dp, cp, ri, v, _ptr(cells, C.c_int64),
                                       C.c_int64(ncell), int(normalize), _ptr(cs, C.c_double), C.c_double(norm_mul),
                                       int(logkind), int(round_digits), rm._h, _ptr(out, C.c_double)))
        return out

    def corrdist(self, mat) -> np.ndarray:
        mat = _f64(mat)
        n, p = mat.shape
        d = np.empty((n, n))
        _check(load().sharp_corrdist(self._h, n, p, _ptr(mat, C.c_double), _ptr(d, C.c_double)))
        return d

    def hclust(self, dist, method=WARD_D):
        dist = _f64(dist)
        n = dist.shape[0]
        if isinstance(method, str):
            method = HMETHODS[method]
        ia = np.empty(max(n - 1, 1), dtype=np.int32)
        ib = np.empty(max(n - 1, 1), dtype=np.int32)
        h = np.empty(max(n - 1, 1))
        _check(load().sharp_hclust(self._h, n, _ptr(dist, C.c_double), int(method), _ptr(ia, C.c_int32),
                                   _ptr(ib, C.c_int32), _ptr(h, C.c_double)))
        return ia[:n - 1], ib[:n - 1], h[:n - 1]

    def opt_hclust(self, mat, symmetric: bool, prm: HcParams, exact=False, want_v=True) -> dict:
        mat = _f64(mat)
        nrow, ncol = mat.shape
        maxlev = 1 if prm.n_cluster else max(1, prm.max_n - prm.min_n + 1)
        f = np.empty(nrow, dtype=np.int32)
        v = np.zeros(nrow * maxlev, dtype=np.int32) if want_v else None
        msil = np.zeros(maxlev)
        ch = np.zeros(maxlev)
        height = np.zeros(max(nrow - 1, 1))
        nlev, optn, oind = C.c_int(), C.c_int(), C.c_int()
        maxsil = C.c_double()
        _check(load().sharp_opt_hclust(self._h, nrow, ncol, _ptr(mat, C.c_double), int(bool(symmetric)), int(bool(exact)),
                                       C.byref(prm), _ptr(f, C.c_int32), _ptr(v, C.c_int32), C.byref(nlev),
                                       _ptr(msil, C.c_double), _ptr(ch, C.c_double), _ptr(height, C.c_double),
                                       C.byref(optn), C.byref(maxsil), C.byref(oind)))
        L = nlev.value
        return {"f": f, "v": None if v is None else v[:nrow * L].reshape(nrow, L), "msil": msil[:L], "CHind": ch[:L],
                "height": height[:nrow - 1], "optN.cluster": optn.value, "maxsil": maxsil.value, "oind": oind.value}

    def getrowcolor(self, emat, prm: HcParams):
        emat = _f64(emat)
        n, p = emat.shape
        color = np.empty(n, dtype=np.int32)
        ms = C.c_double()
        _check(load().sharp_getrowcolor(self._h, n, p, _ptr(emat, C.c_double), C.byref(prm), _ptr(color, C.c_int32),
                                        C.byref(ms)))
        return color, ms.value

    def wmetac(self, labels, prm: HcParams, want_x0=True) -> dict:
        lab = np.asfortranarray(labels, dtype=np.int32)
        N, Cc = lab.shape
        maxc = max(prm.max_n, prm.n_cluster, 2) + 1
        fc = np.empty(N, dtype=np.int32)
        x0 = np.zeros(N * maxc) if want_x0 else None
        w1 = np.empty(N)
        nc = C.c_int()
        _check(load().sharp_wmetac(self._h, N, Cc, _ptr(lab, C.c_int32), C.byref(prm), _ptr(fc, C.c_int32), C.byref(nc),
                                   _ptr(x0, C.c_double), maxc, _ptr(w1, C.c_double)))
        return {"finalC": fc, "x0": None if x0 is None else x0[:N * nc.value].reshape(N, nc.value), "w1": w1,
                "N.cluster": nc.value}

    def smetac(self, labels, se1, prm: HcParams) -> dict:
        lab = _i32(labels)
        se1 = _f64(se1)
        ncells, p = se1.shape
        fc = np.empty(ncells, dtype=np.int32)
        tf = np.empty(len(np.unique(lab)), dtype=np.int32)
        nc = C.c_int()
        _check(load().sharp_smetac(self._h, C.c_int64(ncells), p, _ptr(lab, C.c_int32), _ptr(se1, C.c_double),
                                   C.byref(prm), _ptr(fc, C.c_int32), _ptr(tf, C.c_int32), C.byref(nc)))
        return {"finalColor": fc, "tf": tf[:nc.value]}

    def bind_host(self) -> str:
        """bind this thread to the CPUs next to the GPU (sharp_ctx_bind_host); returns the cpulist applied ("" = no-op)"""
        buf = C.create_string_buffer(4096)
        _check(load().sharp_ctx_bind_host(self._h, buf, 4096))
        return buf.value.decode()

    def smetac_centroids(self, cen, ncells_total, prm: HcParams) -> np.ndarray:
        cen = _f64(cen)
        nC, p = cen.shape
        tf = np.empty(nC, dtype=np.int32)
        _check(load().sharp_smetac_centroids(self._h, nC, p, _ptr(cen, C.c_double), C.c_int64(ncells_total), C.byref(prm),
                                             _ptr(tf, C.c_int32)))
        return tf

    # -- fused pipeline ----------------------------------------------------------------------------
    def run(self, rm: RmDev, prm: RunParams, m=None, n=None, dense=None, csc=None, expr: ExprDev | None = None,
            colsum=None, reind=None, want_vie=True, want_x0=True, max_x0_cols=None) -> dict:
        """SHARP_small / SHARP_large compute for one matrix (host buffers) or one device-resident part."""
        if expr is not None:
            m, n = expr.m, expr.n
        labels = np.empty(n, dtype=np.int32)
        vie = np.empty((n, rm.p)) if want_vie else None
        if max_x0_cols is None:
            max_x0_cols = max(64, prm.hc.max_n + 1, prm.n_cluster + 1)
        x0 = np.zeros(n * max_x0_cols) if want_x0 else None
        x0c = C.c_int()
        cs = _f64(colsum)
        re = _i64(reind)
        if expr is not None:
            _check(load().sharp_run_dev(self._h, expr._h, _ptr(cs, C.c_double), rm._h, _ptr(re, C.c_int64), C.byref(prm),
                                        _ptr(labels, C.c_int32), _ptr(vie, C.c_double), _ptr(x0, C.c_double),
                                        C.byref(x0c), int(max_x0_cols)))
        else:
            keep, dp, cp, ri, v = _expr_args(m, n, dense, csc)
            _check(load().sharp_run(self._h, int(m), C.c_int64(n), dp, cp, ri, v, _ptr(cs, C.c_double), rm._h,
                                    _ptr(re, C.c_int64), C.byref(prm), _ptr(labels, C.c_int32), _ptr(vie, C.c_double),
                                    _ptr(x0, C.c_double), C.byref(x0c), int(max_x0_cols)))
        res = {"labels": labels, "viE": vie, "x0_cols": x0c.value}
        if want_x0:
            res["x0"] = x0[:n * x0c.value].reshape(n, x0c.value)
        return res

    def set_block_budget(self, gigabytes: int):
        _check(load().sharp_ctx_set_block_budget(self._h, int(gigabytes)))

    def set_serial(self, on: bool):
        """sharp_run_parts on ONE stream (isolated per-kernel timings for the roofline object); default off"""
        _check(load().sharp_ctx_set_serial(self._h, int(bool(on))))

    def parts_prefetch(self, m: int, parts: list, group=0, lanes=0, sharded=None):
        """sharp_parts_prefetch: start the uploads of the first group of ``parts`` (same list and group / lanes as the
        run_parts call that follows) and return at once.  -> an object the caller keeps alive until run_parts is done."""
        nparts = len(parts)
        arr = (Part * nparts)()
        keep = [arr]
        for i, pt in enumerate(parts):
            a = arr[i]
            if isinstance(pt, ExprDev):
                a.n = pt.n
                a.dev = pt._h
            else:
                a.n = int(pt["n"])
                a.dev = None
                k, dp, cp, ri, v = _expr_args(m, a.n, pt.get("dense"), pt.get("csc"))
                keep.append(k)
                a.dense, a.colptr, a.rowidx, a.val = dp, cp, ri, v
            a.sharded = int(bool(sharded[i])) if sharded is not None else 0
        _check(load().sharp_parts_prefetch(self._h, int(m), nparts, arr, int(group), int(lanes)))
        return keep

    def run_parts(self, rm: RmDev, prm: RunParams, m: int, parts: list, reinds: list, small_thre=10, cen_cap=64,
                  group=0, lanes=0, sharded=None) -> list:
        """The per-part loop of SHARP_unlimited in one call (sharp_run_parts).  ``parts``: list of ExprDev (device
        resident) or dicts {"n", "dense"} / {"n", "csc": (p, i, x)} (host).  -> per part {"pred_clusters",
        "N.pred_cluster", "cen" (nclust x p), "counts"}."""
        nparts = len(parts)
        arr = (Part * nparts)()
        keep = []
        outs = []
        for i, (pt, re) in enumerate(zip(parts, reinds)):
            a = arr[i]
            if isinstance(pt, ExprDev):
                a.n = pt.n
                a.dev = pt._h
            else:
                a.n = int(pt["n"])
                a.dev = None
                k, dp, cp, ri, v = _expr_args(m, a.n, pt.get("dense"), pt.get("csc"))
                keep.append(k)
                a.dense, a.colptr, a.rowidx, a.val = dp, cp, ri, v
            re = _i64(re)
            keep.append(re)
            a.reind = _ptr(re, C.c_int64)
            a.sharded = int(bool(sharded[i])) if sharded is not None else 0
            pred = np.empty(a.n, dtype=np.int32)
            cen = np.empty((cen_cap, rm.p))
            cnt = np.zeros(cen_cap, dtype=np.int64)
            a.pred, a.cen, a.counts = _ptr(pred, C.c_int32), _ptr(cen, C.c_double), _ptr(cnt, C.c_int64)
            outs.append((pred, cen, cnt))
        _check(load().sharp_run_parts(self._h, int(m), nparts, arr, rm._h, C.byref(prm), int(small_thre), int(cen_cap),
                                      int(group), int(lanes)))
        res = []
        for i, (pred, cen, cnt) in enumerate(outs):
            nc = int(arr[i].nclust)
            res.append({"pred_clusters": pred, "N.pred_cluster": nc, "cen": cen[:nc].copy(), "counts": cnt[:nc].copy()})
        return res

    def centroids(self, labels, nclust, p):
        """colMeans of the last run's viE per cluster id 1..nclust -> (nclust x p means, counts)."""
        lab = _i32(labels)
        cen = np.empty((nclust, p))
        cnt = np.empty(nclust, dtype=np.int64)
        _check(load().sharp_centroids(self._h, C.c_int64(len(lab)), _ptr(lab, C.c_int32), int(nclust),
                                      _ptr(cen, C.c_double), _ptr(cnt, C.c_int64)))
        return cen, cnt

    def last_member(self, k, n, p):
        """member k of the last run -> (colour index per cell, projection n x p)   (SHARP_small's allrpinfo)"""
        col = np.empty(n, dtype=np.int32)
        inde = np.empty((n, p))
        _check(load().sharp_last_member(self._h, int(k), C.c_int64(n), _ptr(col, C.c_int32), _ptr(inde, C.c_double)))
        return col, inde

    def last_vie(self, n, p):
        """viE = enE/K of the last run (n x p, un-shuffled)"""
        vie = np.empty((n, p))
        _check(load().sharp_last_vie(self._h, C.c_int64(n), int(p), _ptr(vie, C.c_double)))
        return vie
