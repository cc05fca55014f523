"""Seeded synthetic scRNA-seq-like matrices for the tests (numpy; the bench has its own device generator)."""
import numpy as np


def make_expression(m, n, n_types=4, seed=0, zero_frac=0.7, kind="tpm", sep=1.0):
    """genes x cells matrix (Fortran order) with planted cell types.

    kind "tpm": columns scaled to 1e6 (TPM-like, dense with ~zero_frac zeros);
    kind "umi": small integer counts."""
    rng = np.random.default_rng(seed)
    base = rng.normal(0.0, 1.0, size=m)
    types = base[:, None] + sep * rng.normal(0.0, 1.0, size=(m, n_types)) * (rng.random((m, n_types)) < 0.15)
    lab = rng.integers(0, n_types, size=n)
    mu = np.exp(types[:, lab] + 0.3 * rng.normal(size=(m, n)))
    keep = rng.random((m, n)) > zero_frac
    x = mu * keep
    if kind == "umi":
        x = rng.poisson(np.minimum(x * 0.6, 50.0)).astype(np.float64)
        # every cell needs a few non-zero genes
        for c in np.flatnonzero(x.sum(0) == 0):
            x[rng.integers(0, m, 5), c] = 1.0
    else:
        x = x / x.sum(0, keepdims=True) * 1e6
    return np.asfortranarray(x), lab + 1


def to_csc(x):
    """dense genes x cells -> (colptr int64[n+1], rowidx int32[nnz], val f64[nnz]) (dgCMatrix slots p, i, x)."""
    m, n = x.shape
    mask = (x != 0)
    counts = mask.sum(0)
    colptr = np.zeros(n + 1, dtype=np.int64)
    colptr[1:] = np.cumsum(counts)
    cols, rows = np.nonzero(mask.T)
    return colptr, rows.astype(np.int32), np.ascontiguousarray(x.T[mask.T], dtype=np.float64)


def ari(a, b):
    """Hubert-Arabie adjusted Rand index (the 'HA' entry of clues::adjustedRand, R/ARI.R:38)."""
    a = np.asarray(a)
    b = np.asarray(b)
    _, ai = np.unique(a, return_inverse=True)
    _, bi = np.unique(b, return_inverse=True)
    ct = np.zeros((ai.max() + 1, bi.max() + 1), dtype=np.int64)
    np.add.at(ct, (ai, bi), 1)
    comb = lambda x: x * (x - 1) // 2
    sij = comb(ct).sum()
    sa = comb(ct.sum(1)).sum()
    sb = comb(ct.sum(0)).sum()
    tot = comb(len(a))
    exp = sa * sb / tot
    mx = (sa + sb) / 2
    if mx == exp:
        return 1.0
    return float((sij - exp) / (mx - exp))
