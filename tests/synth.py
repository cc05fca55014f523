"""Seeded synthetic scRNA-seq-like matrices for the tests (numpy; the bench has its own device generator)."""
import numpy as np


def make_expression(m, n, n_types=4, seed=0, zero_frac=0.7, kind="tpm", sep=1.5, frac=0.3):
    """genes x cells matrix (Fortran order) with planted cell types and realistic sparsity.

    Gene means are log-normal, a fraction `frac` of the genes is differentially expressed (log fold change
    ~ N(0, sep)) per type, counts are Poisson at a sequencing depth chosen so that `zero_frac` of the entries
    are zero (zeros concentrate in lowly expressed genes, like drop-outs do).
    kind "umi": the raw counts; kind "tpm": columns scaled to 1e6."""
    rng = np.random.default_rng(seed)
    base = rng.normal(0.0, 1.5, size=m)
    types = base[:, None] + sep * rng.normal(0.0, 1.0, size=(m, n_types)) * (rng.random((m, n_types)) < frac)
    lab = rng.integers(0, n_types, size=n)
    mu = np.exp(types)
    lo, hi = 1e-4, 1e3
    for _ in range(60):
        mid = np.sqrt(lo * hi)
        if np.exp(-mid * mu).mean() > zero_frac:
            lo = mid
        else:
            hi = mid
    depth = np.sqrt(lo * hi)
    x = rng.poisson(depth * mu[:, lab] * np.exp(0.2 * rng.normal(size=(1, n)))).astype(np.float64)
    # every cell needs a few non-zero genes (a constant projection makes scale() produce NaN in the reference)
    for c in np.flatnonzero((x != 0).sum(0) < 3):
        x[rng.integers(0, m, 5), c] += 1.0
    if kind != "umi":
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
    sij = int(comb(ct).sum())                # Python integers: sa * sb exceeds 2^63 beyond ~1e5 cells
    sa = int(comb(ct.sum(1)).sum())
    sb = int(comb(ct.sum(0)).sum())
    tot = int(comb(len(a)))
    exp = sa * sb / tot
    mx = (sa + sb) / 2
    if mx == exp:
        return 1.0
    return float((sij - exp) / (mx - exp))
