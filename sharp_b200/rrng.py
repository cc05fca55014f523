"""R-compatible random numbers for the host side (no R interpreter in this image).

The reference generates its projection matrices and its cell shuffle with base R:
``set.seed(seedn); sample(c(sqrt(s), 0, -sqrt(s)), m*p, replace=TRUE, prob=...)`` (R/ranM.R:17-30,
R/ranM2.R:17-32, R/RPmat.R:20-30) and ``set.seed(50); sample(ncells)`` (R/SHARP.R:497-498).  Where R
exists those calls stay in R and the results are handed to the C ABI; here the same streams are
reproduced so that a seeded run is identical with or without R:

* ``set.seed``: Mersenne-Twister seeded by R's initial scrambling (50 + 625 steps of the LCG
  ``69069*seed+1``), ``mti = 624``;
* ``unif_rand``: ``genrand_int32 * 2.3283064365386963e-10`` clamped into (0, 1);
* ``sample(x, size, TRUE, prob)`` with < 200 categories: ``ProbSampleReplace`` (``revsort`` + cumulative
  sums, one uniform per draw);
* ``sample(n)`` (R >= 3.6 "Rejection" sampling): ``R_unif_index`` built from 16-bit chunks.

numpy's MT19937 bit generator is the same generator, so the heavy lifting is vectorised.
"""
from __future__ import annotations

import math

import numpy as np

_I2_32M1 = 2.328306437080797e-10  # R's i2_32m1, used by fixup()


class RRandom:
    """State of R's default RNG (Mersenne-Twister, Inversion, Rejection) after ``set.seed(seed)``."""

    def __init__(self, seed: int):
        self.set_seed(seed)

    def set_seed(self, seed: int) -> None:
        s = np.uint32(int(seed) & 0xFFFFFFFF)
        a = np.uint32(69069)
        one = np.uint32(1)
        with np.errstate(over="ignore"):
            for _ in range(50):  # Randomize(): initial scrambling
                s = a * s + one
            dummy = np.empty(625, dtype=np.uint32)
            for j in range(625):  # RNG_Init(): i_seed[j]
                s = a * s + one
                dummy[j] = s
        # FixupSeeds: dummy[0] = mti = 624 (N): the first draw regenerates the whole table
        key = dummy[1:].copy()
        self._bg = np.random.MT19937()
        self._bg.state = {"bit_generator": "MT19937", "state": {"key": key, "pos": 624}}

    def unif_rand(self, size: int) -> np.ndarray:
        """``size`` successive ``unif_rand()`` values (fp64)."""
        raw = self._bg.random_raw(int(size)).astype(np.float64)
        u = raw * 2.3283064365386963e-10
        # fixup(): keep strictly inside (0, 1)
        u = np.where(u <= 0.0, 0.5 * _I2_32M1, u)
        u = np.where(1.0 - u <= 0.0, 1.0 - 0.5 * _I2_32M1, u)
        return u

    # -- sample(n): permutation by rejection sampling -------------------------------------------
    def sample_perm(self, n: int) -> np.ndarray:
        """``sample(n)``: 1-based permutation of 1..n (R >= 3.6, sample.kind = "Rejection")."""
        n = int(n)
        x = np.arange(n, dtype=np.int64)
        y = np.empty(n, dtype=np.int64)
        # rbits() consumes (bits // 16 + 1) uniforms per attempt.  bits changes as dn shrinks, so draw
        # lazily in chunks; the stream is consumed strictly in order.
        buf = np.empty(0)
        pos = 0
        remaining = n
        for i in range(n):
            dn = remaining
            if dn <= 1:  # bits = ceil(log2(dn)) = 0 -> rbits(0) still consumes one uniform, result 0
                bits = 0
            else:
                bits = int(math.ceil(math.log2(dn)))
            nchunk = bits // 16 + 1
            while True:
                if pos + nchunk > buf.shape[0]:
                    need = max(4096, 4 * nchunk * (n - i))
                    buf = np.concatenate([buf[pos:], self.unif_rand(min(need, 1 << 22))])
                    pos = 0
                v = 0
                for q in range(nchunk):
                    v = 65536 * v + int(math.floor(buf[pos + q] * 65536))
                pos += nchunk
                if bits < 64:
                    v &= (1 << bits) - 1
                if v < dn:
                    break
            j = v
            y[i] = x[j] + 1
            remaining -= 1
            x[j] = x[remaining]
        # give unread uniforms back is impossible with MT; callers only use one sample() per set.seed()
        self._leftover = buf[pos:]
        return y


def _revsort_perm(p: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """R's ``revsort(a, ib, n)``: heap sort into DESCENDING order carrying 1-based ids (not stable)."""
    a = [0.0] + [float(v) for v in p]  # 1-based
    ib = [0] + list(range(1, len(p) + 1))
    n = len(p)
    if n <= 1:
        return np.array(a[1:]), np.array(ib[1:])
    l = (n >> 1) + 1
    ir = n
    while True:
        if l > 1:
            l -= 1
            ra, ii = a[l], ib[l]
        else:
            ra, ii = a[ir], ib[ir]
            a[ir], ib[ir] = a[1], ib[1]
            ir -= 1
            if ir == 1:
                a[1], ib[1] = ra, ii
                break
        i = l
        j = l << 1
        while j <= ir:
            if j < ir and a[j] > a[j + 1]:
                j += 1
            if ra > a[j]:
                a[i], ib[i] = a[j], ib[j]
                i = j
                j += i
            else:
                j = ir + 1
        a[i], ib[i] = ra, ii
    return np.array(a[1:]), np.array(ib[1:])


def sample_replace_prob(rng: RRandom, values: np.ndarray, size: int, prob: np.ndarray) -> np.ndarray:
    """``sample(values, size, replace = TRUE, prob = prob)`` for fewer than 200 categories."""
    p = np.asarray(prob, dtype=np.float64).copy()
    # FixupProb: p / sum(p) (sequential sum)
    tot = 0.0
    for v in p:
        tot += float(v)
    p = p / tot
    ps, perm = _revsort_perm(p)
    cum = ps.copy()
    for i in range(1, len(cum)):
        cum[i] = cum[i] + cum[i - 1]
    out = np.empty(int(size), dtype=np.asarray(values).dtype)
    nm1 = len(cum) - 1
    chunk = 1 << 22
    done = 0
    vals = np.asarray(values)
    while done < size:
        cnt = min(chunk, size - done)
        u = rng.unif_rand(cnt)
        # first j in 0..n-2 with u <= cum[j], else n-1
        j = np.full(cnt, nm1, dtype=np.int64)
        for q in range(nm1 - 1, -1, -1):
            j = np.where(u <= cum[q], q, j)
        out[done:done + cnt] = vals[perm[j] - 1]
        done += cnt
    return out


def ranM2(m: int, p: int, seedn) -> dict:
    """``ranM2(m, p, seedn)`` / ``ranM(scdata, p, seedn)`` (R/ranM2.R:11-35, R/ranM.R:11-33).

    Returns the m x p very-sparse ternary matrix as dgCMatrix slots:
    ``{"Dim": (m, p), "p": int32[p+1], "i": int32[nnz], "x": float64[nnz]}`` with rows ascending inside a
    column.  ``R[i, j] = x0[(i-1)*p + j]`` (``Matrix(x0, nrow = m, byrow = TRUE)``).  A non-integer
    ``seedn`` (the reference's 0.5 sentinel) means "unseeded": numpy's entropy is used instead of R's clock.
    """
    m = int(m)
    p = int(p)
    s = math.sqrt(m)
    if float(seedn) % 1 == 0:
        rng = RRandom(int(seedn))
    else:
        rng = RRandom(int(np.random.SeedSequence().entropy % (2 ** 31)))
    vals = np.array([math.sqrt(s), 0.0, -math.sqrt(s)])
    prob = np.array([1 / (2 * s), 1 - 1 / s, 1 / (2 * s)])
    x0 = sample_replace_prob(rng, vals, m * p, prob)
    nz = np.flatnonzero(x0)  # positions in row-major (gene-major) order
    rows = (nz // p).astype(np.int32)
    cols = (nz % p).astype(np.int32)
    order = np.lexsort((rows, cols))  # by column, rows ascending inside a column
    rows = rows[order]
    cols = cols[order]
    x = x0[nz][order]
    colptr = np.zeros(p + 1, dtype=np.int32)
    np.add.at(colptr, cols + 1, 1)
    colptr = np.cumsum(colptr, dtype=np.int64).astype(np.int32)
    return {"Dim": (m, p), "p": colptr, "i": rows, "x": x}


def ranM(scdata, p: int, seedn) -> dict:
    """``ranM(scdata, p, seedn)``: only ``nrow(scdata)`` is used (R/ranM.R:12)."""
    return ranM2(scdata.shape[0], p, seedn)


def r_sample_perm(n: int, seed=None) -> np.ndarray:
    """``set.seed(seed); sample(n)`` (R/SHARP.R:495-498); ``seed=None`` = unseeded."""
    if seed is None:
        rng = RRandom(int(np.random.SeedSequence().entropy % (2 ** 31)))
    else:
        rng = RRandom(int(seed))
    return rng.sample_perm(n)
