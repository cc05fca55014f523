// rrng.cpp -- R-compatible random numbers for the host glue (no R interpreter next to this library).
//
// The reference draws its projection matrices and its cell shuffle with base R:
//   set.seed(seedn); sample(c(sqrt(s), 0, -sqrt(s)), m*p, replace = TRUE, prob = c(1/(2s), 1-1/s, 1/(2s)))
//   Matrix(x0, nrow = m, byrow = TRUE, sparse = TRUE)                       (R/ranM.R:17-30, R/ranM2.R:17-32)
//   set.seed(50); sample(ncells)                                             (R/SHARP.R:495-498)
// Where R exists those calls stay in R and the dgCMatrix slots / the permutation are handed to the C ABI.
// These two entry points reproduce the same streams (Mersenne-Twister + Inversion + Rejection, R >= 3.6;
// SURVEY.md Appendix A.1) so that a seeded run is identical without R.  sharp_b200/rrng.py is the readable
// restatement the tests compare this file with.
#include <algorithm>
#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#endif
#include <cstdlib>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/sharp_b200.h"

namespace {

struct RMersenne {
    static constexpr int N = 624, M = 397;
    uint32_t mt[N];
    int mti;

    explicit RMersenne(uint32_t seed) {
        // set.seed(): Randomize(kind) -> RNG_Init: 50 scrambling steps, then 625 words of i_seed;
        // FixupSeeds sets i_seed[0] (= mti) to N, so the first draw regenerates the table
        for (int j = 0; j < 50; j++) seed = 69069u * seed + 1u;
        seed = 69069u * seed + 1u; /* i_seed[0], overwritten by mti */
        for (int j = 0; j < N; j++) {
            seed = 69069u * seed + 1u;
            mt[j] = seed;
        }
        mti = N;
    }
    void regen() {
        static const uint32_t mag01[2] = {0x0u, 0x9908b0dfu};
        int kk;
        uint32_t y;
        for (kk = 0; kk < N - M; kk++) {
            y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
            mt[kk] = mt[kk + M] ^ (y >> 1) ^ mag01[y & 1u];
        }
        for (; kk < N - 1; kk++) {
            y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
            mt[kk] = mt[kk + (M - N)] ^ (y >> 1) ^ mag01[y & 1u];
        }
        y = (mt[N - 1] & 0x80000000u) | (mt[0] & 0x7fffffffu);
        mt[N - 1] = mt[M - 1] ^ (y >> 1) ^ mag01[y & 1u];
        mti = 0;
    }
    inline uint32_t next32() {
        if (mti >= N) regen();
        uint32_t y = mt[mti++];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        return y;
    }
    static inline double to_unif(uint32_t y) {
        double u = (double)y * 2.3283064365386963e-10; /* MT_genrand: [0,1) */
        const double i2_32m1 = 2.328306437080797e-10;  /* fixup(): strictly inside (0,1) */
        if (u <= 0.0) return 0.5 * i2_32m1;
        if (1.0 - u <= 0.0) return 1.0 - 0.5 * i2_32m1;
        return u;
    }
    inline double unif() { return to_unif(next32()); }
};

inline uint32_t mt_temper(uint32_t y) {
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

// One whole regeneration of the MT19937 table (the same recurrence as RMersenne::regen, written without the table
// look-up so that the compiler vectorises it: the dependency distances are 227 and 397 words), then the indices of the
// words whose TEMPERED value exceeds thr, in order.  Returns their number.
#if defined(__x86_64__) && defined(__GNUC__)
// The same block with AVX-512 intrinsics (16 words per operation, masked tails, the screening result as a mask register):
// 2x the AVX2 clone on the CPUs that have it.  Chosen at run time (mt_block_screen below).
__attribute__((target("avx512f"))) static inline __m512i mt_twist16(__m512i cur, __m512i nxt, __m512i far) {
    const __m512i y = _mm512_or_si512(_mm512_and_si512(cur, _mm512_set1_epi32((int)0x80000000u)),
                                      _mm512_and_si512(nxt, _mm512_set1_epi32(0x7fffffff)));
    const __m512i odd = _mm512_sub_epi32(_mm512_setzero_si512(), _mm512_and_si512(y, _mm512_set1_epi32(1)));
    return _mm512_xor_si512(_mm512_xor_si512(far, _mm512_srli_epi32(y, 1)), _mm512_and_si512(odd, _mm512_set1_epi32((int)0x9908b0dfu)));
}
__attribute__((target("avx512f"))) static int mt_block_screen_avx512(uint32_t *mt, uint32_t thr, uint32_t *hits) {
    constexpr int N = 624, M = 397;
    int kk = 0;
    for (; kk + 16 <= N - M; kk += 16) /* 227 words: 14 full vectors ... */
        _mm512_storeu_si512(mt + kk, mt_twist16(_mm512_loadu_si512(mt + kk), _mm512_loadu_si512(mt + kk + 1), _mm512_loadu_si512(mt + kk + M)));
    {   /* ... and a masked tail of 3 */
        const __mmask16 k = (__mmask16)((1u << (N - M - kk)) - 1u);
        const __m512i r = mt_twist16(_mm512_maskz_loadu_epi32(k, mt + kk), _mm512_maskz_loadu_epi32(k, mt + kk + 1), _mm512_maskz_loadu_epi32(k, mt + kk + M));
        _mm512_mask_storeu_epi32(mt + kk, k, r);
        kk = N - M;
    }
    for (; kk + 16 <= N - 1; kk += 16) /* 396 words that read the NEW words 227 places back: 24 full vectors ... */
        _mm512_storeu_si512(mt + kk, mt_twist16(_mm512_loadu_si512(mt + kk), _mm512_loadu_si512(mt + kk + 1), _mm512_loadu_si512(mt + kk + (M - N))));
    {   /* ... and a masked tail of 12 (up to word 622) */
        const __mmask16 k = (__mmask16)((1u << (N - 1 - kk)) - 1u);
        const __m512i r = mt_twist16(_mm512_maskz_loadu_epi32(k, mt + kk), _mm512_maskz_loadu_epi32(k, mt + kk + 1), _mm512_maskz_loadu_epi32(k, mt + kk + (M - N)));
        _mm512_mask_storeu_epi32(mt + kk, k, r);
    }
    {
        const uint32_t y = (mt[N - 1] & 0x80000000u) | (mt[0] & 0x7fffffffu);
        mt[N - 1] = mt[M - 1] ^ (y >> 1) ^ ((0u - (y & 1u)) & 0x9908b0dfu);
    }
    const __m512i thrv = _mm512_set1_epi32((int)thr);
    int nh = 0;
    for (int i = 0; i < N; i += 16) { /* 39 vectors exactly */
        __m512i y = _mm512_loadu_si512(mt + i);
        y = _mm512_xor_si512(y, _mm512_srli_epi32(y, 11));
        y = _mm512_xor_si512(y, _mm512_and_si512(_mm512_slli_epi32(y, 7), _mm512_set1_epi32((int)0x9d2c5680u)));
        y = _mm512_xor_si512(y, _mm512_and_si512(_mm512_slli_epi32(y, 15), _mm512_set1_epi32((int)0xefc60000u)));
        y = _mm512_xor_si512(y, _mm512_srli_epi32(y, 18));
        unsigned k = (unsigned)_mm512_cmpgt_epu32_mask(y, thrv);
        while (k) {
            hits[nh++] = (uint32_t)(i + __builtin_ctz(k));
            k &= k - 1;
        }
    }
    return nh;
}
__attribute__((target_clones("avx2", "default")))
#endif
int mt_block_screen_generic(uint32_t *mt, uint32_t thr, uint32_t *hits) {
    constexpr int N = 624, M = 397;
    for (int kk = 0; kk < N - M; kk++) {
        const uint32_t y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
        mt[kk] = mt[kk + M] ^ (y >> 1) ^ ((0u - (y & 1u)) & 0x9908b0dfu);
    }
    for (int kk = N - M; kk < N - 1; kk++) {
        const uint32_t y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
        mt[kk] = mt[kk + (M - N)] ^ (y >> 1) ^ ((0u - (y & 1u)) & 0x9908b0dfu);
    }
    {
        const uint32_t y = (mt[N - 1] & 0x80000000u) | (mt[0] & 0x7fffffffu);
        mt[N - 1] = mt[M - 1] ^ (y >> 1) ^ ((0u - (y & 1u)) & 0x9908b0dfu);
    }
    unsigned char over[N];
    for (int i = 0; i < N; i++) over[i] = mt_temper(mt[i]) > thr;
    int nh = 0;
    for (int i = 0; i < N; i += 8) {
        uint64_t any;
        std::memcpy(&any, over + i, 8);
        if (!any) continue;
        for (int j = i; j < i + 8; j++)
            if (over[j]) hits[nh++] = (uint32_t)j;
    }
    return nh;
}

inline int mt_block_screen(uint32_t *mt, uint32_t thr, uint32_t *hits) {
#if defined(__x86_64__) && defined(__GNUC__)
    static const bool wide = __builtin_cpu_supports("avx512f") && !getenv("SHARP_NO_AVX512");
    if (wide) return mt_block_screen_avx512(mt, thr, hits);
#endif
    return mt_block_screen_generic(mt, thr, hits);
}

// R's revsort(a, ib, n): sort a[] into descending order by heapsort, carrying ib[] (NOT stable).
// a and ib are 1-BASED here (element 0 unused), like the shifted pointers of the C original.
void revsort(double *a, int *ib, int n) {
    if (n <= 1) return;
    int l = (n >> 1) + 1, ir = n, i, j, ii;
    double ra;
    for (;;) {
        if (l > 1) {
            l = l - 1;
            ra = a[l];
            ii = ib[l];
        } else {
            ra = a[ir];
            ii = ib[ir];
            a[ir] = a[1];
            ib[ir] = ib[1];
            if (--ir == 1) {
                a[1] = ra;
                ib[1] = ii;
                return;
            }
        }
        i = l;
        j = l << 1;
        while (j <= ir) {
            if (j < ir && a[j] > a[j + 1]) ++j;
            if (ra > a[j]) {
                a[i] = a[j];
                ib[i] = ib[j];
                j += (i = j);
            } else
                j = ir + 1;
        }
        a[i] = ra;
        ib[i] = ii;
    }
}

}  // namespace

extern "C" {

// ranM2(m, p, seedn) for an integer seed: dgCMatrix slots of the m x p matrix (colptr[p+1], rowidx ascending
// inside a column, x = +-sqrt(sqrt(m))).  cap = room in rowidx/x; *nnz is always set; returns SHARP_E_NOMEM when
// cap is too small (call again with *nnz).
int sharp_r_ranm(int m, int p, int64_t seed, int32_t *colptr, int32_t *rowidx, double *x, int64_t cap, int64_t *nnz) {
    if (m <= 0 || p <= 0 || !colptr || !nnz) return SHARP_E_ARG;
    const double s = std::sqrt((double)m);
    const double vals[3] = {std::sqrt(s), 0.0, -std::sqrt(s)};
    double pr1[4] = {0.0, 1.0 / (2.0 * s), 1.0 - 1.0 / s, 1.0 / (2.0 * s)};
    double sum = 0.0;
    for (int i = 1; i <= 3; i++) sum += pr1[i];
    for (int i = 1; i <= 3; i++) pr1[i] /= sum; /* FixupProb */
    int perm1[4] = {0, 1, 2, 3};
    revsort(pr1, perm1, 3);
    const double *pr0 = pr1 + 1;
    const int *perm = perm1 + 1;
    double pr[3] = {pr0[0], pr0[1], pr0[2]};
    for (int i = 1; i < 3; i++) pr[i] += pr[i - 1];
    const double v0 = vals[perm[0] - 1], v1 = vals[perm[1] - 1], v2 = vals[perm[2] - 1];
    // ProbSampleReplace: first j in 0..n-2 with rU <= p[j], else n-1.  to_unif is monotone in the 32-bit word, so
    // the common case (the most probable value, 0) is decided by one integer comparison against the largest word
    // whose uniform is <= p[0]
    uint32_t thr = (uint32_t)std::floor(pr[0] / 2.3283064365386963e-10);
    while (thr < 0xffffffffu && RMersenne::to_unif(thr + 1u) <= pr[0]) thr++;
    while (thr > 0u && !(RMersenne::to_unif(thr) <= pr[0])) thr--;
    const bool fast = (v0 == 0.0);
    RMersenne rng((uint32_t)seed);
    std::vector<uint32_t> pos;   /* (row * p + col) of the non-zeros, row-major order */
    std::vector<double> val;
    pos.reserve((size_t)((double)p * s * 1.1) + 1024);
    val.reserve(pos.capacity());
    const int64_t total = (int64_t)m * p;
    if (total > 0xffffffffLL) return SHARP_E_LIMIT;
    auto classify = [&](int64_t q, uint32_t w) {
        const double u = RMersenne::to_unif(w);
        const double v = (u <= pr[0]) ? v0 : (u <= pr[1]) ? v1 : v2;
        if (v != 0.0) {
            pos.push_back((uint32_t)q);
            val.push_back(v);
        }
    };
    if (fast) {
        /* m * p draws (14 M for the 10x-brain shape) of which ~1/sqrt(m) are non-zero: the state table is regenerated,
           tempered and screened against the threshold 624 words at a time in vectorised loops (AVX2 where the CPU has
           it); only the words above the threshold take the scalar path.  Same stream, same order. */
        uint32_t hits[RMersenne::N];
        for (int64_t q0 = 0; q0 < total; q0 += RMersenne::N) {
            const int cnt = (int)std::min<int64_t>(RMersenne::N, total - q0);
            const int nh = mt_block_screen(rng.mt, thr, hits);
            for (int h = 0; h < nh; h++) {
                const int i = (int)(hits[h] >> 0) & 0x3ff;
                if (i < cnt) classify(q0 + i, mt_temper(rng.mt[i]));
            }
        }
    } else {
        for (int64_t q = 0; q < total; q++) classify(q, rng.next32());
    }
    const int64_t nz = (int64_t)pos.size();
    *nnz = nz;
    std::memset(colptr, 0, sizeof(int32_t) * (size_t)(p + 1));
    for (int64_t q = 0; q < nz; q++) colptr[pos[q] % (uint32_t)p + 1]++;
    for (int j = 0; j < p; j++) colptr[j + 1] += colptr[j];
    if (nz > cap || (nz > 0 && (!rowidx || !x))) return SHARP_E_NOMEM;
    std::vector<int32_t> fill(colptr, colptr + p);
    for (int64_t q = 0; q < nz; q++) { /* row-major scan => rows ascend inside every column */
        const uint32_t j = pos[q] % (uint32_t)p, i = pos[q] / (uint32_t)p;
        const int32_t at = fill[j]++;
        rowidx[at] = (int32_t)i;
        x[at] = val[q];
    }
    return SHARP_OK;
}

// set.seed(seed); sample(n): 1-based permutation (R >= 3.6, sample.kind = "Rejection")
int sharp_r_sample_perm(int64_t seed, int64_t n, int64_t *out) {
    if (n < 0 || (n > 0 && !out)) return SHARP_E_ARG;
    if (n > 2147483647LL) return SHARP_E_LIMIT;
    RMersenne rng((uint32_t)seed);
    std::vector<int32_t> x((size_t)n);
    for (int64_t i = 0; i < n; i++) x[i] = (int32_t)i;
    int64_t remaining = n;
    for (int64_t i = 0; i < n; i++) {
        const double dn = (double)remaining;
        // R_unif_index(dn): bits = ceil(log2(dn)); repeat dv = rbits(bits) until dv < dn
        int64_t j;
        if (dn <= 0) j = 0;
        else {
            // bits = (int) ceil(log2(dn)); dn is an integer < 2^31, so this is the bit length of dn - 1 (exact, no libm)
            const uint32_t dm1 = (uint32_t)remaining - 1u;
            const int bits = dm1 ? 32 - __builtin_clz(dm1) : 0;
            double dv;
            do {
                int64_t v = 0;
                for (int nb = 0; nb <= bits; nb += 16) {
                    int v1 = (int)std::floor(rng.unif() * 65536);
                    v = 65536 * v + v1;
                }
                const int64_t one64 = 1L;
                if (bits < 64) v &= ((one64 << bits) - 1);
                dv = (double)v;
            } while (dn <= dv);
            j = (int64_t)dv;
        }
        out[i] = (int64_t)x[j] + 1;
        x[j] = x[--remaining];
    }
    return SHARP_OK;
}


// The label tail of SHARP_unlimited (R/SHARP_unlimited.R:166-183), one pass each instead of R's table()/match() on
// 1.3 M-element vectors: final = tf[fColor code]; clusters with fewer than `merge_thre` cells take the smallest such id
// (:168-176, skipped for merge_thre <= 0); ids renumbered 1.. by decreasing size, ties in the STRING order of the ids
// (names(sort(table(f), decreasing = TRUE)): table() orders its names as character, the sort is stable).
int sharp_labels_combine(int nparts, const int64_t *part_start, const int32_t *part_off, const int32_t *pred,
                         const int32_t *tf, int ntf, int merge_thre, int32_t *out, int64_t *counts, int *n_labels) {
    if (nparts < 1 || !part_start || !part_off || !pred || !tf || ntf < 1 || !out || !n_labels) return SHARP_E_ARG;
    const int64_t n = part_start[nparts];
    int mx = 0;
    for (int q = 0; q < ntf; q++) {
        if (tf[q] < 1) return SHARP_E_ARG;
        mx = std::max(mx, tf[q]);
    }
    std::vector<int64_t> cnt((size_t)mx + 1, 0);
    for (int t = 0; t < nparts; t++) {
        const int32_t *lut = tf + part_off[t];
        const int lim = ntf - part_off[t];
        for (int64_t i = part_start[t]; i < part_start[t + 1]; i++) {
            const int c = pred[i];
            if (c < 1 || c > lim) return SHARP_E_ARG;
            const int v = lut[c - 1];
            out[i] = v;
            cnt[v]++;
        }
    }
    std::vector<int32_t> map((size_t)mx + 1, 0);
    for (int v = 0; v <= mx; v++) map[v] = v;
    if (merge_thre > 0) {
        int smallest = 0;
        for (int v = 1; v <= mx; v++)
            if (cnt[v] > 0 && cnt[v] < merge_thre) { smallest = v; break; }
        if (smallest) {
            int64_t moved = 0;
            for (int v = smallest + 1; v <= mx; v++)
                if (cnt[v] > 0 && cnt[v] < merge_thre) { map[v] = smallest; moved += cnt[v]; cnt[v] = 0; }
            cnt[smallest] += moved;
        }
    }
    std::vector<int> ids;
    for (int v = 1; v <= mx; v++)
        if (cnt[v] > 0) ids.push_back(v);
    std::vector<std::string> name((size_t)mx + 1);
    for (int v : ids) name[v] = std::to_string(v);
    std::sort(ids.begin(), ids.end(), [&](int a, int b) { return name[a] < name[b]; });
    std::stable_sort(ids.begin(), ids.end(), [&](int a, int b) { return cnt[a] > cnt[b]; });
    std::vector<int32_t> code((size_t)mx + 1, 0);
    for (size_t r = 0; r < ids.size(); r++) {
        code[ids[r]] = (int32_t)(r + 1);
        if (counts) counts[r] = cnt[ids[r]];
    }
    for (int v = 0; v <= mx; v++) map[v] = code[map[v]];
    for (int64_t i = 0; i < n; i++) out[i] = map[out[i]];
    *n_labels = (int)ids.size();
    return SHARP_OK;
}


// clusterID = match(y, unique(y)) (R/SHARP.R:429-432, 828-832) for non-negative integer ids below `nvals`: codes 1.. in
// order of first appearance, one pass.  uniq (optional, nvals entries) receives unique(y); returns the number of
// distinct ids, or -1 for an id out of range.
int sharp_first_appearance_codes(const int32_t *y, int64_t n, int nvals, int32_t *codes, int32_t *uniq) {
    if (!y || !codes || nvals < 1) return -1;
    std::vector<int32_t> code((size_t)nvals, 0);
    int next = 0;
    for (int64_t i = 0; i < n; i++) {
        const int32_t v = y[i];
        if (v < 0 || v >= nvals) return -1;
        int32_t &c = code[v];
        if (!c) {
            c = ++next;
            if (uniq) uniq[next - 1] = v;
        }
        codes[i] = c;
    }
    return next;
}

}  // extern "C"
