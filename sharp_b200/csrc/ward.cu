// ward.cu -- K3: agglomerative clustering with the exact merge order of stats::hclust (Fortran hclust.f).
//
// Replaces stats::hclust(d, hmethod) at R/get_opt_hclust.R:77.  One CTA per problem, many problems per
// launch (the 2000-cell blocks of every ensemble member, or the small meta-clustering problems of wMetaC /
// sMetaC).  The algorithm is hclust.f's nearest-neighbour list with a global minimum per step -- NOT an
// NN-chain -- because the similarity matrices of wMetaC are full of exact ties (S = 0 / 1) and the labels
// must come out identical to the reference: every argmin below is a reduction under the order
// (value, index), which returns what hclust.f's sequential scan with a strict `<` returns.
//
// Memory: the working distance matrix is the full symmetric n x ld array in global memory (it lives in L2 /
// HBM; 32 MB for a 2000-cell block), so every read is a contiguous row segment; the only strided accesses are
// the mirror writes of the updated row.  NN list, flags and cluster sizes live in shared memory.
// The kernel is latency-bound by design (n-1 dependent steps); throughput comes from the number of problems
// resident at once (3 CTAs of 256 threads per SM) -- callers keep several waves in flight on different streams.
#include <algorithm>
#include <cstdlib>

#include "devutil.cuh"
#include "internal.cuh"

namespace sharp {

// Lance-Williams update of hclust.f, with the same (unfused) operation order.
__device__ __forceinline__ double lance_williams(int method, double d1, double d2, double d12, double mi, double mj,
                                                 double mk) {
    switch (method) {
    case SHARP_WARD_D:
    case SHARP_WARD_D2: {
        double t1 = __dmul_rn(mi + mk, d1);
        double t2 = __dmul_rn(mj + mk, d2);
        double t3 = __dmul_rn(mk, d12);
        double r = __dsub_rn(__dadd_rn(t1, t2), t3);
        return __ddiv_rn(r, mi + mj + mk);
    }
    case SHARP_SINGLE: return fmin(d1, d2);
    case SHARP_COMPLETE: return fmax(d1, d2);
    case SHARP_AVERAGE: return __ddiv_rn(__dadd_rn(__dmul_rn(mi, d1), __dmul_rn(mj, d2)), mi + mj);
    case SHARP_MCQUITTY: return __ddiv_rn(__dadd_rn(d1, d2), 2.0);
    case SHARP_MEDIAN: return __ddiv_rn(__dsub_rn(__dadd_rn(d1, d2), __ddiv_rn(d12, 2.0)), 2.0);
    default: { /* SHARP_CENTROID */
        double a = __dadd_rn(__dmul_rn(mi, d1), __dmul_rn(mj, d2));
        double b = __ddiv_rn(__dmul_rn(__dmul_rn(mi, mj), d12), mi + mj);
        return __ddiv_rn(__dsub_rn(a, b), mi + mj);
    }
    }
}

// Lane-strided scan of row[j], j in [j0, j1), for the first minimum under (value, index).  The loads of a whole
// 256-element chunk are issued before any of them is used (8 independent loads in flight per lane): the scan is a
// chain of L2 round trips otherwise.  CHECK: skip retired columns (flag == 0).
constexpr int SCAN_U = 8;
template <bool CHECK>
__device__ __forceinline__ DI scan_row(const double *row, const unsigned char *flag, int j0, int j1, int lane) {
    DI best;
    best.d = SHARP_INF;
    best.i = INT_MAX;
    for (int base = j0; base < j1; base += 32 * SCAN_U) {
        double v[SCAN_U];
#pragma unroll
        for (int u = 0; u < SCAN_U; u++) {
            const int j = base + u * 32 + lane;
            v[u] = (j < j1) ? row[j] : SHARP_INF;
        }
#pragma unroll
        for (int u = 0; u < SCAN_U; u++) {
            const int j = base + u * 32 + lane;
            if (j < j1 && (!CHECK || flag[j]) && v[u] < best.d) { best.d = v[u]; best.i = j; }
        }
    }
    return best;
}

constexpr int LW_U = 8;                   // Lance-Williams columns per thread per pass

constexpr int RESCAN_PF = 16;  // row elements per lane prefetched for a rescan before the Lance-Williams pass

// Step structure (n - 1 dependent steps, one CTA per problem; the matrices of all resident problems exceed L2, so
// every batch of loads is an HBM round trip and the step is organised to overlap them):
//   A  argmin over the NN list                      -> (i2, j2)
//   B  list of rows whose NN was i2 or j2 (from the PRE-update NN list)
//   C  prefetch: the first 512 columns of each (row, segment) rescan are loaded into registers
//   D  Lance-Williams update of row / column i2 (its loads overlap C's); the new column is also kept in shared
//      memory (rcol) because the prefetched rescans saw the OLD column i2
//   E  rescans: minimum over active j > i, j != i2, j2 of the prefetched data (+ the rest of long rows), combined
//      with the candidate (rcol[i], i2) -- exactly the minimum hclust.f finds after its update
template <int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) hclust_kernel(HcProb *probs, int method, int only_fallback) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ DI red[THREADS / 32];
    __shared__ int s_cnt;

    HcProb &P = probs[blockIdx.x];
    const int n = P.n;
    if (n < 2) return;
    const int ld = P.ld;
    double *D = P.Dw; /* read and written by the whole CTA: no __restrict__, no ld.global.nc */
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = THREADS / 32;

    double *disnn = reinterpret_cast<double *>(smem_raw);  // [n]
    double *rcol = disnn + n;                               // [n] new column i2 of the current step
    DI *part = reinterpret_cast<DI *>(rcol + n);            // [NW][NW] partial minima of the rescans
    int *nn = reinterpret_cast<int *>(part + NW * NW);      // [n]
    int *membr = nn + n;                                    // [n]
    int *list = membr + n;                                  // [n]
    unsigned char *flag = reinterpret_cast<unsigned char *>(list + n);  // [n]

    if (only_fallback) { /* second pass after hclust_rnn_kernel: only the problems it handed back (exact ties) */
        if (!P.fallback) return;
        const double *src = P.D;
        for (size_t idx = tid; idx < (size_t)n * ld; idx += THREADS) D[idx] = src[idx];
        __syncthreads();
    }
    if (method == SHARP_WARD_D2) { /* hclust.f: iOpt 8 works on squared dissimilarities */
        for (size_t idx = tid; idx < (size_t)n * ld; idx += THREADS) D[idx] = __dmul_rn(D[idx], D[idx]);
    }
    for (int i = tid; i < n; i += THREADS) {
        flag[i] = 1;
        membr[i] = 1;
        disnn[i] = SHARP_INF;
        nn[i] = -1;
    }
    if (tid == 0) s_cnt = 0;
    __syncthreads();

    // initial nearest neighbours: NN(i) = first minimum over j > i
    for (int i = warp; i < n - 1; i += NW) {
        DI best = warp_argmin_redux(scan_row<false>(D + (size_t)i * ld, flag, i + 1, n, lane));
        if (lane == 0) {
            nn[i] = (best.i == INT_MAX) ? -1 : best.i;
            disnn[i] = best.d;
        }
    }
    __syncthreads();

    for (int step = 0; step < n - 1; ++step) {
        // ---- A: least dissimilarity over the NN list (first strict minimum over i) ----
        DI c;
        c.d = SHARP_INF;
        c.i = INT_MAX;
        for (int i = tid; i < n - 1; i += THREADS) {
            if (flag[i]) {
                double d = disnn[i];
                if (d < c.d) { c.d = d; c.i = i; }
            }
        }
        c = block_argmin<THREADS>(c, red);
        if (c.i == INT_MAX) { /* NaN / Inf in the dissimilarities: R's hclust stops */
            if (tid == 0) P.status = 12;
            return;
        }
        const int im = c.i, jm = nn[im];
        const int i2 = min(im, jm), j2 = max(im, jm);
        const double *rowi = D + (size_t)i2 * ld;
        const double *rowj = D + (size_t)j2 * ld;
        const double d12 = rowi[j2];
        const double mi = (double)membr[i2], mj = (double)membr[j2];
        if (tid == 0) {
            P.ia[step] = i2 + 1;
            P.ib[step] = j2 + 1;
            P.crit[step] = (method == SHARP_WARD_D2) ? sqrt(c.d) : c.d;
        }
        // ---- B: rows whose NN is about to disappear ----
        for (int i = tid; i < n - 1; i += THREADS) {
            if (flag[i] && i != i2 && i != j2) {
                const int q = nn[i];
                if (q == i2 || q == j2) list[atomicAdd(&s_cnt, 1)] = i;
            }
        }
        __syncthreads();
        const int cnt = s_cnt;
        // ---- C: prefetch the first rescan pass (NW rows at most; fewer rows -> several warps per row) ----
        const int nr = min(NW, cnt);
        const int wpr = nr > 0 ? NW / nr : 1;
        const int rr = warp / wpr, seg = warp - rr * wpr;
        const bool scanning = rr < nr;
        int srow = 0, j0 = 0, j1 = 0;
        double pv[RESCAN_PF];
        if (scanning) {
            srow = list[rr];
            const int len = n - (srow + 1);
            const int seglen = (((len + wpr - 1) / wpr) + 31) & ~31;
            j0 = srow + 1 + seg * seglen;
            j1 = min(n, j0 + seglen);
            const double *row = D + (size_t)srow * ld;
#pragma unroll
            for (int u = 0; u < RESCAN_PF; u++) {
                const int j = j0 + u * 32 + lane;
                pv[u] = (j < j1) ? row[j] : SHARP_INF;
            }
        }
        // ---- D: update dissimilarities from the new cluster (kept under index i2); j2 retires ----
        DI nb;
        nb.d = SHARP_INF;
        nb.i = INT_MAX;
        for (int kb = 0; kb < n; kb += THREADS * LW_U) { /* both rows' loads of LW_U columns in flight per thread */
            double a[LW_U], b[LW_U];
            bool act[LW_U];
#pragma unroll
            for (int u = 0; u < LW_U; u++) {
                const int k = kb + u * THREADS + tid;
                act[u] = k < n && flag[k] && k != i2 && k != j2;
                a[u] = act[u] ? rowi[k] : 0.0;
                b[u] = act[u] ? rowj[k] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < LW_U; u++) {
                if (!act[u]) continue;
                const int k = kb + u * THREADS + tid;
                const double r = lance_williams(method, a[u], b[u], d12, mi, mj, (double)membr[k]);
                D[(size_t)i2 * ld + k] = r;
                D[(size_t)k * ld + i2] = r;
                if (k > i2) {
                    if (r < nb.d) { nb.d = r; nb.i = k; }
                } else {
                    rcol[k] = r;
                    if (r < disnn[k]) { /* hclust.f "FIX by JB": i2 may become the NN of a smaller index */
                        disnn[k] = r;
                        nn[k] = i2;
                    }
                }
            }
        }
        nb = block_argmin<THREADS>(nb, red); /* its barriers also order the reads of nn / membr / s_cnt above */
        if (tid == 0) {
            flag[j2] = 0;
            membr[i2] = membr[i2] + membr[j2];
            disnn[i2] = nb.d;
            nn[i2] = (nb.i == INT_MAX) ? -1 : nb.i;
            s_cnt = 0;
        }
        // ---- E: rescans ----
        if (scanning) {
            DI best;
            best.d = SHARP_INF;
            best.i = INT_MAX;
#pragma unroll
            for (int u = 0; u < RESCAN_PF; u++) {
                const int j = j0 + u * 32 + lane;
                if (j < j1 && j != i2 && j != j2 && flag[j] && pv[u] < best.d) { best.d = pv[u]; best.i = j; }
            }
            const double *row = D + (size_t)srow * ld;
            for (int base = j0 + 32 * RESCAN_PF; base < j1; base += 32 * SCAN_U) { /* rest of a long segment */
                double v[SCAN_U];
#pragma unroll
                for (int u = 0; u < SCAN_U; u++) {
                    const int j = base + u * 32 + lane;
                    v[u] = (j < j1) ? row[j] : SHARP_INF;
                }
#pragma unroll
                for (int u = 0; u < SCAN_U; u++) {
                    const int j = base + u * 32 + lane;
                    if (j < j1 && j != i2 && j != j2 && flag[j] && v[u] < best.d) { best.d = v[u]; best.i = j; }
                }
            }
            best = warp_argmin_redux(best);
            if (lane == 0) part[rr * NW + seg] = best;
        }
        __syncthreads(); /* part[], rcol[], flag[j2] = 0 and the new row / column i2 in D are visible */
        if (tid < nr) {
            DI best = part[tid * NW];
            for (int sgm = 1; sgm < wpr; sgm++) best = di_better(best, part[tid * NW + sgm]);
            const int i = list[tid];
            if (i < i2) { /* the updated column i2, which the prefetched scan skipped */
                DI cand;
                cand.d = rcol[i];
                cand.i = i2;
                best = di_better(best, cand);
            }
            nn[i] = (best.i == INT_MAX) ? -1 : best.i;
            disnn[i] = best.d;
        }
        // further passes when more than NW rows need a rescan (D is up to date now)
        for (int r0 = NW; r0 < cnt; r0 += NW) {
            __syncthreads(); /* part[] of the previous pass has been consumed */
            const int nr2 = min(NW, cnt - r0);
            const int wpr2 = NW / nr2;
            const int r = warp / wpr2, sg = warp - r * wpr2;
            if (r < nr2) {
                const int i = list[r0 + r];
                const int len = n - (i + 1);
                const int seglen = (((len + wpr2 - 1) / wpr2) + 31) & ~31;
                const int b0 = i + 1 + sg * seglen;
                DI best = warp_argmin_redux(scan_row<true>(D + (size_t)i * ld, flag, b0, min(n, b0 + seglen), lane));
                if (lane == 0) part[r * NW + sg] = best;
            }
            __syncthreads();
            if (tid < nr2) {
                DI best = part[tid * NW];
                for (int sgm = 1; sgm < wpr2; sgm++) best = di_better(best, part[tid * NW + sgm]);
                const int i = list[r0 + tid];
                nn[i] = (best.i == INT_MAX) ? -1 : best.i;
                disnn[i] = best.d;
            }
        }
        __syncthreads();
    }
}

// =====================================================================================================
// K3, round-parallel variant for the 2000-cell feature blocks: all reciprocal-nearest-neighbour pairs per round
// =====================================================================================================
// hclust.f merges the globally closest pair, n - 1 dependent steps.  For a REDUCIBLE linkage (ward.D, ward.D2, single,
// complete, average, mcquitty: d(k, i+j) >= min(d(k,i), d(k,j))) every pair of clusters that are each other's nearest
// neighbour is merged by that sequential algorithm sooner or later, whatever happens elsewhere, and the merge heights
// are monotone -- so all reciprocal pairs of the current partition can be merged AT ONCE and the reference's merge
// order is recovered at the end by sorting the merges by height.  A 2000-cell block takes ~40 rounds instead of 1999
// steps.  Each round is one streaming pass: the n_r x n_r matrix of the current partition is read once and the
// (n_r - m) x (n_r - m) matrix of the next partition is written, compacted, one warp per output row (which needs only
// the one or two input rows of its cluster), with the row minimum -- the next round's nearest neighbour -- found on
// the way.  No strided mirror stores, no dependent round trips: the kernel is bound by HBM throughput
// (about 10.6 n^2 x 8 bytes per problem on the benchmark's blocks).
//   * The merges and therefore the labels are those of hclust.f.  The heights agree to rounding only: a distance
//     between two clusters that were both built after the previous global step is reached through the same
//     Lance-Williams recurrences associated in a different order (a few ulps).
//   * Everything above assumes there are no exact ties.  Any tie that could matter -- a row minimum attained twice,
//     two merges of equal height, no reciprocal pair at all (NaN / Inf) -- and any input on which the rounds are not
//     productive (work guard) sets P.fallback, and the problem is redone from the pristine matrix by the exact kernel
//     (launch_hclust runs hclust_kernel in only_fallback mode right after): tie-heavy input costs time, not parity.
// Buffers: round 1 reads the pristine P.D and writes P.Dw, round 2 writes P.E, round 3 P.Dw, ...
constexpr int RNN_THREADS = 512;
constexpr int RNN_UC = 8;   // columns per lane in flight: rows that do not merge (copy + a few updates)
constexpr int RNN_UM = 4;   // merged rows (two source rows)
typedef unsigned short u16;
constexpr u16 RNN_NONE = 0xffffu;

// ascending bitonic sort of (key, payload) pairs in shared memory; P2 a power of two
template <int THREADS>
__device__ __forceinline__ void block_bitonic_sort_kv(double *key, int *val, int P2) {
    for (int k = 2; k <= P2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < P2 / 2; t += THREADS) {
                const int i = 2 * t - (t & (j - 1));
                const int l = i + j;
                const bool up = ((i & k) == 0);
                const double x = key[i], y = key[l];
                if ((x > y) == up && x != y) {
                    const int vi = val[i], vl = val[l];
                    key[i] = y; key[l] = x;
                    val[i] = vl; val[l] = vi;
                }
            }
            __syncthreads();
        }
    }
}

struct RnnBest {
    double d;
    int i;
    bool tie;
};
__device__ __forceinline__ void rnn_consider(RnnBest &b, double v, int j) {
    if (v < b.d) { b.d = v; b.i = j; b.tie = false; }
    else if (v == b.d) b.tie = true;
}
// warp-wide result in all lanes; *tie: the minimum is attained more than once (or there is no finite entry)
__device__ __forceinline__ DI rnn_finish(const RnnBest &b, bool *tie) {
    DI mine;
    mine.d = b.d;
    mine.i = b.i;
    const DI w = warp_argmin_redux(mine);
    const bool dup = b.i != INT_MAX && b.d == w.d && (b.i != w.i || b.tie);
    *tie = __any_sync(0xffffffffu, dup) || w.i == INT_MAX;
    return w;
}

__global__ void __launch_bounds__(RNN_THREADS, 2) hclust_rnn_kernel(HcProb *probs, int method) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_tie, s_nm;
    __shared__ int tmp_scan[RNN_THREADS];
    constexpr int THREADS = RNN_THREADS, NW = RNN_THREADS / 32;

    HcProb &P = probs[blockIdx.x];
    const int n = P.n;
    if (n < 2) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // shared state of the current partition (index cur) and the next one (cur ^ 1)
    double *dnn0 = reinterpret_cast<double *>(smem_raw);   // [2][n] distance to the nearest neighbour
    double *hrow = dnn0 + 2 * (size_t)n;                    // [n] height of the merge that built new cluster i'
    int *rank = reinterpret_cast<int *>(hrow + n);          // [n] scan scratch
    u16 *nn0 = reinterpret_cast<u16 *>(rank + n);           // [2][n] nearest neighbour
    u16 *size0 = nn0 + 2 * (size_t)n;                       // [2][n] cluster sizes
    u16 *orig0 = size0 + 2 * (size_t)n;                     // [2][n] representative (smallest original index)
    u16 *sA = orig0 + 2 * (size_t)n;                        // [n] first source cluster of new cluster i'
    u16 *sB = sA + n;                                       // [n] second source (RNN_NONE: not merged this round)
    u16 *cmap = sB + n;                                     // [n] old cluster -> index of its cluster in the next partition
    u16 *pl = cmap + n;                                     // [n/2 + 1] kept members of this round's pairs
    unsigned char *kind = reinterpret_cast<unsigned char *>(pl + (n / 2 + 1));  // [n] 0 not merging, 1 kept member, 2 retired member

    int cur = 0;
    for (int i = tid; i < n; i += THREADS) {
        size0[i] = 1;
        orig0[i] = (u16)i;
    }
    if (tid == 0) { s_tie = 0; s_nm = 0; P.fallback = 0; }
    __syncthreads();

    // ---- nearest neighbours of the singletons: one pass over the pristine matrix ----
    const double *A = P.D;
    int ldA = P.ld;
    bool sq = (method == SHARP_WARD_D2); /* hclust.f: iOpt 8 works on squared dissimilarities */
    for (int i = warp; i < n; i += NW) {
        const double *row = A + (size_t)i * ldA;
        RnnBest b;
        b.d = SHARP_INF; b.i = INT_MAX; b.tie = false;
        for (int base = 0; base < n; base += 32 * SCAN_U) {
            double v[SCAN_U];
#pragma unroll
            for (int u = 0; u < SCAN_U; u++) {
                const int j = base + u * 32 + lane;
                v[u] = (j < n) ? row[j] : SHARP_INF;
            }
#pragma unroll
            for (int u = 0; u < SCAN_U; u++) {
                const int j = base + u * 32 + lane;
                if (j < n && j != i) rnn_consider(b, sq ? __dmul_rn(v[u], v[u]) : v[u], j);
            }
        }
        bool tie;
        const DI w = rnn_finish(b, &tie);
        if (lane == 0) {
            nn0[i] = (u16)(w.i == INT_MAX ? 0 : w.i);
            dnn0[i] = w.d;
            if (tie) s_tie = 1;
        }
    }
    __syncthreads();

    int nr = n;
    bool failed = false;
    double work = 0.0;
    int round = 0;
    while (nr > 1) {
        if (s_tie) { failed = true; break; }
        const u16 *nn = nn0 + (size_t)cur * n, *size = size0 + (size_t)cur * n, *orig = orig0 + (size_t)cur * n;
        const double *dnn = dnn0 + (size_t)cur * n;
        u16 *nn2 = nn0 + (size_t)(cur ^ 1) * n, *size2 = size0 + (size_t)(cur ^ 1) * n, *orig2 = orig0 + (size_t)(cur ^ 1) * n;
        double *dnn2 = dnn0 + (size_t)(cur ^ 1) * n;
        // ---- survivors: every cluster except the larger-indexed member of a reciprocal pair ----
        for (int i = tid; i < nr; i += THREADS) {
            const int j = nn[i];
            rank[i] = (nn[j] == i && j < i) ? 0 : 1;
        }
        __syncthreads();
        const int nnew = block_exclusive_scan<THREADS>(rank, nr, tmp_scan);
        const int m = nr - nnew;
        if (m == 0) { failed = true; break; }
        const int base = s_nm;
        for (int i = tid; i < nr; i += THREADS) {
            const int j = nn[i];
            const bool paired = nn[j] == i;
            if (paired && j < i) { /* retired: i - rank[i] = number of retired clusters before i, a dense numbering */
                const int q = base + (i - rank[i]);
                P.ia[q] = (int)orig[j] + 1;
                P.ib[q] = (int)orig[i] + 1;
                P.crit[q] = dnn[i];
                cmap[i] = (u16)rank[j];
                kind[i] = 2;
                pl[i - rank[i]] = (u16)j;
                continue;
            }
            const int ip = rank[i];
            cmap[i] = (u16)ip;
            kind[i] = paired ? 1 : 0;
            sA[ip] = (u16)i;
            orig2[ip] = orig[i];
            if (paired) {
                sB[ip] = (u16)j;
                hrow[ip] = dnn[i];
                size2[ip] = (u16)(size[i] + size[j]);
            } else {
                sB[ip] = RNN_NONE;
                hrow[ip] = 0.0;
                size2[ip] = size[i];
            }
        }
        __syncthreads();
        round++;
        double *B = (round & 1) ? P.Dw : P.E;
        const int ldB = (nnew + 3) & ~3;
        const size_t capB = (round & 1) ? (size_t)n * P.ld : (size_t)P.ecap;
        work += (double)nr * nr;
        if ((size_t)nnew * ldB > capB || !B || (work > 16.0 * n * n && nr > 256)) { failed = true; break; }
        if (nnew == 1) { /* the last merge: nothing left to build */
            if (tid == 0) s_nm = base + m;
            nr = 1;
            __syncthreads();
            break;
        }
        // ---- one warp per row of the next partition's matrix.  Pass 1 walks the columns in the OLD order (contiguous,
        //      address-independent loads; every value lands at the compacted position of its cluster) and handles the
        //      clusters that do not merge; pass 2 walks this round's pairs densely (no divergence on the 3-update path) ----
        for (int ip = warp; ip < nnew; ip += NW) {
            const int a = sA[ip];
            const u16 b16 = sB[ip];
            const bool im = b16 != RNN_NONE;
            const double *rowa = A + (size_t)a * ldA;
            double *out = B + (size_t)ip * ldB;
            const double ma = (double)size[a];
            RnnBest best;
            best.d = SHARP_INF; best.i = INT_MAX; best.tie = false;
            if (!im) {
                for (int jb = 0; jb < nr; jb += 32 * RNN_UC) {
                    double x[RNN_UC];
#pragma unroll
                    for (int u = 0; u < RNN_UC; u++) {
                        const int j = jb + u * 32 + lane;
                        x[u] = (j < nr) ? rowa[j] : 0.0;
                    }
#pragma unroll
                    for (int u = 0; u < RNN_UC; u++) {
                        const int j = jb + u * 32 + lane;
                        if (j >= nr || kind[j] != 0) continue;
                        const int jp = cmap[j];
                        double v = sq ? __dmul_rn(x[u], x[u]) : x[u];
                        if (jp == ip) v = 0.0;
                        out[jp] = v;
                        if (jp != ip) rnn_consider(best, v, jp);
                    }
                }
                for (int q = lane; q < m; q += 32) { /* this row's cluster a against the new cluster (c, d): I2 = c, J2 = d, K = a */
                    const int c = pl[q], d = nn[c], jp = cmap[c];
                    double xc = rowa[c], xd = rowa[d];
                    if (sq) { xc = __dmul_rn(xc, xc); xd = __dmul_rn(xd, xd); }
                    const double v = lance_williams(method, xc, xd, hrow[jp], (double)size[c], (double)size[d], ma);
                    out[jp] = v;
                    rnn_consider(best, v, jp);
                }
            } else {
                const int b = (int)b16;
                const double *rowb = A + (size_t)b * ldA;
                const double mb = (double)size[b], hi = hrow[ip];
                for (int jb = 0; jb < nr; jb += 32 * RNN_UM) {
                    double x[RNN_UM], y[RNN_UM];
#pragma unroll
                    for (int u = 0; u < RNN_UM; u++) {
                        const int j = jb + u * 32 + lane;
                        x[u] = (j < nr) ? rowa[j] : 0.0;
                        y[u] = (j < nr) ? rowb[j] : 0.0;
                    }
#pragma unroll
                    for (int u = 0; u < RNN_UM; u++) {
                        const int j = jb + u * 32 + lane;
                        if (j >= nr || kind[j] != 0) continue;
                        const int jp = cmap[j];
                        double x1 = x[u], y1 = y[u];
                        if (sq) { x1 = __dmul_rn(x1, x1); y1 = __dmul_rn(y1, y1); }
                        /* new cluster (a, b) against c = j: I2 = a, J2 = b, K = c */
                        const double v = lance_williams(method, x1, y1, hi, ma, mb, (double)size[j]);
                        out[jp] = v;
                        rnn_consider(best, v, jp);
                    }
                }
                for (int q = lane; q < m; q += 32) { /* both are new: the two updates in the order of their heights, like the reference */
                    const int c = pl[q], d = nn[c], jp = cmap[c];
                    if (jp == ip) { out[jp] = 0.0; continue; }
                    double x1 = rowa[c], y1 = rowb[c], x2 = rowa[d], y2 = rowb[d];
                    if (sq) {
                        x1 = __dmul_rn(x1, x1); y1 = __dmul_rn(y1, y1);
                        x2 = __dmul_rn(x2, x2); y2 = __dmul_rn(y2, y2);
                    }
                    const double hj = hrow[jp], mc = (double)size[c], md = (double)size[d];
                    if (hi == hj) s_tie = 1;
                    const bool first = hi < hj; /* this row's pair merges first */
                    const double h1 = first ? hi : hj, h2 = first ? hj : hi;
                    const double p1 = first ? ma : mc, p2 = first ? mb : md; /* sizes of the pair merging first */
                    const double r1 = first ? mc : ma, r2 = first ? md : mb; /* sizes of the other pair's members */
                    /* first merge: (I2, J2) of the earlier pair against each member of the later pair */
                    const double t1 = lance_williams(method, x1, first ? y1 : x2, h1, p1, p2, r1);
                    const double t2 = lance_williams(method, first ? x2 : y1, y2, h1, p1, p2, r2);
                    /* second merge: the later pair (I2 = its kept member, J2 = retired) against the merged earlier pair */
                    const double v = lance_williams(method, t1, t2, h2, r1, r2, p1 + p2);
                    out[jp] = v;
                    rnn_consider(best, v, jp);
                }
            }
            bool tie;
            const DI w = rnn_finish(best, &tie);
            if (lane == 0) {
                nn2[ip] = (u16)(w.i == INT_MAX ? 0 : w.i);
                dnn2[ip] = w.d;
                if (tie) s_tie = 1;
            }
        }
        if (tid == 0) s_nm = base + m;
        __syncthreads();
        A = B;
        ldA = ldB;
        sq = false;
        nr = nnew;
        cur ^= 1;
    }
    if (!failed && s_tie) failed = true;
    // ---- the reference's merge order: ascending height ----
    const int nm = n - 1;
    int P2 = 1;
    while (P2 < nm) P2 <<= 1;
    double *key = reinterpret_cast<double *>(smem_raw);          // [P2]
    int *perm = reinterpret_cast<int *>(key + P2);               // [P2]
    int *ias = perm + P2;                                        // [nm]
    int *ibs = ias + nm;                                         // [nm]
    if (!failed) {
        __syncthreads();
        if (s_nm != nm) failed = true;
    }
    if (!failed) {
        for (int i = tid; i < P2; i += THREADS) {
            key[i] = (i < nm) ? P.crit[i] : SHARP_INF;
            perm[i] = i;
        }
        for (int i = tid; i < nm; i += THREADS) { ias[i] = P.ia[i]; ibs[i] = P.ib[i]; }
        __syncthreads();
        block_bitonic_sort_kv<THREADS>(key, perm, P2);
        for (int i = tid; i + 1 < nm; i += THREADS)
            if (!(key[i] < key[i + 1])) s_tie = 1; /* equal heights (or NaN): the order is the reference's business */
        __syncthreads();
        if (s_tie) failed = true;
    }
    if (failed) {
        if (tid == 0) P.fallback = 1;
        return;
    }
    for (int i = tid; i < nm; i += THREADS) {
        P.ia[i] = ias[perm[i]];
        P.ib[i] = ibs[perm[i]];
        P.crit[i] = (method == SHARP_WARD_D2) ? sqrt(key[i]) : key[i];
    }
}

// =====================================================================================================
// K3, round-parallel variant on TRIANGULAR storage: half the bytes and half the arithmetic of hclust_rnn_kernel
// =====================================================================================================
// Same rounds as hclust_rnn_kernel (all reciprocal-nearest-neighbour pairs of a reducible linkage merge at once, the
// reference's merge order is recovered by sorting the merges by height, any tie hands the problem to the exact
// kernel), but the dissimilarities are symmetric and this kernel reads and writes each of them ONCE per round:
//   * the working matrices hold the strict upper triangle only, row i = columns tri_c0(i) .. nrp-1 (rows start on
//     16-byte boundaries); round 1 reads the upper triangle of the pristine full matrix P.D;
//   * one warp per row of the next partition's matrix, as before, but only the columns RIGHT of the row: the value
//     v(i', j'), i' < j', is computed once.  It is a candidate for the nearest neighbour of row i' -- reduced by the warp,
//     "upper" minimum and its index, like hclust.f's NN list -- and of COLUMN j': column minima are kept as 64-bit
//     order-preserving keys in shared memory, lowered with atomicMin behind a plain-read filter (a few updates per column);
//   * the reciprocal pairs follow from the two: (k, x), k < x, merges iff x is the upper nearest neighbour of k, that
//     distance beats k's column minimum, EQUALS x's column minimum and beats x's upper minimum.  Equal candidates
//     anywhere (a key that meets its own value) raise the tie flag: the problem is redone by the exact kernel;
//   * POSITIONS.  A merged cluster (a, b), a < b, takes the place of its RIGHT member b (its representative, the smallest
//     original index, is tracked separately in orig[]): row b' then needs only the columns right of b, all of them
//     contiguous in rows a and b, and what other rows need from the retired row a -- element j of row a for the rows j
//     between a and b -- is read by consecutive rows from consecutive addresses.  Keeping the LEFT place instead (r2's
//     first version, like hclust.f's I2) made row a' gather column b from the rows between a and b: one 32-byte sector
//     (64 fetched) per 8 bytes used, the larger part of that version's 8.7 n^2 x 8 B of DRAM traffic per problem;
//   * rows are dealt to the warps in (top, bottom) pairs of the triangle: the same amount of work for every warp.
// Tried and dropped (B200, r2): a per-warp shared-memory stash of the pair members' row values, so that the pair pass
// would not re-read them -- 98 KB more shared memory left 28 KB of L1 and the kernel ran 25 % slower (446 vs 357 ms per
// 1.3 M-cell step); one CTA of 16 warps per problem, one or two problems per SM: 1.4x slower than 32 warps per problem.

__device__ __forceinline__ int tri_c0(int i) { return (i + 1) & ~1; }
__device__ __forceinline__ size_t tri_rowoff(int i, int nrp) {
    const size_t t = (size_t)(i >> 1);
    size_t o = 2 * (size_t)nrp * t - 2 * t * t;
    if (i & 1) o += (size_t)nrp - 2 * t;
    return o;
}
__host__ __device__ inline size_t tri_elems(int nr) {
    const size_t nrp = (size_t)((nr + 1) & ~1);
    return nrp * nrp / 2 + nrp;
}
// source matrix of a round: the full pristine matrix (round 1) or a triangular working matrix
struct TriSrc {
    const double *p;
    int full, ld, nrp;
    __device__ __forceinline__ const double *row(int i) const { /* pointer such that row(i)[j] is element (i, j), i < j */
        return full ? p + (size_t)i * ld : p + tri_rowoff(i, nrp) - tri_c0(i);
    }
};

// Running minimum of a row (registers of one lane) and the per-element update, written without branches around the
// common path: the r2 profile of the first version showed ~130 warp instructions per element slot, most of them
// divergence bookkeeping (BSSY / BRA / BSYNC around `continue`s and nested ifs).
struct TriRow {
    double d;
    int i;
    int tie;   /* the running minimum has been met twice */
};

// Column minima live in shared memory as plain doubles and are lowered with a 64-bit integer atomicMin on their bit
// patterns -- the order of the patterns is the numeric order for non-negative values.  A negative value or a NaN
// (impossible for Ward / single / complete / average / mcquitty on proper dissimilarities) raises the flag, and the
// problem is redone by the exact kernel like any tie.
__device__ __forceinline__ void tri_elem(bool valid, double v, int jp, double *out, TriRow &r, double *colmin, int &flag) {
    if (out && valid) out[jp] = v;
    const bool lt = valid && v < r.d;
    const bool eq = valid && v == r.d;
    r.tie = lt ? 0 : (r.tie | (int)eq);
    r.i = lt ? jp : r.i;
    r.d = lt ? v : r.d;
    const double cur = *reinterpret_cast<volatile double *>(colmin + (valid ? jp : 0));
    flag |= (int)(valid && (!(v >= 0.0) || v == cur));
    if (valid && v < cur) {
        const unsigned long long k = (unsigned long long)__double_as_longlong(v + 0.0);
        if (atomicMin(reinterpret_cast<unsigned long long *>(colmin + jp), k) == k) flag |= 1;
    }
}

__device__ __forceinline__ DI tri_row_finish(const TriRow &b, bool *tie) { /* warp-wide (value, index) minimum, all lanes */
    DI mine;
    mine.d = b.d;
    mine.i = b.i;
    const DI w = warp_argmin_redux(mine);
    const bool dup = b.i != INT_MAX && b.d == w.d && (b.i != w.i || b.tie);
    *tie = __any_sync(0xffffffffu, dup);
    return w;
}

// Lance-Williams update with the method fixed at compile time for Ward (METHOD = SHARP_WARD_D; 0 = run-time switch)
template <int METHOD>
__device__ __forceinline__ double tri_lw(int method, double d1, double d2, double d12, double mi, double mj, double mk) {
    if (METHOD == SHARP_WARD_D) {
        const double t1 = __dmul_rn(mi + mk, d1);
        const double t2 = __dmul_rn(mj + mk, d2);
        const double t3 = __dmul_rn(mk, d12);
        return __ddiv_rn(__dsub_rn(__dadd_rn(t1, t2), t3), mi + mj + mk);
    }
    return lance_williams(method, d1, d2, d12, mi, mj, mk);
}

constexpr u16 TRI_MERGING = 0x8000u; /* flag in the column map: the old cluster merges this round (handled by the pair pass) */

template <int TRI_THREADS, int TRI_MINB, int METHOD>
__global__ void __launch_bounds__(TRI_THREADS, TRI_MINB) hclust_tri_kernel(HcProb *probs, int method) {
    constexpr int TRI_NW = TRI_THREADS / 32;
    constexpr bool TRI_PIPE = (TRI_THREADS == 768); /* experiment: fewer warps, software-pipelined row streams */
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_tie, s_nm;
    __shared__ int tmp_scan[TRI_THREADS];
    constexpr int THREADS = TRI_THREADS, NW = TRI_NW;

    HcProb &P = probs[blockIdx.x];
    const int n = P.n;
    if (n < 2) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // [2][n] doubles: half `cur` holds the upper minimum of every row of the current partition, the other half the
    // column minima; both are dead once the round's setup is done and swap roles for the next partition
    double *dnn0 = reinterpret_cast<double *>(smem_raw);
    double *hrow = dnn0 + 2 * (size_t)n;                    // [n] height of the merge that built new cluster i'
    int *rank = reinterpret_cast<int *>(hrow + n);          // [n] scan scratch
    u16 *nn0 = reinterpret_cast<u16 *>(rank + n);           // [2][n] upper nearest neighbour (current / next partition)
    u16 *size0 = nn0 + 2 * (size_t)n;                       // [2][n] cluster sizes
    u16 *orig0 = size0 + 2 * (size_t)n;                     // [2][n] representative (smallest original index)
    u16 *sA = orig0 + 2 * (size_t)n;                        // [n] first source cluster of new cluster i'
    u16 *sB = sA + n;                                       // [n] second source (RNN_NONE: not merged this round)
    u16 *cmap = sB + n;                                     // [n] old cluster -> index in the next partition | TRI_MERGING
    u16 *partner = cmap + n;                                // [n] this round's partner of every current cluster, or RNN_NONE
    u16 *pl = partner + n;                                  // [n/2 + 1] kept members of this round's pairs

    int cur = 0;
    for (int i = tid; i < n; i += THREADS) {
        size0[i] = 1;
        orig0[i] = (u16)i;
    }
    if (tid == 0) { s_tie = 0; s_nm = 0; P.fallback = 0; }
    __syncthreads();

    TriSrc A;
    A.p = P.D; A.full = 1; A.ld = P.ld; A.nrp = 0;
    bool sq = (method == SHARP_WARD_D2); /* hclust.f: iOpt 8 works on squared dissimilarities */
    int nr = n;
    bool failed = false;
    double work = 0.0;
    int round = 0;          /* round 0: nearest neighbours of the singletons, nothing merges, nothing is written */
    int base = 0;
    while (nr > 1) {
        const bool first = (round == 0);
        const u16 *size = size0 + (size_t)cur * n, *orig = orig0 + (size_t)cur * n;
        const u16 *nnu = nn0 + (size_t)cur * n;                             /* upper nearest neighbour of the current rows */
        const double *dup = dnn0 + (size_t)cur * n;                         /* upper minimum of the current rows */
        const double *cmin = dnn0 + (size_t)(cur ^ 1) * n;                  /* column minimum of the current clusters */
        u16 *size2 = size0 + (size_t)(cur ^ 1) * n, *orig2 = orig0 + (size_t)(cur ^ 1) * n;
        int nnew = nr, m = 0;
        if (first) {
            for (int i = tid; i < nr; i += THREADS) {
                cmap[i] = (u16)i; sA[i] = (u16)i; sB[i] = RNN_NONE; hrow[i] = 0.0;
                size2[i] = 1; orig2[i] = (u16)i;
            }
            __syncthreads();
        } else {
            if (s_tie) { failed = true; break; }
            // ---- this round's pairs (k, x), k < x: x = nnu[k], d = dup[k] < column minimum of k, == column minimum of x,
            //      < upper minimum of x.  Equalities that would make the choice ambiguous raise s_tie. ----
            for (int i = tid; i < nr; i += THREADS) partner[i] = RNN_NONE;
            __syncthreads();
            for (int k = tid; k < nr; k += THREADS) {
                const double d = dup[k];
                if (!(d < SHARP_INF)) continue;
                const int x = nnu[k];
                if (d == cmin[k]) s_tie = 1; /* the nearest neighbour of k is not unique (one above, one below) */
                if (d < cmin[k] && d == cmin[x] && d < dup[x]) { partner[k] = (u16)x; partner[x] = (u16)k; }
            }
            __syncthreads();
            if (s_tie) { failed = true; break; }
            for (int i = tid; i < nr; i += THREADS) {
                const int j = partner[i];
                rank[i] = (j != RNN_NONE && j > i) ? 0 : 1; /* the member on the LEFT retires (see the note on positions above) */
            }
            __syncthreads();
            nnew = block_exclusive_scan<THREADS>(rank, nr, tmp_scan);
            m = nr - nnew;
            if (m == 0) { failed = true; break; }
            base = s_nm;
            for (int i = tid; i < nr; i += THREADS) {
                const int j = partner[i];
                const bool paired = j != RNN_NONE;
                if (paired && j > i) { /* retired: i - rank[i] = number of retired clusters before i, a dense numbering */
                    const int q = base + (i - rank[i]);
                    const int oi = (int)orig[i], oj = (int)orig[j];
                    P.ia[q] = min(oi, oj) + 1;
                    P.ib[q] = max(oi, oj) + 1;
                    P.crit[q] = dup[i];
                    cmap[i] = (u16)(rank[j] | TRI_MERGING);
                    pl[i - rank[i]] = (u16)j;
                    continue;
                }
                const int ip = rank[i];
                cmap[i] = (u16)(ip | (paired ? TRI_MERGING : 0));
                if (paired) { /* the merged cluster takes the place of its RIGHT member i; j < i is the left one */
                    sA[ip] = (u16)j;
                    sB[ip] = (u16)i;
                    hrow[ip] = dup[j];
                    size2[ip] = (u16)(size[i] + size[j]);
                    orig2[ip] = min(orig[i], orig[j]);
                } else {
                    sA[ip] = (u16)i;
                    sB[ip] = RNN_NONE;
                    hrow[ip] = 0.0;
                    size2[ip] = size[i];
                    orig2[ip] = orig[i];
                }
            }
            __syncthreads();
            if (nnew == 1) { /* the last merge: nothing left to build */
                if (tid == 0) s_nm = base + m;
                nr = 1;
                __syncthreads();
                break;
            }
        }
        double *B = nullptr;
        const int nrpB = (nnew + 1) & ~1;
        if (!first) {
            B = (round & 1) ? P.Dw : P.E;
            const size_t capB = (round & 1) ? (size_t)n * P.ld : (size_t)P.ecap;
            work += (double)nr * nr;
            if (tri_elems(nnew) > capB || !B || (work > 16.0 * n * n && nr > 256)) { failed = true; break; }
        }
        // the halves of dnn0 / nn0 swap roles: upper minima of the NEXT partition go where this round's column minima were
        // (consumed by the setup above), the next column minima where this round's upper minima were
        double *dup2 = dnn0 + (size_t)(cur ^ 1) * n;
        u16 *nnu2 = nn0 + (size_t)(cur ^ 1) * n;
        double *cmin2 = dnn0 + (size_t)cur * n;
        for (int i = tid; i < nnew; i += THREADS) cmin2[i] = SHARP_INF;
        __syncthreads();
        // ---- one warp per row of the next partition's matrix, columns right of the row only; rows in (top, bottom) pairs ----
        const int half = (nnew + 1) / 2;
        for (int rr = warp; rr < half; rr += NW) {
#pragma unroll 1
            for (int side = 0; side < 2; side++) {
                const int ip = side ? nnew - 1 - rr : rr;
                if (side && ip == rr) break;
                const int a = sA[ip];
                const u16 b16 = sB[ip];
                const bool im = b16 != RNN_NONE;
                const double *rowa = A.row(a);
                double *out = B ? B + tri_rowoff(ip, nrpB) - tri_c0(ip) : nullptr;
                const double ma = (double)size[a];
                TriRow best;
                best.d = SHARP_INF; best.i = INT_MAX; best.tie = 0;
                int flag = 0;
                const int jstart = (a + 1) & ~31; /* aligned start: full sectors */
                if (!im && TRI_PIPE) {
                    /* 24 warps, 85 registers: the next batch of the row is requested before this one is worked on */
                    double x[RNN_UC];
                    unsigned cm[RNN_UC];
#pragma unroll
                    for (int u = 0; u < RNN_UC; u++) {
                        const int j = jstart + u * 32 + lane;
                        const bool in = j > a && j < nr;
                        x[u] = in ? rowa[j] : 0.0;
                        cm[u] = in ? (unsigned)cmap[j] : (unsigned)TRI_MERGING;
                    }
                    for (int jb = jstart; jb < nr; jb += 32 * RNN_UC) {
                        double xn[RNN_UC];
                        unsigned cmn[RNN_UC];
#pragma unroll
                        for (int u = 0; u < RNN_UC; u++) {
                            const int j = jb + 32 * RNN_UC + u * 32 + lane;
                            const bool in = j < nr; /* j > a holds: this is a later batch */
                            xn[u] = in ? rowa[j] : 0.0;
                            cmn[u] = in ? (unsigned)cmap[j] : (unsigned)TRI_MERGING;
                        }
#pragma unroll
                        for (int u = 0; u < RNN_UC; u++) {
                            const double v = sq ? __dmul_rn(x[u], x[u]) : x[u];
                            tri_elem(!(cm[u] & TRI_MERGING), v, (int)(cm[u] & 0x7fffu), out, best, cmin2, flag);
                        }
#pragma unroll
                        for (int u = 0; u < RNN_UC; u++) { x[u] = xn[u]; cm[u] = cmn[u]; }
                    }
                }
                if (!im && !TRI_PIPE) {
                    for (int jb = jstart; jb < nr; jb += 32 * RNN_UC) {
                        double x[RNN_UC];
                        unsigned cm[RNN_UC];
#pragma unroll
                        for (int u = 0; u < RNN_UC; u++) {
                            const int j = jb + u * 32 + lane;
                            const bool in = j > a && j < nr;
                            x[u] = in ? rowa[j] : 0.0;
                            cm[u] = in ? (unsigned)cmap[j] : (unsigned)TRI_MERGING;
                        }
#pragma unroll
                        for (int u = 0; u < RNN_UC; u++) {
                            const double v = sq ? __dmul_rn(x[u], x[u]) : x[u];
                            tri_elem(!(cm[u] & TRI_MERGING), v, (int)(cm[u] & 0x7fffu), out, best, cmin2, flag);
                        }
                    }
                }
                if (!im) {
                    for (int q0 = 0; q0 < m; q0 += 32) { /* this row's cluster a against the new cluster (d, c), d < c: I2 = d, J2 = c, K = a */
                        const int q = q0 + lane;
                        const int c = q < m ? (int)pl[q] : 0;
                        const bool valid = q < m && c > a;
                        const int d = valid ? (int)partner[c] : a + 1, jp = valid ? (int)(cmap[c] & 0x7fffu) : 0;
                        double xc = valid ? rowa[c] : 0.0;
                        double xd = valid ? ((d > a) ? rowa[d] : A.row(d)[a]) : 0.0; /* left of a: element a of row d (consecutive rows a share its sectors) */
                        if (sq) { xc = __dmul_rn(xc, xc); xd = __dmul_rn(xd, xd); }
                        const double v = tri_lw<METHOD>(method, xd, xc, hrow[jp], (double)size[valid ? d : 0], (double)size[c], ma);
                        tri_elem(valid, v, jp, out, best, cmin2, flag);
                    }
                } else {
                    const int b = (int)b16; /* a < b: the new cluster sits at b, so only the columns right of b are needed, all of them in rows a and b */
                    const double *rowb = A.row(b);
                    const double mb = (double)size[b], hi = hrow[ip];
                    const int jstart_b = (b + 1) & ~31;
                    for (int jb = jstart_b; jb < nr; jb += 32 * RNN_UM) {
                        double x[RNN_UM], y[RNN_UM];
                        unsigned cm[RNN_UM];
#pragma unroll
                        for (int u = 0; u < RNN_UM; u++) {
                            const int j = jb + u * 32 + lane;
                            const bool in = j > b && j < nr;
                            x[u] = in ? rowa[j] : 0.0;
                            y[u] = in ? rowb[j] : 0.0;
                            cm[u] = in ? (unsigned)cmap[j] : (unsigned)TRI_MERGING;
                        }
#pragma unroll
                        for (int u = 0; u < RNN_UM; u++) {
                            const int j = jb + u * 32 + lane;
                            double x1 = x[u], y1 = y[u];
                            if (sq) { x1 = __dmul_rn(x1, x1); y1 = __dmul_rn(y1, y1); }
                            /* new cluster (a, b) against c = j: I2 = a, J2 = b, K = c */
                            const double v = tri_lw<METHOD>(method, x1, y1, hi, ma, mb, (double)size[j < nr ? j : 0]);
                            tri_elem(!(cm[u] & TRI_MERGING), v, (int)(cm[u] & 0x7fffu), out, best, cmin2, flag);
                        }
                    }
                    for (int q0 = 0; q0 < m; q0 += 32) { /* both are new: the two updates in the order of their heights, like the reference */
                        const int q = q0 + lane;
                        const int c2 = q < m ? (int)pl[q] : 0;      /* right member of the other pair (its place) */
                        const bool valid = q < m && c2 > b;
                        const int dd = valid ? c2 : b + 1;           /* any index right of b keeps the loads of idle lanes in bounds */
                        const int cc = valid ? (int)partner[c2] : b + 1; /* its left member: anywhere left of dd */
                        const int jp = valid ? (int)(cmap[c2] & 0x7fffu) : 0;
                        double x1 = valid ? ((cc > a) ? rowa[cc] : A.row(cc)[a]) : 0.0, x2 = valid ? rowa[dd] : 0.0;
                        double y1 = valid ? ((cc > b) ? rowb[cc] : A.row(cc)[b]) : 0.0, y2 = valid ? rowb[dd] : 0.0;
                        if (sq) {
                            x1 = __dmul_rn(x1, x1); y1 = __dmul_rn(y1, y1);
                            x2 = __dmul_rn(x2, x2); y2 = __dmul_rn(y2, y2);
                        }
                        const double hj = hrow[jp], mc = (double)size[cc < nr ? cc : 0], md = (double)size[dd < nr ? dd : 0];
                        flag |= (int)(valid && hi == hj);
                        const bool fst = hi < hj; /* this row's pair merges first */
                        const double h1 = fst ? hi : hj, h2 = fst ? hj : hi;
                        const double p1 = fst ? ma : mc, p2 = fst ? mb : md; /* sizes of the pair merging first */
                        const double r1 = fst ? mc : ma, r2 = fst ? md : mb; /* sizes of the other pair's members */
                        /* first merge: (I2, J2) of the earlier pair against each member of the later pair */
                        const double t1 = tri_lw<METHOD>(method, x1, fst ? y1 : x2, h1, p1, p2, r1);
                        const double t2 = tri_lw<METHOD>(method, fst ? x2 : y1, y2, h1, p1, p2, r2);
                        /* second merge: the later pair (I2 = its left member, J2 = right) against the merged earlier pair */
                        const double v = tri_lw<METHOD>(method, t1, t2, h2, r1, r2, p1 + p2);
                        tri_elem(valid, v, jp, out, best, cmin2, flag);
                    }
                }
                bool tie;
                const DI w = tri_row_finish(best, &tie); /* no candidate at all (the last row): no upper neighbour, not a tie */
                if (__any_sync(0xffffffffu, flag != 0)) tie = true;
                if (lane == 0) {
                    nnu2[ip] = (u16)(w.i == INT_MAX ? 0 : w.i);
                    dup2[ip] = w.d;
                    if (tie) s_tie = 1;
                }
            }
        }
        if (tid == 0 && !first) s_nm = base + m;
        __syncthreads();
        if (!first) {
            A.p = B; A.full = 0; A.ld = 0; A.nrp = nrpB;
            sq = false;
        }
        nr = nnew;
        cur ^= 1;
        round++;
    }
    if (!failed && s_tie) failed = true;
    // ---- the reference's merge order: ascending height ----
    const int nm = n - 1;
    int P2 = 1;
    while (P2 < nm) P2 <<= 1;
    double *key = reinterpret_cast<double *>(smem_raw);          // [P2]
    int *perm = reinterpret_cast<int *>(key + P2);               // [P2]
    int *ias = perm + P2;                                        // [nm]
    int *ibs = ias + nm;                                         // [nm]
    if (!failed) {
        __syncthreads();
        if (s_nm != nm) failed = true;
    }
    if (!failed) {
        for (int i = tid; i < P2; i += THREADS) {
            key[i] = (i < nm) ? P.crit[i] : SHARP_INF;
            perm[i] = i;
        }
        for (int i = tid; i < nm; i += THREADS) { ias[i] = P.ia[i]; ibs[i] = P.ib[i]; }
        __syncthreads();
        block_bitonic_sort_kv<THREADS>(key, perm, P2);
        for (int i = tid; i + 1 < nm; i += THREADS)
            if (!(key[i] < key[i + 1])) s_tie = 1; /* equal heights (or NaN): the order is the reference's business */
        __syncthreads();
        if (s_tie) failed = true;
    }
    if (failed) {
        if (tid == 0) P.fallback = 1;
        return;
    }
    for (int i = tid; i < nm; i += THREADS) {
        P.ia[i] = ias[perm[i]];
        P.ib[i] = ibs[perm[i]];
        P.crit[i] = (method == SHARP_WARD_D2) ? sqrt(key[i]) : key[i];
    }
}

static size_t hclust_rnn_smem_bytes(int n) {
    size_t rounds = (size_t)n * (16 + 8 + 4) + (size_t)n * 2 * 9 + ((size_t)n / 2 + 1) * 2 + (size_t)n + 16;
    size_t p2 = 1;
    while ((int)p2 < n - 1) p2 <<= 1;
    size_t sort = p2 * 12 + (size_t)n * 8;
    return (std::max(rounds, sort) + 31) & ~(size_t)15;
}

static size_t hclust_tri_smem_bytes(int n) {
    size_t rounds = (size_t)n * (16 + 8 + 4) + (size_t)n * 2 * 10 + ((size_t)n / 2 + 1) * 2 + 16;
    size_t p2 = 1;
    while ((int)p2 < n - 1) p2 <<= 1;
    size_t sort = p2 * 12 + (size_t)n * 8;
    return (std::max(rounds, sort) + 31) & ~(size_t)15;
}

bool hclust_fast_ok(int max_n, int method) {
    static const bool no_rnn = getenv("SHARP_HCLUST_EXACT") != nullptr; /* development switch */
    const bool reducible = method != SHARP_MEDIAN && method != SHARP_CENTROID;
    return reducible && !no_rnn && max_n > 384 && max_n < 32768 && hclust_tri_smem_bytes(max_n) <= (size_t)SHARP_SMEM_OPTIN - 4096;
}

static size_t hclust_smem_bytes(int n) {
    return ((size_t)n * (8 + 8 + 4 + 4 + 4 + 1) + (size_t)16 * 16 * sizeof(DI) + 15) & ~(size_t)15;
}

static int launch_exact(sharp_ctx *c, HcProb *probs_dev, int nprob, int max_n, int method, int only_fallback) {
    size_t smem = hclust_smem_bytes(max_n);
    if (smem > 200 * 1024)
        return set_error(SHARP_E_LIMIT, "hclust: %d objects exceed the shared-memory NN list (max ~7000)", max_n);
    prof_begin(c, max_n > 384 ? KID_HCLUST : KID_HCLUST_SMALL);
    if (max_n > 384) {
        SHARP_SMEM_OPTIN_ONCE((hclust_kernel<256, 2>), c->device);
        hclust_kernel<256, 2><<<nprob, 256, smem, c->stream>>>(probs_dev, method, only_fallback);
    } else {
        SHARP_SMEM_OPTIN_ONCE((hclust_kernel<128, 4>), c->device);
        hclust_kernel<128, 4><<<nprob, 128, smem, c->stream>>>(probs_dev, method, only_fallback);
    }
    prof_end(c);
    SHARP_CUDA(cudaGetLastError());
    return 0;
}

// fast = 1: feature problems (continuous data, P.D holds a pristine copy of the dissimilarities): the round-parallel
// kernel, then the exact kernel for the problems that reported ties.  fast = 0: hclust.f's order to the last bit.
int launch_hclust(sharp_ctx *c, HcProb *probs_dev, int nprob, int max_n, int method, int fast) {
    if (nprob <= 0) return 0;
    if (method < 1 || method > 8) return set_error(SHARP_E_ARG, "invalid clustering method %d", method);
    if (fast && hclust_fast_ok(max_n, method)) {
        static const bool no_tri = getenv("SHARP_HCLUST_FULL") != nullptr; /* development switch: the full-matrix kernel */
        prof_begin(c, KID_HCLUST);
        if (!no_tri) {
            /* 32 warps per problem, one problem per SM: measured on B200 1.4x faster than 16 warps (one or two problems per SM
               made no difference: the kernel sits at the DRAM efficiency of many concurrent 2 KB row streams, ~2.5 TB/s) */
            static const int tri_threads = getenv("SHARP_TRI_THREADS") ? atoi(getenv("SHARP_TRI_THREADS")) : 1024; /* development switch */
            const size_t tsm = hclust_tri_smem_bytes(max_n);
            if (tri_threads == 1024 && method == SHARP_WARD_D) {
                SHARP_SMEM_OPTIN_ONCE((hclust_tri_kernel<1024, 1, SHARP_WARD_D>), c->device);
                hclust_tri_kernel<1024, 1, SHARP_WARD_D><<<nprob, 1024, tsm, c->stream>>>(probs_dev, method);
            } else if (tri_threads == 1024) {
                SHARP_SMEM_OPTIN_ONCE((hclust_tri_kernel<1024, 1, 0>), c->device);
                hclust_tri_kernel<1024, 1, 0><<<nprob, 1024, tsm, c->stream>>>(probs_dev, method);
            } else if (tri_threads == 768) {
                SHARP_SMEM_OPTIN_ONCE((hclust_tri_kernel<768, 1, SHARP_WARD_D>), c->device);
                SHARP_SMEM_OPTIN_ONCE((hclust_tri_kernel<768, 1, 0>), c->device);
                if (method == SHARP_WARD_D) hclust_tri_kernel<768, 1, SHARP_WARD_D><<<nprob, 768, tsm, c->stream>>>(probs_dev, method);
                else hclust_tri_kernel<768, 1, 0><<<nprob, 768, tsm, c->stream>>>(probs_dev, method);
            } else if (method == SHARP_WARD_D) {
                SHARP_SMEM_OPTIN_ONCE((hclust_tri_kernel<512, 2, SHARP_WARD_D>), c->device);
                hclust_tri_kernel<512, 2, SHARP_WARD_D><<<nprob, 512, tsm, c->stream>>>(probs_dev, method);
            } else {
                SHARP_SMEM_OPTIN_ONCE((hclust_tri_kernel<512, 2, 0>), c->device);
                hclust_tri_kernel<512, 2, 0><<<nprob, 512, tsm, c->stream>>>(probs_dev, method);
            }
        } else {
            SHARP_SMEM_OPTIN_ONCE((hclust_rnn_kernel), c->device);
            hclust_rnn_kernel<<<nprob, RNN_THREADS, hclust_rnn_smem_bytes(max_n), c->stream>>>(probs_dev, method);
        }
        prof_end(c);
        SHARP_CUDA(cudaGetLastError());
        return launch_exact(c, probs_dev, nprob, max_n, method, 1);
    }
    return launch_exact(c, probs_dev, nprob, max_n, method, 0);
}

}  // namespace sharp
