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
// order is recovered at the end by sorting the merges by height.  On a 2000-cell block this takes ~40 rounds instead
// of 1999 steps; a round is wide (hundreds of independent Lance-Williams row updates and row scans), so the kernel is
// bound by memory throughput instead of by a chain of dependent round trips.
//   * The merges and therefore the labels are those of hclust.f.  The heights agree to rounding only: a distance
//     between two clusters that were both built after the previous global step is reached through the same
//     Lance-Williams recurrences associated in a different order (a few ulps).
//   * Everything above assumes there are no exact ties.  Any tie that could matter -- a row minimum attained twice, an
//     updated distance that is not strictly above the row's nearest-neighbour distance, two merges of equal height,
//     no reciprocal pair at all (NaN / Inf) -- sets P.fallback and the problem is redone by the exact kernel
//     (launch_hclust runs hclust_kernel in only_fallback mode right after), so tie-heavy input costs time, not parity.
constexpr int RNN_THREADS = 512;
constexpr int RNN_U = 4;

__device__ __forceinline__ DI rnn_scan_row(const double *row, const unsigned char *flag, int n, int self, int lane, bool *tie) {
    DI best;
    best.d = SHARP_INF;
    best.i = INT_MAX;
    bool t = false;
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
            if (j < n && j != self && flag[j]) {
                if (v[u] < best.d) { best.d = v[u]; best.i = j; t = false; }
                else if (v[u] == best.d) t = true;
            }
        }
    }
    const DI w = warp_argmin_redux(best);
    const bool mine = best.i != INT_MAX && best.d == w.d && (best.i != w.i || t);
    *tie = __any_sync(0xffffffffu, mine);
    return w;
}

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

__global__ void __launch_bounds__(RNN_THREADS, 2) hclust_rnn_kernel(HcProb *probs, int method) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_np, s_cnt, s_tie, s_nm;
    constexpr int THREADS = RNN_THREADS, NW = RNN_THREADS / 32;

    HcProb &P = probs[blockIdx.x];
    const int n = P.n;
    if (n < 2) return;
    const int ld = P.ld;
    double *D = P.Dw; /* read and written by the whole CTA: no __restrict__, no ld.global.nc */
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int hp = n / 2 + 1;

    double *dnn = reinterpret_cast<double *>(smem_raw);  // [n] distance to the nearest neighbour (all j != i)
    double *ph = dnn + n;                                 // [hp] heights of this round's pairs
    int *nn = reinterpret_cast<int *>(ph + hp);           // [n]
    int *size = nn + n;                                   // [n]
    int *list = size + n;                                 // [n] rows to rescan
    int *mrg = list + n;                                  // [n] 0, +(q+1): kept representative of pair q, -(q+1): retired
    int *pa = mrg + n;                                    // [hp]
    int *pb = pa + hp;                                    // [hp]
    unsigned char *flag = reinterpret_cast<unsigned char *>(pb + hp);  // [n]

    if (method == SHARP_WARD_D2) { /* hclust.f: iOpt 8 works on squared dissimilarities */
        for (size_t idx = tid; idx < (size_t)n * ld; idx += THREADS) D[idx] = __dmul_rn(D[idx], D[idx]);
    }
    for (int i = tid; i < n; i += THREADS) {
        flag[i] = 1;
        size[i] = 1;
        mrg[i] = 0;
    }
    if (tid == 0) { s_tie = 0; s_nm = 0; P.fallback = 0; }
    __syncthreads();
    for (int i = warp; i < n; i += NW) {
        bool tie;
        const DI b = rnn_scan_row(D + (size_t)i * ld, flag, n, i, lane, &tie);
        if (lane == 0) {
            nn[i] = (b.i == INT_MAX) ? -1 : b.i;
            dnn[i] = b.d;
            if (tie || b.i == INT_MAX) s_tie = 1;
        }
    }
    __syncthreads();

    int nact = n;
    bool failed = false;
    while (nact > 1) {
        if (s_tie) { failed = true; break; }
        if (tid == 0) s_np = 0;
        __syncthreads();
        // ---- reciprocal nearest neighbours ----
        for (int i = tid; i < n; i += THREADS) {
            if (!flag[i]) continue;
            const int j = nn[i];
            if (j > i && nn[j] == i) {
                const int q = atomicAdd(&s_np, 1);
                pa[q] = i;
                pb[q] = j;
                ph[q] = dnn[i];
            }
        }
        __syncthreads();
        const int np = s_np;
        if (np == 0) { failed = true; break; }
        const int base = s_nm;
        for (int q = tid; q < np; q += THREADS) {
            const int a = pa[q], b = pb[q];
            mrg[a] = q + 1;
            mrg[b] = -(q + 1);
            P.ia[base + q] = a + 1;
            P.ib[base + q] = b + 1;
            P.crit[base + q] = ph[q];
        }
        __syncthreads();
        // ---- Lance-Williams: merged cluster (kept under a) against every cluster that does not merge this round ----
        for (int q = warp; q < np; q += NW) {
            const int a = pa[q], b = pb[q];
            const double h = ph[q], mi = (double)size[a], mj = (double)size[b];
            double *rowa = D + (size_t)a * ld;
            const double *rowb = D + (size_t)b * ld;
            bool bad = false;
            for (int kb = 0; kb < n; kb += 32 * RNN_U) {
                double x[RNN_U], y[RNN_U];
                bool act[RNN_U];
#pragma unroll
                for (int u = 0; u < RNN_U; u++) {
                    const int k = kb + u * 32 + lane;
                    act[u] = k < n && flag[k] && mrg[k] == 0;
                    x[u] = act[u] ? rowa[k] : 0.0;
                    y[u] = act[u] ? rowb[k] : 0.0;
                }
#pragma unroll
                for (int u = 0; u < RNN_U; u++) {
                    if (!act[u]) continue;
                    const int k = kb + u * 32 + lane;
                    const double r = lance_williams(method, x[u], y[u], h, mi, mj, (double)size[k]);
                    rowa[k] = r;
                    D[(size_t)k * ld + a] = r;
                    /* a row whose nearest neighbour survives keeps it only if the new distance is strictly larger */
                    if (!(r > dnn[k]) && mrg[nn[k]] == 0) bad = true;
                }
            }
            if (bad) s_tie = 1;
        }
        // ---- two clusters that both merge this round: the two updates in the order of their heights ----
        for (int idx = tid; idx < np * np; idx += THREADS) {
            const int p1 = idx / np, q1 = idx - p1 * np;
            if (p1 >= q1) continue;
            int f = p1, g = q1; /* f merges first */
            if (ph[q1] < ph[p1]) { f = q1; g = p1; }
            if (ph[p1] == ph[q1]) s_tie = 1;
            const int a = pa[f], b = pb[f], c = pa[g], d = pb[g];
            const double ma = (double)size[a], mb = (double)size[b], mc = (double)size[c], md = (double)size[d];
            const double tc = lance_williams(method, D[(size_t)a * ld + c], D[(size_t)b * ld + c], ph[f], ma, mb, mc);
            const double td = lance_williams(method, D[(size_t)a * ld + d], D[(size_t)b * ld + d], ph[f], ma, mb, md);
            const double r = lance_williams(method, tc, td, ph[g], mc, md, ma + mb);
            D[(size_t)a * ld + c] = r;
            D[(size_t)c * ld + a] = r;
        }
        __syncthreads();
        for (int q = tid; q < np; q += THREADS) {
            const int a = pa[q], b = pb[q];
            size[a] += size[b];
            flag[b] = 0;
        }
        if (tid == 0) { s_nm = base + np; s_cnt = 0; }
        __syncthreads();
        nact -= np;
        if (nact > 1) {
            // ---- rows whose nearest neighbour merged, and the merged rows themselves ----
            for (int k = tid; k < n; k += THREADS)
                if (flag[k] && (mrg[k] > 0 || mrg[nn[k]] != 0)) list[atomicAdd(&s_cnt, 1)] = k;
            __syncthreads();
            const int cnt = s_cnt;
            for (int r = warp; r < cnt; r += NW) {
                const int i = list[r];
                bool tie;
                const DI b = rnn_scan_row(D + (size_t)i * ld, flag, n, i, lane, &tie);
                if (lane == 0) {
                    nn[i] = (b.i == INT_MAX) ? -1 : b.i;
                    dnn[i] = b.d;
                    if (tie || b.i == INT_MAX) s_tie = 1;
                }
            }
            __syncthreads();
        }
        for (int q = tid; q < np; q += THREADS) {
            mrg[pa[q]] = 0;
            mrg[pb[q]] = 0;
        }
        __syncthreads();
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
    const size_t hp = (size_t)n / 2 + 1;
    size_t rounds = (size_t)n * 8 + hp * 8 + (size_t)n * 16 + hp * 8 + (size_t)n;
    size_t p2 = 1;
    while ((int)p2 < n - 1) p2 <<= 1;
    size_t sort = p2 * 12 + (size_t)n * 8;
    return (std::max(rounds, sort) + 31) & ~(size_t)15;
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
        SHARP_CUDA(cudaFuncSetAttribute(hclust_kernel<256, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SHARP_SMEM_OPTIN));
        hclust_kernel<256, 2><<<nprob, 256, smem, c->stream>>>(probs_dev, method, only_fallback);
    } else {
        SHARP_CUDA(cudaFuncSetAttribute(hclust_kernel<128, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, SHARP_SMEM_OPTIN));
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
    static const bool no_rnn = getenv("SHARP_HCLUST_EXACT") != nullptr; /* development switch */
    const bool reducible = method != SHARP_MEDIAN && method != SHARP_CENTROID;
    const size_t rsmem = hclust_rnn_smem_bytes(max_n);
    if (fast && reducible && !no_rnn && max_n > 384 && rsmem <= (size_t)SHARP_SMEM_OPTIN) {
        SHARP_CUDA(cudaFuncSetAttribute(hclust_rnn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SHARP_SMEM_OPTIN));
        prof_begin(c, KID_HCLUST);
        hclust_rnn_kernel<<<nprob, RNN_THREADS, rsmem, c->stream>>>(probs_dev, method);
        prof_end(c);
        SHARP_CUDA(cudaGetLastError());
        return launch_exact(c, probs_dev, nprob, max_n, method, 1);
    }
    return launch_exact(c, probs_dev, nprob, max_n, method, 0);
}

}  // namespace sharp
