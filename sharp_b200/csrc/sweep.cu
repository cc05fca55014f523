// sweep.cu -- K4/K5/K6 + the selection rule of get_opt_hclust (R/get_opt_hclust.R:111-229).
//
// For every candidate number of clusters k in minN..min(maxN, n-1) the reference calls cutree, silhouette
// (+ median) and clues::get_CH -- 39 passes over the distance matrix per 2000-cell block.  Two kernels:
//
//  * sweep_nested_kernel (feature problems, the 2000-cell blocks): the cuts are nested, so the per-point
//    distance sums diC[x][c] are accumulated ONCE for the finest cut (one coalesced pass over D, summed in
//    ascending y like cluster::sildist) and every coarser level only adds two columns, following the dendrogram
//    (so a cluster that is unchanged between two levels keeps bit-identical sums and the reference's exact
//    ties msil[k] == msil[k+1] survive).  The CH index comes from the Gram matrix of the finest-level cluster
//    sums.  One CTA per problem.
//  * sweep_exact_kernel (similarity problems of wMetaC / sMetaC, n <= ~1000, full of exact ties): one CTA per
//    (problem, level); every sum runs in the reference's order, so msil is bit-identical to the CPU oracle.
//
// Both are followed by the selection rule (sil -> CH -> height gap) and the final cutree.
#include "devutil.cuh"
#include "internal.cuh"

namespace sharp {

constexpr int SW_THREADS = 256;
constexpr int NESTED_MAXK = 64;
constexpr int SIL_U = 16;

// ---- dendrogram helpers --------------------------------------------------------------------------
// hclust.f keeps the merged cluster under the smaller representative I2 and retires J2, so the representative
// of a cluster is always its minimum index.  step_of[j] = merge step at which j was retired (INT_MAX if never),
// link[j] = the representative it was merged into.
// IT = int, or unsigned short (n < 65535) where shared memory is tight; "never" is the largest value of IT.
template <class IT>
__device__ __forceinline__ void build_links(int n, const int *ia, const int *ib, IT *step_of, IT *link) {
    const IT never = sizeof(IT) == 2 ? (IT)0xffff : (IT)INT_MAX;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        step_of[i] = never;
        link[i] = (IT)i;
    }
    __syncthreads();
    for (int s = threadIdx.x; s < n - 1; s += blockDim.x) {
        int j2 = ib[s] - 1;
        step_of[j2] = (IT)s;
        link[j2] = (IT)(ia[s] - 1);
    }
    __syncthreads();
}

template <class IT>
__device__ __forceinline__ int root_at(int x, int nmerge, const IT *step_of, const IT *link) {
    while ((int)step_of[x] < nmerge) x = link[x];
    return x;
}

// cutree(h, k): labels numbered by first appearance = rank of the cluster's representative (its minimum index)
// among the live representatives.  rank[] (shared, n ints) is scratch; lab[] receives 0-based ids.
template <int THREADS, class IT>
__device__ __forceinline__ void labels_at(int n, int k, const IT *step_of, const IT *link, int *rank, int *tmp,
                                          int *lab) {
    const int nmerge = n - k;
    for (int i = threadIdx.x; i < n; i += THREADS) rank[i] = ((int)step_of[i] >= nmerge) ? 1 : 0;
    __syncthreads();
    block_exclusive_scan<THREADS>(rank, n, tmp);
    for (int i = threadIdx.x; i < n; i += THREADS) lab[i] = rank[root_at(i, nmerge, step_of, link)];
    __syncthreads();
}

// ---- selection rule (R/get_opt_hclust.R:162-217), run by one thread ----------------------------------
// Returns the 1-based column `oind` of v (level kmin + oind - 1) or a negative status.
__device__ int select_rule(int n, int nlev, const double *msil, const double *chind, const double *height,
                           double sil_thre, double height_ntimes, double *maxsil_out) {
    double mx = msil[0];
    bool anynan = (msil[0] != msil[0]);
    for (int i = 1; i < nlev; i++) {
        if (msil[i] != msil[i]) anynan = true;
        if (msil[i] > mx) mx = msil[i];
    }
    if (anynan) return -SWEEP_E_NAN;
    int ntie = 0;
    for (int i = 0; i < nlev; i++) ntie += (msil[i] == mx);
    int want = (ntie + 1) / 2; /* tmp[ceiling(length(tmp)/2)] */
    int oind = 0;
    for (int i = 0, seen = 0; i < nlev; i++)
        if (msil[i] == mx && ++seen == want) { oind = i + 1; break; }
    *maxsil_out = mx;
    if (mx <= sil_thre) {
        int best = -1; /* which.max(CHind): first maximum, NaN skipped */
        for (int i = 0; i < nlev; i++)
            if (chind[i] == chind[i] && (best < 0 || chind[i] > chind[best])) best = i;
        if (best < 0) return -SWEEP_E_CHNAN;
        oind = best + 1;
        if (oind == 1) {
            const int nh = n - 1;
            const int nt = nh < 10 ? nh : 10;
            const double *t = height + (nh - nt); /* tail(h$height, 10) */
            int pind = -1;
            for (int i = 0; i + 1 < nt; i++) {
                double dif = __dsub_rn(t[i + 1], t[i]);
                if (dif > __dmul_rn(height_ntimes - 1.0, t[i])) { pind = i; break; }
            }
            if (pind >= 0) {
                double opth = __ddiv_rn(__dadd_rn(t[pind], t[pind + 1]), 2.0);
                for (int i = 0; i + 1 < nh; i++)
                    if (height[i + 1] < height[i]) return -SWEEP_E_UNSORTED;
                int idx = nh;
                for (int i = 0; i < nh; i++)
                    if (height[i] > opth) { idx = i; break; }
                int kcut = n + 1 - (idx + 1);
                oind = kcut - 1; /* used as a COLUMN index by the reference (quirk B2) */
            }
        }
    }
    if (oind < 1 || oind > nlev) return -SWEEP_E_OIND;
    return oind;
}

// =====================================================================================================
// nested sweep (feature problems)
// =====================================================================================================
// One CTA of NT threads per problem; thread t owns point x = c0 + t of the current chunk of NT points.
//   * the rows are visited CLUSTER BY CLUSTER of the finest cut (counting sort `order`, ascending y inside a cluster --
//     the order cluster::sildist adds in), so the distance sum of (x, cluster) is a register accumulation, stored once
//     per cluster: no shared-memory read-modify-write per element (r2 profile: 30 % of the stalls of the first version);
//   * shared memory holds the MEANS q[c][t] = sum / count of every live cluster (dead clusters: +Inf), the sums live in
//     an L2-resident scratch: a level touches two sums, and a rescan for the nearest other cluster is NT-uniform loads
//     and compares instead of kmax fp64 divisions in a handful of divergent lanes (23 % of the stalls);
//   * the medians are radix selections (one warp per level) instead of 39 block-wide bitonic sorts (18 %).
// dynamic shared memory layout (bytes):
//   q        [kcap][NT] double   (also: Gram matrix kcap*kcap doubles; the scan scratch rank [n] int of cutree, which
//            runs before and after the passes that use q; the radix histograms [NT/32][256] int)
//   cidf [n] int, step_of [n] u16, link [n] u16, order [n] u16
//   small tables: cnt_lv [nlevcap][kcap] int, own map slot_lv [nlevcap][kcap] uint8, mergeA/B [nlevcap] int, cstart
struct NestedLayout {
    size_t acc_off, ints_off, cnt_off, slot_off, merge_off, msil_off, cstart_off, total;
};
__host__ __device__ inline NestedLayout nested_layout(int n, int kcap, int nlevcap, int nt) {
    NestedLayout L;
    size_t acc_bytes = (size_t)kcap * nt * 8;
    size_t gram_bytes = (size_t)kcap * kcap * 8;
    size_t hist_bytes = (size_t)(nt / 32) * 256 * 4;
    size_t a = acc_bytes > gram_bytes ? acc_bytes : gram_bytes;
    a = a > hist_bytes ? a : hist_bytes;
    a = a > (size_t)n * 4 ? a : (size_t)n * 4;
    L.acc_off = 0;
    L.ints_off = (a + 15) & ~(size_t)15;
    L.cnt_off = (L.ints_off + (size_t)n * 4 + (size_t)3 * n * 2 + 15) & ~(size_t)15;
    L.slot_off = L.cnt_off + (size_t)nlevcap * kcap * 4;
    L.merge_off = (L.slot_off + (size_t)nlevcap * kcap + 15) & ~(size_t)15;
    L.msil_off = (L.merge_off + (size_t)nlevcap * 2 * 4 + 15) & ~(size_t)15;
    L.cstart_off = L.msil_off + (size_t)nlevcap * 2 * 8;
    L.total = L.cstart_off + (size_t)(kcap + 1) * 4;
    return L;
}

constexpr int NESTED_NT_MAX = 512;
constexpr size_t NESTED_SMEM_MAX = 216 * 1024;

static int nested_threads(int max_n, int kcap, int nlevcap) {
    return nested_layout(max_n, kcap, nlevcap, 512).total <= NESTED_SMEM_MAX ? 512 : 256;
}

size_t sweep_nested_scratch_bytes(int max_n, int max_p, HcParamsDev prm) {
    int nlev = prm.n_cluster ? 1 : (prm.max_n - prm.min_n + 1);
    if (nlev < 1) nlev = 1;
    size_t sil = (size_t)nlev * max_n * 8;
    size_t csum = (size_t)NESTED_MAXK * max_p * 8;
    size_t sums = (size_t)NESTED_MAXK * NESTED_NT_MAX * 8;
    return ((sil + csum + sums) + 255) & ~(size_t)255;
}

// order-preserving 64-bit key of a double (NaN excluded by the caller), and back
__device__ __forceinline__ unsigned long long sil_key(double v) {
    const long long b = __double_as_longlong(v + 0.0); /* -0 -> +0 */
    return b < 0 ? ~(unsigned long long)b : ((unsigned long long)b | 0x8000000000000000ull);
}
__device__ __forceinline__ double sil_unkey(unsigned long long k) {
    return __longlong_as_double((long long)((k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k));
}

// stats::median of a[0..n) by one warp: radix selection of the element of rank (n + 1) / 2 - 1, eight 8-bit passes over
// the (L2-resident) values with a 256-bin histogram in shared memory; for even n the next element is either the same
// value or the smallest larger one.  Returns the same bits as median_sorted() on the sorted array.
__device__ double warp_median_select(const double *a, int n, int *hist, int lane) /* a[] was written by this block: plain loads, no read-only path */ {
    const int r0 = (n + 1) / 2 - 1;
    int r = r0, less = 0, eq = 0;
    unsigned long long prefix = 0ull, mask = 0ull;
    bool nan = false;
    for (int b = 7; b >= 0; b--) {
        for (int i = lane; i < 256; i += 32) hist[i] = 0;
        __syncwarp();
        for (int i0 = lane; i0 < n; i0 += 32 * 8) { /* eight loads in flight per lane: the values live in L2, not in registers */
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) v[u] = (i0 + 32 * u < n) ? a[i0 + 32 * u] : 0.0;
#pragma unroll
            for (int u = 0; u < 8; u++) {
                if (i0 + 32 * u >= n) continue;
                if (v[u] != v[u]) { nan = true; continue; }
                const unsigned long long k = sil_key(v[u]);
                if ((k & mask) == prefix) atomicAdd(&hist[(int)((k >> (8 * b)) & 255ull)], 1);
            }
        }
        __syncwarp();
        if (__any_sync(0xffffffffu, nan)) return __longlong_as_double(0x7ff8000000000000ll);
        int h[8], mine = 0;
#pragma unroll
        for (int u = 0; u < 8; u++) { h[u] = hist[lane * 8 + u]; mine += h[u]; }
        int inc = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        const int exc = inc - mine;
        const bool here = r >= exc && r < inc;
        int bin = 0, before = 0, cnt = 0;
        if (here) {
            int run = exc;
#pragma unroll
            for (int u = 0; u < 8; u++) {
                if (r >= run && r < run + h[u]) { bin = lane * 8 + u; before = run; cnt = h[u]; }
                run += h[u];
            }
        }
        const int src = __ffs(__ballot_sync(0xffffffffu, here)) - 1;
        bin = __shfl_sync(0xffffffffu, bin, src);
        before = __shfl_sync(0xffffffffu, before, src);
        cnt = __shfl_sync(0xffffffffu, cnt, src);
        less += before;
        r -= before;
        eq = cnt;
        prefix |= (unsigned long long)bin << (8 * b);
        mask |= 255ull << (8 * b);
        __syncwarp();
    }
    const double x = sil_unkey(prefix);
    if (n & 1) return x;
    double y = x;
    if (less + eq <= r0 + 1) { /* the next element in sorted order is the smallest value above x */
        unsigned long long best = ~0ull;
        for (int i0 = lane; i0 < n; i0 += 32 * 8) {
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) v[u] = (i0 + 32 * u < n) ? a[i0 + 32 * u] : 0.0;
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const unsigned long long k = sil_key(v[u]);
                if (i0 + 32 * u < n && k > prefix && k < best) best = k;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long t = __shfl_xor_sync(0xffffffffu, best, o);
            best = t < best ? t : best;
        }
        y = sil_unkey(best);
    }
    double s = __ddiv_rn(__dadd_rn(x, y), 2.0); /* mean(c(x, y)) like median_sorted */
    const double t = __dadd_rn(__dsub_rn(x, s), __dsub_rn(y, s));
    s = __dadd_rn(s, __ddiv_rn(t, 2.0));
    return s;
}

template <int NT>
__global__ void __launch_bounds__(NT, 1)
sweep_nested_kernel(HcProb *probs, SweepOut *outs, HcParamsDev prm, int n_cap, int kcap, int nlevcap, double *scratch,
                    size_t scratch_per_prob) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int tmp_scan[NT];
    __shared__ int s_oind;
    constexpr int NW = NT / 32;

    HcProb &P = probs[blockIdx.x];
    SweepOut &O = outs[blockIdx.x];
    const int n = P.n;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (n <= 0) return;
    if (P.status != 0) { if (tid == 0) O.meta[3] = P.status; return; }

    // levels
    int kmin, kmax;
    if (prm.n_cluster) { kmin = kmax = prm.n_cluster; }
    else { kmin = prm.min_n; kmax = min(prm.max_n, n - 1); }
    if (kmax < kmin || kmax > n - 1 || kmin < 2) {
        /* R: silhouette() returns NA when k >= n or nc runs backwards; cutree stops for k > n */
        if (tid == 0) { O.meta[0] = 0; O.meta[3] = (prm.n_cluster && kmax > n) ? SWEEP_E_KRANGE : SWEEP_E_FEWPOINTS; }
        return;
    }
    const int nlev = kmax - kmin + 1;
    const int ld = P.ld;
    const double *__restrict__ D = P.D;

    NestedLayout L = nested_layout(n_cap, kcap, nlevcap, NT);
    double *q = reinterpret_cast<double *>(smem + L.acc_off);         // [kmax][NT] mean distance to every live cluster
    int *cidf = reinterpret_cast<int *>(smem + L.ints_off);
    unsigned short *step_of = reinterpret_cast<unsigned short *>(cidf + n_cap);
    unsigned short *link = step_of + n_cap;
    unsigned short *order = link + n_cap;                             // rows by (finest cluster, index)
    int *rank = reinterpret_cast<int *>(q);
    int *cnt_lv = reinterpret_cast<int *>(smem + L.cnt_off);          // [nlev][kcap], level 0 = finest (k = kmax)
    unsigned char *slot_lv = smem + L.slot_off;                       // [nlev][kcap]: fine id -> slot at level
    int *mergeA = reinterpret_cast<int *>(smem + L.merge_off);        // [nlev] slot kept when going to level l
    int *mergeB = mergeA + nlevcap;                                   // [nlev] slot absorbed
    double *msil = reinterpret_cast<double *>(smem + L.msil_off);     // [nlev] indexed by level (0 = finest)
    double *chv = msil + nlevcap;
    int *cstart = reinterpret_cast<int *>(smem + L.cstart_off);       // [kmax + 1] first position of a cluster in `order`

    build_links(n, P.ia, P.ib, step_of, link);
    // finest cut
    labels_at<NT>(n, kmax, step_of, link, rank, tmp_scan, cidf);
    for (int i = tid; i < nlevcap * kcap; i += NT) cnt_lv[i] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += NT) atomicAdd(&cnt_lv[cidf[i]], 1);
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        for (int c = 0; c < kmax; c++) { cstart[c] = run; run += cnt_lv[c]; }
        cstart[kmax] = run;
        for (int c = 0; c < kmax; c++) slot_lv[c] = (unsigned char)c;
        mergeA[0] = mergeB[0] = -1;
        for (int l = 1; l < nlev; l++) {
            /* level l has k = kmax - l clusters: apply merge step s = n - (kmax - l) - 1 (0-based) */
            int s = n - (kmax - l) - 1;
            int a = cidf[P.ia[s] - 1], b = cidf[P.ib[s] - 1]; /* fine ids of the two representatives */
            a = slot_lv[(l - 1) * kcap + a];
            b = slot_lv[(l - 1) * kcap + b];
            mergeA[l] = a;
            mergeB[l] = b;
            for (int c = 0; c < kmax; c++) {
                unsigned char sl = slot_lv[(l - 1) * kcap + c];
                slot_lv[l * kcap + c] = (sl == b) ? (unsigned char)a : sl;
                cnt_lv[l * kcap + c] = cnt_lv[(l - 1) * kcap + c];
            }
            cnt_lv[l * kcap + a] += cnt_lv[l * kcap + b];
            cnt_lv[l * kcap + b] = 0;
        }
    }
    __syncthreads();
    // counting sort of the rows by finest cluster, ascending index inside a cluster: one warp per cluster
    for (int c = warp; c < kmax; c += NW) {
        int pos = cstart[c];
        for (int y0 = 0; y0 < n; y0 += 32) {
            const int y = y0 + lane;
            const bool in = y < n && cidf[y] == c;
            const unsigned bal = __ballot_sync(0xffffffffu, in);
            if (in) order[pos + __popc(bal & ((1u << lane) - 1u))] = (unsigned short)y;
            pos += __popc(bal);
        }
    }
    __syncthreads();

    double *silg = scratch + blockIdx.x * (scratch_per_prob / 8);   // [nlev][n]
    double *csum = silg + (size_t)nlev * n;                          // [kmax][p]
    double *sums = csum + (size_t)NESTED_MAXK * P.p;                 // [kmax][NT] distance sums of the current chunk

    // ---- silhouettes: one pass over D per chunk of NT points ----
    for (int c0 = 0; c0 < n; c0 += NT) {
        const int x = c0 + tid;
        const bool valid = x < n;
        const double *col = D + (valid ? x : 0);
        {
            /* SIL_U rows in flight per thread; the adds of a cluster stay in ascending y like cluster::sildist */
            int c = 0, end = cstart[1];
            double acc = 0.0;
            for (int i0 = 0; i0 < n; i0 += SIL_U) {
                double d[SIL_U];
#pragma unroll
                for (int u = 0; u < SIL_U; u++) d[u] = (i0 + u < n) ? col[(size_t)order[i0 + u] * ld] : 0.0;
#pragma unroll
                for (int u = 0; u < SIL_U; u++) {
                    if (i0 + u == end && i0 + u < n) { /* uniform over the block: the cluster is complete */
                        sums[c * NT + tid] = acc;
                        q[c * NT + tid] = __ddiv_rn(acc, (double)(end - cstart[c]));
                        acc = 0.0;
                        c++;
                        end = cstart[c + 1];
                    }
                    if (i0 + u < n) acc += d[u];
                }
            }
            sums[c * NT + tid] = acc;
            q[c * NT + tid] = __ddiv_rn(acc, (double)(end - cstart[c]));
        }
        const int myfine = valid ? cidf[x] : 0;
        /* b_i = min over the OTHER clusters of the mean distance.  Going one level up merges two clusters (B into A) and
           leaves every other mean untouched, so the minimum is carried along: one division for the new cluster, and a
           rescan of the cached means only when the cluster that held the minimum is one of the two (the quotients that
           are compared are the same ones a full scan computes, so b_i is bit-identical to the scan's) */
        double a_i = 0.0, b_i = SHARP_INF;
        int b_arg = -1, own = 0, n_own = 0;
        for (int l = 0; l < nlev; l++) {
            const int *cnt = cnt_lv + l * kcap;
            bool rescan = (l == 0);
            if (l > 0) {
                const int A = mergeA[l], B = mergeB[l];
                const double sA = sums[A * NT + tid] + sums[B * NT + tid];
                sums[A * NT + tid] = sA;
                const double qA = __ddiv_rn(sA, (double)cnt[A]);
                q[A * NT + tid] = qA;
                q[B * NT + tid] = SHARP_INF;
                if (own == A || own == B) { /* the own cluster grows; the other of the two stops being a candidate */
                    const int other = (own == A) ? B : A;
                    own = A;
                    n_own = cnt[A];
                    a_i = (n_own > 1) ? __ddiv_rn(sA, (double)(n_own - 1)) : 0.0;
                    if (b_arg == other) rescan = true;
                } else if (b_arg == A || b_arg == B) rescan = true;
                else if (qA < b_i) { b_i = qA; b_arg = A; }
            } else {
                own = slot_lv[myfine];
                n_own = cnt[own];
                a_i = (n_own > 1) ? __ddiv_rn(sums[own * NT + tid], (double)(n_own - 1)) : 0.0;
            }
            if (rescan) {
                b_i = SHARP_INF;
                b_arg = -1;
                for (int c = 0; c < kmax; c++) {
                    const double v = q[c * NT + tid];
                    if (c != own && v < b_i) { b_i = v; b_arg = c; }
                }
            }
            double s = (n_own > 1 && b_i != a_i) ? __ddiv_rn(__dsub_rn(b_i, a_i), fmax(a_i, b_i)) : 0.0;
            if (valid) silg[(size_t)l * n + x] = s;
        }
    }
    __syncthreads();

    // ---- medians: one warp per level ----
    {
        int *hist = reinterpret_cast<int *>(q) + warp * 256;
        for (int l = warp; l < nlev; l += NW) {
            const double md = warp_median_select(silg + (size_t)l * n, n, hist, lane);
            if (lane == 0) msil[l] = md;
        }
        __syncthreads();
    }

    // ---- CH index from the finest-level cluster sums of the unit rows ----
    {
        const double *__restrict__ Y = P.Y;
        const int p = P.p, ldy = P.ldy;
        for (int d0 = 0; d0 < p; d0 += NT) {
            const int d = d0 + tid;
            if (d < p) { /* SIL_U rows in flight per thread, cluster by cluster (ascending rows inside a cluster) */
                const double *colY = Y + d;
                int c = 0, end = cstart[1];
                double acc = 0.0;
                for (int i0 = 0; i0 < n; i0 += SIL_U) {
                    double yv[SIL_U];
#pragma unroll
                    for (int u = 0; u < SIL_U; u++) yv[u] = (i0 + u < n) ? colY[(size_t)order[i0 + u] * ldy] : 0.0;
#pragma unroll
                    for (int u = 0; u < SIL_U; u++) {
                        if (i0 + u == end && i0 + u < n) {
                            csum[(size_t)c * p + d] = acc;
                            acc = 0.0;
                            c++;
                            end = cstart[c + 1];
                        }
                        if (i0 + u < n) acc += yv[u];
                    }
                }
                csum[(size_t)c * p + d] = acc;
            }
        }
        __syncthreads();
        double *G = q; /* [kmax][kmax] Gram matrix of the cluster sums */
        for (int pr = warp; pr < kmax * kmax; pr += NW) {
            int a = pr / kmax, b = pr % kmax;
            if (b < a) continue;
            double s = 0.0;
            for (int d = lane; d < p; d += 32) s = fma(csum[(size_t)a * p + d], csum[(size_t)b * p + d], s);
            s = warp_sum(s);
            if (lane == 0) { G[a * kmax + b] = s; G[b * kmax + a] = s; }
        }
        __syncthreads();
        if (tid == 0) {
            double gtot = 0.0;
            for (int a = 0; a < kmax; a++)
                for (int b = 0; b < kmax; b++) gtot += G[a * kmax + b];
            /* total sum of squares about the grand mean: sum ||y||^2 - ||sum y||^2 / n; rows are unit vectors: sum ||y||^2 = n */
            const double T = (double)n - gtot / (double)n;
            for (int l = 0; l < nlev; l++) {
                if (l > 0) {
                    int a = mergeA[l], b = mergeB[l];
                    for (int c = 0; c < kmax; c++) {
                        if (c == a || c == b) continue;
                        G[a * kmax + c] += G[b * kmax + c];
                        G[c * kmax + a] = G[a * kmax + c];
                    }
                    G[a * kmax + a] += G[b * kmax + b] + 2.0 * G[a * kmax + b];
                }
                const int *cnt = cnt_lv + l * kcap;
                double B = 0.0;
                for (int c = 0; c < kmax; c++)
                    if (cnt[c] > 0) B += G[c * kmax + c] / (double)cnt[c];
                B -= gtot / (double)n;
                double W = T - B;
                int k = kmax - l;
                chv[l] = (B / (double)(k - 1)) / (W / (double)(n - k));
            }
        }
        __syncthreads();
    }

    // ---- outputs in ascending-k order + selection ----
    if (tid == 0) {
        for (int l = 0; l < nlev; l++) {
            O.msil[nlev - 1 - l] = msil[l];
            O.chind[nlev - 1 - l] = chv[l];
        }
        int oind;
        double mx = 0.0;
        if (prm.n_cluster) { oind = 1; mx = msil[0]; }
        else oind = select_rule(n, nlev, O.msil, O.chind, P.crit, prm.sil_thre, prm.height_ntimes, &mx);
        O.meta[0] = nlev;
        O.meta[4] = kmin;
        O.meta[5] = kmax;
        if (oind < 0) { O.meta[3] = -oind; oind = 0; }
        else O.meta[3] = 0;
        O.meta[2] = oind;
        *O.maxsil = mx;
        s_oind = oind;
    }
    __syncthreads();
    const int oind = s_oind;
    if (oind <= 0) return;
    // final cutree at the chosen level (+ every level when v is requested)
    int *lab = cidf;
    {
        const int k = kmin + oind - 1;
        labels_at<NT>(n, k, step_of, link, rank, tmp_scan, lab);
        for (int i = tid; i < n; i += NT) O.f[i] = lab[i] + 1;
        if (tid == 0) O.meta[1] = k; /* optN.cluster = length(unique(f)) = k */
        __syncthreads();
    }
    if (O.v) {
        for (int l = 0; l < nlev; l++) {
            labels_at<NT>(n, kmin + l, step_of, link, rank, tmp_scan, lab);
            for (int i = tid; i < n; i += NT) O.v[(size_t)i * nlev + l] = lab[i] + 1;
            __syncthreads();
        }
    }
}

int launch_sweep_nested(sharp_ctx *c, HcProb *probs_dev, SweepOut *outs_dev, int nprob, int max_n, int max_p,
                        HcParamsDev prm, double *scratch, size_t scratch_per_prob) {
    if (nprob <= 0) return 0;
    int kcap = prm.n_cluster ? prm.n_cluster : prm.max_n;
    if (kcap > max_n - 1) kcap = max_n - 1;
    if (kcap < 2) kcap = 2;
    if (kcap > NESTED_MAXK)
        return set_error(SHARP_E_LIMIT, "nested sweep supports at most %d clusters per level (got %d)", NESTED_MAXK, kcap);
    int nlevcap = prm.n_cluster ? 1 : (prm.max_n - prm.min_n + 1);
    if (nlevcap < 1) nlevcap = 1;
    const int nt = nested_threads(max_n, kcap, nlevcap);
    NestedLayout L = nested_layout(max_n, kcap, nlevcap, nt);
    if (L.total > NESTED_SMEM_MAX)
        return set_error(SHARP_E_LIMIT, "nested sweep: %d objects need %zu bytes of shared memory", max_n, L.total);
    if (scratch_per_prob < sweep_nested_scratch_bytes(max_n, max_p, prm))
        return set_error(SHARP_E_ARG, "nested sweep: scratch too small");
    prof_begin(c, KID_SWEEP_NESTED);
    if (nt == 512) {
        SHARP_SMEM_OPTIN_ONCE((sweep_nested_kernel<512>), c->device);
        sweep_nested_kernel<512><<<nprob, 512, L.total, c->stream>>>(probs_dev, outs_dev, prm, max_n, kcap, nlevcap, scratch, scratch_per_prob);
    } else {
        SHARP_SMEM_OPTIN_ONCE((sweep_nested_kernel<256>), c->device);
        sweep_nested_kernel<256><<<nprob, 256, L.total, c->stream>>>(probs_dev, outs_dev, prm, max_n, kcap, nlevcap, scratch, scratch_per_prob);
    }
    prof_end(c);
    SHARP_CUDA(cudaGetLastError());
    return 0;
}

// =====================================================================================================
// exact sweep (similarity problems; also selectable for feature problems through sharp_opt_hclust(exact=1))
// =====================================================================================================
// Pre-pass per problem: row-standardised copy of Y for the CH index (clues::get_CH "1-corr").
__global__ void __launch_bounds__(SW_THREADS) exact_prep_kernel(HcProb *probs, double *ystd_all, size_t ystd_per_prob) {
    HcProb &P = probs[blockIdx.y];
    const int n = P.n, p = P.p, ldy = P.ldy;
    if (n <= 0) return;
    double *ys = ystd_all + blockIdx.y * ystd_per_prob;
    for (int i = blockIdx.x * SW_THREADS + threadIdx.x; i < n; i += gridDim.x * SW_THREADS) {
        const double *x = P.Y + (size_t)i * ldy;
        double s = 0.0;
        for (int j = 0; j < p; j++) s = __dadd_rn(s, x[j]);
        double mean = __ddiv_rn(s, (double)p);
        double ss = 0.0;
        for (int j = 0; j < p; j++) {
            double d = __dsub_rn(x[j], mean);
            ss = __dadd_rn(ss, __dmul_rn(d, d));
        }
        double sd = sqrt(__ddiv_rn(ss, (double)(p - 1)));
        for (int j = 0; j < p; j++) ys[(size_t)i * p + j] = (p > 1) ? __ddiv_rn(__dsub_rn(x[j], mean), sd) : x[j];
    }
}

// levels of a problem under its own parameters
__device__ __forceinline__ bool level_range(const HcParamsDev &prm, int n, int *kmin, int *kmax) {
    if (prm.n_cluster) { *kmin = *kmax = prm.n_cluster; }
    else { *kmin = prm.min_n; *kmax = min(prm.max_n, n - 1); }
    return !(*kmax < *kmin || *kmax > n - 1 || *kmin < 2);
}

// grid: (max_levels, nprob).  dynamic smem: 4 int arrays of n_cap + sort buffer of P2(n_cap) doubles + coff
__global__ void __launch_bounds__(SW_THREADS)
sweep_exact_kernel(HcProb *probs, SweepOut *outs, const HcParamsDev *prms, int n_cap, int kcap, const double *ystd_all,
                   size_t ystd_per_prob) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int tmp_scan[SW_THREADS];
    __shared__ double red[SW_THREADS / 32];

    HcProb &P = probs[blockIdx.y];
    SweepOut &O = outs[blockIdx.y];
    const HcParamsDev prm = prms[blockIdx.y];
    const int n = P.n;
    const int tid = threadIdx.x;
    if (n <= 0 || P.status != 0) return;
    int kmin, kmax;
    if (!level_range(prm, n, &kmin, &kmax)) return;
    const int lev = blockIdx.x;
    const int k = kmin + lev;
    if (k > kmax) return;
    const int ld = P.ld;
    const double *__restrict__ D = P.D;

    int p2cap = 1;
    while (p2cap < n_cap) p2cap <<= 1;
    double *sb = reinterpret_cast<double *>(smem);             // [P2] silhouette widths / sort buffer
    int *step_of = reinterpret_cast<int *>(sb + p2cap);        // [n]
    int *link = step_of + n_cap;
    int *rank = link + n_cap;
    int *lab = rank + n_cap;
    int *order = lab + n_cap;                                   // [n] members grouped by cluster, ascending inside
    int *coff = order + n_cap;                                  // [kcap + 1]

    build_links(n, P.ia, P.ib, step_of, link);
    labels_at<SW_THREADS>(n, k, step_of, link, rank, tmp_scan, lab);
    // stable counting sort of the points by label
    for (int c = tid; c <= k; c += SW_THREADS) coff[c] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += SW_THREADS) atomicAdd(&coff[lab[i] + 1], 1);
    __syncthreads();
    if (tid == 0)
        for (int c = 0; c < k; c++) coff[c + 1] += coff[c];
    __syncthreads();
    for (int c = tid; c < k; c += SW_THREADS) {
        int pos = coff[c];
        for (int i = 0; i < n; i++)
            if (lab[i] == c) order[pos++] = i;
    }
    __syncthreads();

    // ---- silhouette widths: diC[x][c] summed over the members of c in ascending order (cluster::sildist) ----
    for (int x = tid; x < n; x += SW_THREADS) {
        const int own = lab[x];
        const int n_own = coff[own + 1] - coff[own];
        double a_i = 0.0, b_i = SHARP_INF;
        for (int c = 0; c < k; c++) {
            double s = 0.0;
            const int q1 = coff[c + 1];
            for (int q = coff[c]; q < q1; q++) s = __dadd_rn(s, D[(size_t)order[q] * ld + x]);
            if (c == own) {
                if (n_own > 1) a_i = __ddiv_rn(s, (double)(n_own - 1));
            } else {
                s = __ddiv_rn(s, (double)(q1 - coff[c]));
                if (s < b_i) b_i = s;
            }
        }
        sb[x] = (n_own > 1 && b_i != a_i) ? __ddiv_rn(__dsub_rn(b_i, a_i), fmax(a_i, b_i)) : 0.0;
    }
    const int P2 = next_pow2(n);
    for (int i = n + tid; i < P2; i += SW_THREADS) sb[i] = SHARP_INF;
    __syncthreads();
    block_bitonic_sort<SW_THREADS>(sb, P2);
    if (tid == 0) O.msil[lev] = median_sorted(sb, n);
    __syncthreads();

    // ---- CH index on the row-standardised data: per-dimension between / within sums, fixed-tree reduction ----
    {
        const double *__restrict__ Y = ystd_all ? ystd_all + blockIdx.y * ystd_per_prob : P.Y;
        const int p = P.p;
        const int ldy = ystd_all ? p : P.ldy;
        double Bsum = 0.0, Wsum = 0.0;
        for (int j = tid; j < p; j += SW_THREADS) {
            double tot = 0.0;
            for (int i = 0; i < n; i++) tot += Y[(size_t)i * ldy + j];
            tot /= (double)n;
            for (int c = 0; c < k; c++) {
                const int q0 = coff[c], q1 = coff[c + 1];
                double s = 0.0;
                for (int q = q0; q < q1; q++) s += Y[(size_t)order[q] * ldy + j];
                const double cen = s / (double)(q1 - q0);
                const double db = cen - tot;
                Bsum += (double)(q1 - q0) * db * db;
                for (int q = q0; q < q1; q++) {
                    double dw = Y[(size_t)order[q] * ldy + j] - cen;
                    Wsum += dw * dw;
                }
            }
        }
        double B = block_sum<SW_THREADS>(Bsum, red);
        double W = block_sum<SW_THREADS>(Wsum, red);
        if (tid == 0) O.chind[lev] = (B / (double)(k - 1)) / (W / (double)(n - k));
    }
}

// selection + final labels for the exact path: one CTA per problem
__global__ void __launch_bounds__(SW_THREADS)
sweep_select_kernel(HcProb *probs, SweepOut *outs, const HcParamsDev *prms, int n_cap) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int tmp_scan[SW_THREADS];
    __shared__ int s_oind;
    HcProb &P = probs[blockIdx.x];
    SweepOut &O = outs[blockIdx.x];
    const HcParamsDev prm = prms[blockIdx.x];
    const int n = P.n;
    const int tid = threadIdx.x;
    if (n <= 0) return;
    if (P.status != 0) { if (tid == 0) { O.meta[3] = P.status; O.meta[0] = 0; } return; }
    int kmin, kmax;
    if (!level_range(prm, n, &kmin, &kmax)) {
        if (tid == 0) { O.meta[0] = 0; O.meta[3] = (prm.n_cluster && kmax > n) ? SWEEP_E_KRANGE : SWEEP_E_FEWPOINTS; }
        return;
    }
    const int nlev = kmax - kmin + 1;
    int *step_of = reinterpret_cast<int *>(smem);
    int *link = step_of + n_cap;
    int *rank = link + n_cap;
    int *lab = rank + n_cap;
    if (tid == 0) {
        int oind;
        double mx = 0.0;
        if (prm.n_cluster) { oind = 1; mx = O.msil[0]; }
        else oind = select_rule(n, nlev, O.msil, O.chind, P.crit, prm.sil_thre, prm.height_ntimes, &mx);
        O.meta[0] = nlev;
        O.meta[4] = kmin;
        O.meta[5] = kmax;
        if (oind < 0) { O.meta[3] = -oind; oind = 0; }
        else O.meta[3] = 0;
        O.meta[2] = oind;
        *O.maxsil = mx;
        s_oind = oind;
    }
    __syncthreads();
    const int oind = s_oind;
    if (oind <= 0) return;
    build_links(n, P.ia, P.ib, step_of, link);
    const int k = kmin + oind - 1;
    labels_at<SW_THREADS>(n, k, step_of, link, rank, tmp_scan, lab);
    for (int i = tid; i < n; i += SW_THREADS) O.f[i] = lab[i] + 1;
    if (tid == 0) O.meta[1] = k;
    __syncthreads();
    if (O.v) {
        for (int l = 0; l < nlev; l++) {
            labels_at<SW_THREADS>(n, kmin + l, step_of, link, rank, tmp_scan, lab);
            for (int i = tid; i < n; i += SW_THREADS) O.v[(size_t)i * nlev + l] = lab[i] + 1;
            __syncthreads();
        }
    }
}

size_t sweep_exact_scratch_bytes(int max_n, int max_p) { return (((size_t)max_n * max_p * 8) + 255) & ~(size_t)255; }

int launch_sweep_exact(sharp_ctx *c, HcProb *probs_dev, SweepOut *outs_dev, int nprob, int max_n, int max_p,
                       const HcParamsDev *prm_dev, int max_levels, int kcap, double *scratch, size_t scratch_per_prob) {
    if (nprob <= 0) return 0;
    if (max_levels < 1) max_levels = 1;
    int p2 = 1;
    while (p2 < max_n) p2 <<= 1;
    size_t smem = (size_t)p2 * 8 + (size_t)5 * max_n * 4 + (size_t)(kcap + 2) * 4;
    smem = (smem + 15) & ~(size_t)15;
    if (smem > 220 * 1024)
        return set_error(SHARP_E_LIMIT, "exact sweep: %d objects need %zu bytes of shared memory", max_n, smem);
    {
        dim3 g((max_n + SW_THREADS - 1) / SW_THREADS, nprob);
        prof_begin(c, KID_SWEEP_EXACT);
        exact_prep_kernel<<<g, SW_THREADS, 0, c->stream>>>(probs_dev, scratch, scratch_per_prob / 8);
        prof_end(c);
    }
    SHARP_SMEM_OPTIN_ONCE((sweep_exact_kernel), c->device);
    dim3 grid(max_levels, nprob);
    prof_begin(c, KID_SWEEP_EXACT);
    sweep_exact_kernel<<<grid, SW_THREADS, smem, c->stream>>>(probs_dev, outs_dev, prm_dev, max_n, kcap, scratch,
                                                              scratch_per_prob / 8);
    prof_end(c);
    size_t smem2 = (size_t)4 * max_n * 4;
    SHARP_SMEM_OPTIN_ONCE((sweep_select_kernel), c->device);
    prof_begin(c, KID_SWEEP_EXACT);
    sweep_select_kernel<<<nprob, SW_THREADS, smem2, c->stream>>>(probs_dev, outs_dev, prm_dev, max_n);
    prof_end(c);
    SHARP_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace sharp
