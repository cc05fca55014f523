// metac.cu -- K7/K8/K9 (wMetaC) and K10 (sMetaC): the meta-clustering that merges ensemble members and blocks.
//
// wMetaC (R/wMetaC.R:15-226), batched over the cell blocks of a SHARP_large call (one grid row / CTA per block):
//   wm_enumerate  : R = unique(paste(nC[,c], "_", c)) -- global cluster ids in first-appearance order,
//                   ascending member lists per cluster                                   (R/wMetaC.R:60-67)
//   wm_weights    : co-association AA = mean_c A_c, point weight w1 = (4/N rowSums(AA(1-AA)) + .01)/1.01,
//                   as label-indicator products evaluated on the fly -- the N x N matrices of getA() are never
//                   materialised                                                        (R/wMetaC.R:24-44, 242-283)
//   wm_similarity : weighted Jaccard S[k,j] = sum w1[k & j] / sum w1[k | j] over all cluster pairs, summed in
//                   the reference's order (so identical clusters give exactly 1)         (R/wMetaC.R:70-77, 299-320)
//   (Ward + exact sweep on d = 1 - S: ward.cu / sweep.cu)
//   wm_vote       : per-cell majority vote with the reference's tie-break (levels of a CHARACTER table are
//                   sorted as strings), one-cluster fallback, unique(finalC)                (R/wMetaC.R:141-161)
//   wm_x0         : the soft indicator x0                                                  (R/wMetaC.R:180-208)
// sMetaC (R/sMetaC.R:17-210):
//   sm_centroids / sm_stats / sm_cor : colMeans per cluster and Pearson correlation of centroid pairs, in the
//                   reference's summation order                                            (R/sMetaC.R:58-85)
//   sm_setup      : the k-range tweak by ncells                                            (R/sMetaC.R:101-119)
//   sm_finish     : "second best if 2 clusters" rule and relabel                           (R/sMetaC.R:139-182)
// All tie-sensitive arithmetic uses explicit round-to-nearest operations (no FMA contraction) so that the CPU
// oracle reproduces it bit for bit.
#include "devutil.cuh"
#include "metac.cuh"

namespace sharp {

constexpr int MT = 256;

// decimal-string order of two positive ints (levels of table() on a character vector)
__device__ __forceinline__ bool str_less(int a, int b) {
    // compare most-significant digits first after aligning: equivalent to strcmp on the decimal strings
    int da[10], db[10], na = 0, nb = 0;
    int x = a;
    do { da[na++] = x % 10; x /= 10; } while (x);
    x = b;
    do { db[nb++] = x % 10; x /= 10; } while (x);
    int ia = na - 1, ib = nb - 1;
    while (ia >= 0 && ib >= 0) {
        if (da[ia] != db[ib]) return da[ia] < db[ib];
        ia--; ib--;
    }
    return na < nb; /* a is a proper prefix of b */
}

// ---------------------------------------------------------------------------------------------------
// wm_enumerate: one CTA per block
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(MT) wm_enumerate_kernel(WmArgs A) {
    __shared__ int first_pos[WM_MAXL + 1];
    __shared__ int rank_of[WM_MAXL + 1];
    __shared__ int ccount[WM_MAXL + 1];
    __shared__ int s_base, s_nk;
    const int t = blockIdx.x, tid = threadIdx.x;
    const int64_t s0 = A.start[t];
    const int N = (int)(A.start[t + 1] - s0);
    int *gid = A.gid + (size_t)A.K * s0;          // [K][N] for this block
    int *members = A.members + (size_t)A.K * s0;  // [K][N] grouped by cluster, ascending inside a cluster
    int *moff = A.moff + (size_t)t * (A.capC + 1);
    int *colof = A.col_of + (size_t)t * A.capC;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int k = 0; k < A.K; k++) {
        const int32_t *lab = A.labels + (size_t)k * A.ncells + s0;
        for (int l = tid; l <= WM_MAXL; l += MT) first_pos[l] = INT_MAX;
        __syncthreads();
        for (int i = tid; i < N; i += MT) {
            int l = lab[i];
            if (l >= 1 && l <= WM_MAXL) atomicMin(&first_pos[l], i);
        }
        __syncthreads();
        if (tid == 0) {
            // order the labels that occur by first appearance (insertion sort over <= WM_MAXL items)
            int order[WM_MAXL];
            int cnt = 0;
            for (int l = 1; l <= WM_MAXL; l++) {
                if (first_pos[l] == INT_MAX) continue;
                int q = cnt++;
                while (q > 0 && first_pos[order[q - 1]] > first_pos[l]) { order[q] = order[q - 1]; q--; }
                order[q] = l;
            }
            for (int q = 0; q < cnt; q++) rank_of[order[q]] = q;
            s_nk = cnt;
        }
        __syncthreads();
        const int base = s_base, nk = s_nk;
        const bool fits = (base + nk <= A.capC);
        for (int c = tid; c < nk; c += MT) ccount[c] = 0;
        __syncthreads();
        for (int i = tid; i < N; i += MT) {
            int l = lab[i];
            int r = (l >= 1 && l <= WM_MAXL) ? rank_of[l] : -1;
            gid[(size_t)k * N + i] = (r >= 0) ? base + r : -1;
            if (r >= 0) atomicAdd(&ccount[r], 1);
        }
        __syncthreads();
        if (tid == 0 && fits) {
            int run = k * N;
            for (int c = 0; c < nk; c++) {
                moff[base + c] = run;
                colof[base + c] = k;
                run += ccount[c];
            }
            moff[base + nk] = run;
        }
        __syncthreads();
        if (fits) {
            for (int c = tid >> 5; c < nk; c += MT / 32) { /* stable: one warp per cluster, ascending ballots */
                int pos = moff[base + c];
                const int lane = tid & 31;
                for (int i0 = 0; i0 < N; i0 += 32) {
                    const int i = i0 + lane;
                    const bool in = i < N && gid[(size_t)k * N + i] == base + c;
                    const unsigned bal = __ballot_sync(0xffffffffu, in);
                    if (in) members[pos + __popc(bal & ((1u << lane) - 1u))] = i;
                    pos += __popc(bal);
                }
            }
        }
        __syncthreads();
        if (tid == 0) s_base = base + nk;
        __syncthreads();
    }
    if (tid == 0) {
        const int allC = s_base;
        A.allc[t] = (allC <= A.capC) ? allC : -allC; /* negative: capacity exceeded */
    }
}

// ---------------------------------------------------------------------------------------------------
// wm_weights: grid (ceil(maxN / MT), T); one thread per cell
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(MT) wm_weights_kernel(WmArgs A, int packed_cap) {
    extern __shared__ unsigned long long wpk[]; /* [packed_cap]: the K <= 8 labels of every cell of the block, one byte each */
    const int t = blockIdx.y;
    const int64_t s0 = A.start[t];
    const int N = (int)(A.start[t + 1] - s0);
    const int i = blockIdx.x * MT + threadIdx.x;
    const int K = A.K;
    __shared__ double term[WM_MAXK + 1]; /* x (1 - x) for x = cnt / K: the K + 1 values an entry of AA can take */
    __shared__ int s_wide;
    if (threadIdx.x <= K) {
        const double x = __ddiv_rn((double)threadIdx.x, (double)K);
        term[threadIdx.x] = __dmul_rn(x, __dsub_rn(1.0, x));
    }
    if (threadIdx.x == 0) s_wide = (K > 8 || N > packed_cap) ? 1 : 0;
    __syncthreads();
    const int wide0 = s_wide;
    __syncthreads();
    if (!wide0) { /* labels 1..256 as bytes: one 64-bit word per cell, compared four bytes at a time */
        for (int j = threadIdx.x; j < N; j += MT) {
            unsigned long long w = 0ull;
            for (int k = 0; k < K; k++) {
                const int l = A.labels[(size_t)k * A.ncells + s0 + j];
                if (l < 1 || l > 256) s_wide = 1;
                w |= (unsigned long long)((unsigned)(l - 1) & 255u) << (8 * k);
            }
            wpk[j] = w;
        }
    }
    __syncthreads();
    if (i >= N) return;
    double rs = 0.0;
    if (!s_wide) {
        const unsigned long long me = wpk[i];
        const unsigned mlo = K >= 4 ? 0xffffffffu : ((1u << (8 * K)) - 1u);
        const unsigned mhi = K <= 4 ? 0u : (K >= 8 ? 0xffffffffu : ((1u << (8 * (K - 4))) - 1u));
        const unsigned melo = (unsigned)me, mehi = (unsigned)(me >> 32);
        for (int j = 0; j < N; j++) {
            const unsigned long long o = wpk[j];
            const int cnt = (__popc(__vcmpeq4((unsigned)o, melo) & mlo) + __popc(__vcmpeq4((unsigned)(o >> 32), mehi) & mhi)) >> 3;
            rs = __dadd_rn(rs, term[cnt]); /* term[0] = +0: the sum is the one over the pairs that share a cluster */
        }
    } else {
        int mine[WM_MAXK];
        for (int k = 0; k < K; k++) mine[k] = A.labels[(size_t)k * A.ncells + s0 + i];
        for (int j = 0; j < N; j++) {
            int cnt = 0;
            for (int k = 0; k < K; k++) cnt += (A.labels[(size_t)k * A.ncells + s0 + j] == mine[k]);
            rs = __dadd_rn(rs, term[cnt]);
        }
    }
    double w0 = __dmul_rn(__ddiv_rn(4.0, (double)N), rs);
    A.w1[s0 + i] = __ddiv_rn(__dadd_rn(w0, 0.01), __dadd_rn(1.0, 0.01));
}

// ---------------------------------------------------------------------------------------------------
// wm_similarity: grid (ceil(capC*capC / MT), T); one thread per ordered pair (a <= b)
// writes S (for the CH index), D = 1 - S and its working copy; fills the HcProb descriptor (thread 0,0)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(MT) wm_similarity_kernel(WmArgs A) {
    const int t = blockIdx.y;
    const int allC = A.allc[t];
    const int capC = A.capC;
    double *S = A.S + (size_t)t * capC * capC;
    double *D = A.D + (size_t)t * capC * capC;
    double *Dw = A.Dw + (size_t)t * capC * capC;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        HcProb &P = A.probs[t];
        P.n = allC > 0 ? allC : 0;
        P.ld = capC;
        P.D = D;
        P.Dw = Dw;
        P.ia = A.ia + (size_t)t * capC;
        P.ib = A.ib + (size_t)t * capC;
        P.crit = A.crit + (size_t)t * capC;
        P.Y = S;
        P.p = allC > 0 ? allC : 0;
        P.ldy = capC;
        P.status = allC > 1 ? 0 : (allC < 0 ? WM_E_TOOMANY : 23);
    }
    if (allC < 2) return;
    const int idx = blockIdx.x * MT + threadIdx.x;
    const int a = idx / allC, b = idx - a * allC;
    if (a >= allC || b < a) return;
    if (a == b) {
        S[(size_t)a * capC + a] = 1.0;
        D[(size_t)a * capC + a] = 0.0;
        Dw[(size_t)a * capC + a] = 0.0;
        return;
    }
    const int64_t s0 = A.start[t];
    const int N = (int)(A.start[t + 1] - s0);
    const int *gid = A.gid + (size_t)A.K * s0;
    const int *members = A.members + (size_t)A.K * s0;
    const int *moff = A.moff + (size_t)t * (capC + 1);
    const int *colof = A.col_of + (size_t)t * capC;
    const double *w1 = A.w1 + s0;
    const int ka = colof[a], kb = colof[b];
    double ss = 0.0;
    if (ka != kb) {
        const int *gb = gid + (size_t)kb * N;
        const int *ga = gid + (size_t)ka * N;
        double si = 0.0;
        int ni = 0;
        for (int q = moff[a]; q < moff[a + 1]; q++) {
            int i = members[q];
            if (gb[i] == b) { si = __dadd_rn(si, w1[i]); ni++; }
        }
        if (ni != 0) {
            double su = 0.0; /* union(a, b) = unique(c(a, b)): all of a, then the cells of b not in a */
            for (int q = moff[a]; q < moff[a + 1]; q++) su = __dadd_rn(su, w1[members[q]]);
            for (int q = moff[b]; q < moff[b + 1]; q++) {
                int i = members[q];
                if (ga[i] != a) su = __dadd_rn(su, w1[i]);
            }
            ss = __ddiv_rn(si, su);
        }
    }
    const double d = __dsub_rn(1.0, ss);
    S[(size_t)a * capC + b] = S[(size_t)b * capC + a] = ss;
    D[(size_t)a * capC + b] = D[(size_t)b * capC + a] = d;
    Dw[(size_t)a * capC + b] = Dw[(size_t)b * capC + a] = d;
}

// ---------------------------------------------------------------------------------------------------
// wm_vote: one CTA per block.  Outputs finalc (meta id per cell), ucount[t], ulist[t][capU] (unique(finalC)
// in first-appearance order), fcode (index into ulist) and the block status.
// ---------------------------------------------------------------------------------------------------
struct Vote {
    int v1, c1, v2, c2, nd; /* top value/count, second value/count, number of distinct values */
};

__device__ __forceinline__ Vote vote_row(const int *d, int K) {
    // names(sort(table(d), decreasing = TRUE)): distinct values in string order, stable sort by count desc
    int vals[WM_MAXK], cnts[WM_MAXK], nd = 0;
    for (int k = 0; k < K; k++) {
        int q = 0;
        while (q < nd && vals[q] != d[k]) q++;
        if (q == nd) { vals[nd] = d[k]; cnts[nd] = 1; nd++; }
        else cnts[q]++;
    }
    Vote r;
    r.nd = nd;
    r.v1 = r.v2 = -1;
    r.c1 = r.c2 = 0;
    // best: max count, ties -> smallest in string order
    for (int q = 0; q < nd; q++)
        if (r.v1 < 0 || cnts[q] > r.c1 || (cnts[q] == r.c1 && str_less(vals[q], r.v1))) { r.v1 = vals[q]; r.c1 = cnts[q]; }
    for (int q = 0; q < nd; q++) {
        if (vals[q] == r.v1) continue;
        if (r.v2 < 0 || cnts[q] > r.c2 || (cnts[q] == r.c2 && str_less(vals[q], r.v2))) { r.v2 = vals[q]; r.c2 = cnts[q]; }
    }
    return r;
}

__global__ void __launch_bounds__(MT) wm_vote_kernel(WmArgs A, const SweepOut *outs) {
    extern __shared__ int vsm[]; /* first_pos[capC + 1] */
    __shared__ int s_multi, s_err;
    const int t = blockIdx.x, tid = threadIdx.x;
    const int64_t s0 = A.start[t];
    const int N = (int)(A.start[t + 1] - s0);
    const int K = A.K;
    const int allC = A.allc[t];
    const SweepOut &O = outs[t];
    int st = (allC < 2) ? (allC < 0 ? WM_E_TOOMANY : 23) : O.meta[3];
    if (st != 0) {
        if (tid == 0) { A.status[t] = st; A.ucount[t] = 0; }
        return;
    }
    const int *tf = O.f; /* [allC] meta-cluster id (1-based) of every cluster */
    const int *gid = A.gid + (size_t)K * s0;
    int *finalc = A.finalc + s0;
    int *first_pos = vsm;
    if (tid == 0) { s_multi = 0; s_err = 0; }
    for (int c = tid; c <= A.capC; c += MT) first_pos[c] = INT_MAX;
    __syncthreads();
    int d[WM_MAXK];
    int v0 = -1;
    {
        /* value of cell 0, to detect N.cluster == 1 */
        for (int k = 0; k < K; k++) d[k] = tf[gid[(size_t)k * N + 0]];
        v0 = vote_row(d, K).v1;
    }
    for (int i = tid; i < N; i += MT) {
        for (int k = 0; k < K; k++) d[k] = tf[gid[(size_t)k * N + i]];
        Vote r = vote_row(d, K);
        finalc[i] = r.v1;
        if (r.v1 != v0) s_multi = 1;
    }
    __syncthreads();
    if (!s_multi) { /* only one cluster: take everybody's second choice (R/wMetaC.R:148-161, quirk B7) */
        for (int i = tid; i < N; i += MT) {
            for (int k = 0; k < K; k++) d[k] = tf[gid[(size_t)k * N + i]];
            Vote r = vote_row(d, K);
            if (r.nd < 2) s_err = 1; /* R: `if (NA >= ...)` -> error */
            else finalc[i] = r.v2;
        }
        __syncthreads();
        if (s_err) {
            if (tid == 0) { A.status[t] = WM_E_ONECLUSTER; A.ucount[t] = 0; }
            return;
        }
    }
    __syncthreads();
    // unique(finalC) in first-appearance order
    for (int i = tid; i < N; i += MT) atomicMin(&first_pos[finalc[i]], i);
    __syncthreads();
    int *ulist = A.ulist + (size_t)t * A.capU;
    if (tid == 0) {
        int cnt = 0;
        int st2 = 0;
        for (int v = 1; v <= A.capC; v++) {
            if (first_pos[v] == INT_MAX) continue;
            if (cnt >= A.capU) { st2 = WM_E_TOOMANY; break; }
            int q = cnt++;
            while (q > 0 && first_pos[ulist[q - 1]] > first_pos[v]) { ulist[q] = ulist[q - 1]; q--; }
            ulist[q] = v;
        }
        A.ucount[t] = st2 ? 0 : cnt;
        A.status[t] = st2;
        /* reuse first_pos[v] as the position of v in ulist */
        if (!st2)
            for (int q = 0; q < cnt; q++) first_pos[ulist[q]] = -(q + 1);
    }
    __syncthreads();
    if (A.status[t] != 0) return;
    int *fcode = A.fcode + s0;
    for (int i = tid; i < N; i += MT) fcode[i] = -first_pos[finalc[i]] - 1;
}

// ---------------------------------------------------------------------------------------------------
// wm_x0: x0 rows.  For cell i of block t and every q in unique(finalC)_t:
//   y0 = #members voting uC[q];  x0 = 1 for the chosen cluster, 0.5 * y0 / y0[chosen] otherwise.
// Output column = colmap[coloff[t] + q] (identity / block offset / sMetaC group) accumulated in ascending q, and
// the row is written at out_row[s0 + i] (un-shuffle).  grid (ceil(maxN/MT), T)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(MT)
wm_x0_kernel(WmArgs A, const SweepOut *outs, const int *coloff, const int *colmap, const int64_t *out_row,
             double *x0, int ncol) {
    const int t = blockIdx.y;
    const int64_t s0 = A.start[t];
    const int N = (int)(A.start[t + 1] - s0);
    const int i = blockIdx.x * MT + threadIdx.x;
    if (i >= N || A.status[t] != 0) return;
    const int K = A.K;
    const int *tf = outs[t].f;
    const int *gid = A.gid + (size_t)K * s0;
    const int *ulist = A.ulist + (size_t)t * A.capU;
    const int nu = A.ucount[t];
    int d[WM_MAXK];
    for (int k = 0; k < K; k++) d[k] = tf[gid[(size_t)k * N + i]];
    const int chosen = A.fcode[s0 + i];
    int ych = 0;
    for (int k = 0; k < K; k++) ych += (d[k] == ulist[chosen]);
    const int64_t row = out_row ? out_row[s0 + i] : (s0 + i);
    double *xr = x0 + (size_t)row * ncol;
    const int off = coloff ? coloff[t] : 0;
    for (int q = 0; q < nu; q++) {
        int y0 = 0;
        for (int k = 0; k < K; k++) y0 += (d[k] == ulist[q]);
        double val = (q == chosen) ? 1.0 : ((y0 != 0) ? __ddiv_rn(__dmul_rn(0.5, (double)y0), (double)ych) : 0.0);
        if (val != 0.0) {
            const int col = colmap ? colmap[off + q] : (off + q);
            xr[col] = __dadd_rn(xr[col], val);
        }
    }
}

int launch_wmetac_front(sharp_ctx *c, const WmArgs &A, int T, int max_block_n) {
    if (A.K > WM_MAXK) return set_error(SHARP_E_LIMIT, "wMetaC: at most %d clustering solutions per cell (got %d)", WM_MAXK, A.K);
    prof_begin(c, KID_WMETAC);
    wm_enumerate_kernel<<<T, MT, 0, c->stream>>>(A);
    prof_end(c);
    dim3 g1((max_block_n + MT - 1) / MT, T);
    prof_begin(c, KID_WM_WEIGHTS);
    const int packed_cap = std::min(max_block_n, 5000); /* 40 KB of dynamic shared memory at most; larger blocks take the plain loop */
    wm_weights_kernel<<<g1, MT, (size_t)packed_cap * 8, c->stream>>>(A, packed_cap);
    prof_end(c);
    dim3 g2(((size_t)A.capC * A.capC + MT - 1) / MT, T);
    prof_begin(c, KID_WM_SIMILARITY);
    wm_similarity_kernel<<<g2, MT, 0, c->stream>>>(A);
    prof_end(c);
    SHARP_CUDA(cudaGetLastError());
    return 0;
}

int launch_wmetac_vote(sharp_ctx *c, const WmArgs &A, const SweepOut *outs, int T) {
    size_t smem = (size_t)(A.capC + 2) * 4;
    prof_begin(c, KID_WMETAC);
    wm_vote_kernel<<<T, MT, smem, c->stream>>>(A, outs);
    prof_end(c);
    SHARP_CUDA(cudaGetLastError());
    return 0;
}

int launch_wmetac_x0(sharp_ctx *c, const WmArgs &A, const SweepOut *outs, int T, int max_block_n, const int *coloff,
                     const int *colmap, const int64_t *out_row, double *x0, int ncol) {
    dim3 g((max_block_n + MT - 1) / MT, T);
    prof_begin(c, KID_WMETAC);
    wm_x0_kernel<<<g, MT, 0, c->stream>>>(A, outs, coloff, colmap, out_row, x0, ncol);
    prof_end(c);
    SHARP_CUDA(cudaGetLastError());
    return 0;
}

// ===================================================================================================
// sMetaC
// ===================================================================================================
// sm_codes (pipeline): global cluster code of every cell = coloff[t] + fcode, cluster member lists (cells are
// block-local, so one CTA per block builds the lists of its own clusters).  coloff[T+1] is the exclusive prefix
// of ucount (computed here by block 0 ... we need it before: separate tiny kernel).
__global__ void sm_prefix_kernel(const int *ucount, const int *status, int T, int *coloff, int *nc_out, int *status_out) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        int run = 0, st = 0;
        for (int t = 0; t < T; t++) {
            coloff[t] = run;
            run += ucount[t];
            if (status[t] != 0 && st == 0) st = status[t];
        }
        coloff[T] = run;
        *nc_out = run;
        *status_out = st;
    }
}

__global__ void __launch_bounds__(MT)
sm_codes_kernel(WmArgs A, const int *coloff, int *code, int *corder, int *coff) {
    const int t = blockIdx.x, tid = threadIdx.x;
    const int64_t s0 = A.start[t];
    const int N = (int)(A.start[t + 1] - s0);
    const int nu = A.ucount[t];
    const int off = coloff[t];
    __shared__ int cnt[WM_MAXU];
    if (A.status[t] != 0) {
        if (tid == 0 && t == gridDim.x - 1) coff[off] = (int)A.start[t + 1];
        return;
    }
    for (int q = tid; q < nu; q += MT) cnt[q] = 0;
    __syncthreads();
    for (int i = tid; i < N; i += MT) {
        int f = A.fcode[s0 + i];
        code[s0 + i] = off + f;
        atomicAdd(&cnt[f], 1);
    }
    __syncthreads();
    if (tid == 0) {
        int run = (int)s0;
        for (int q = 0; q < nu; q++) {
            coff[off + q] = run;
            run += cnt[q];
        }
        if (t == gridDim.x - 1) coff[off + nu] = run;
    }
    __syncthreads();
    for (int q = tid >> 5; q < nu; q += MT / 32) { /* one warp per cluster, ascending ballots */
        int pos = coff[off + q];
        const int lane = tid & 31;
        for (int i0 = 0; i0 < N; i0 += 32) {
            const int i = i0 + lane;
            const bool in = i < N && A.fcode[s0 + i] == q;
            const unsigned bal = __ballot_sync(0xffffffffu, in);
            if (in) corder[pos + __popc(bal & ((1u << lane) - 1u))] = (int)(s0 + i);
            pos += __popc(bal);
        }
    }
}

// colMeans(sE1[cluster, ]) in ascending row order: one CTA per cluster, threads over the p columns
__global__ void __launch_bounds__(MT, 2)
sm_centroids_kernel(const double *__restrict__ E1, int p, const int *__restrict__ corder, const int *__restrict__ coff,
                    const int *nc_ptr, double *__restrict__ cen, int64_t *counts) {
    const int c = blockIdx.x;
    if (c >= *nc_ptr) return;
    const int q0 = coff[c], q1 = coff[c + 1];
    for (int d = threadIdx.x; d < p; d += MT) {
        double s = 0.0;
        int q = q0;
        for (; q + 32 <= q1; q += 32) { /* 32 rows in flight (part-level clusters hold thousands of cells); the adds stay in ascending row order */
            double v[32];
#pragma unroll
            for (int u = 0; u < 32; u++) v[u] = E1[(size_t)corder[q + u] * p + d];
#pragma unroll
            for (int u = 0; u < 32; u++) s = __dadd_rn(s, v[u]);
        }
        for (; q + 8 <= q1; q += 8) {
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) v[u] = E1[(size_t)corder[q + u] * p + d];
#pragma unroll
            for (int u = 0; u < 8; u++) s = __dadd_rn(s, v[u]);
        }
        for (; q < q1; q++) s = __dadd_rn(s, E1[(size_t)corder[q] * p + d]);
        cen[(size_t)c * p + d] = __ddiv_rn(s, (double)(q1 - q0));
    }
    if (counts && threadIdx.x == 0) counts[c] = q1 - q0;
}

// per-centroid mean (with R's refinement pass) and standard deviation: one thread per centroid
__global__ void __launch_bounds__(MT)
sm_stats_kernel(const double *__restrict__ cen, int p, const int *nc_ptr, double *__restrict__ mean, double *__restrict__ sdev) {
    const int c = blockIdx.x * MT + threadIdx.x;
    if (c >= *nc_ptr) return;
    const double *v = cen + (size_t)c * p;
    double sum = 0.0;
    for (int k = 0; k < p; k++) sum = __dadd_rn(sum, v[k]);
    double tmp = __ddiv_rn(sum, (double)p);
    if (isfinite(tmp)) {
        sum = 0.0;
        for (int k = 0; k < p; k++) sum = __dadd_rn(sum, __dsub_rn(v[k], tmp));
        tmp = __dadd_rn(tmp, __ddiv_rn(sum, (double)p));
    }
    double s = 0.0;
    for (int k = 0; k < p; k++) {
        double d = __dsub_rn(v[k], tmp);
        s = __dadd_rn(s, __dmul_rn(d, d));
    }
    mean[c] = tmp;
    sdev[c] = sqrt(__ddiv_rn(s, (double)(p - 1)));
}

// S[i,j] = cor(aG[i,], aG[j,]) (stats C cov_complete2 order); D = 1 - S.  one thread per pair i <= j
__global__ void __launch_bounds__(MT)
sm_cor_kernel(const double *__restrict__ cen, int p, const int *nc_ptr, const double *__restrict__ mean,
              const double *__restrict__ sdev, int ld, double *S, double *D, double *Dw) {
    const int nC = *nc_ptr;
    const size_t idx = (size_t)blockIdx.x * MT + threadIdx.x;
    if (nC <= 0 || idx >= (size_t)nC * nC) return;
    const int i = (int)(idx / nC), j = (int)(idx - (size_t)i * nC);
    if (j < i) return;
    if (i == j) {
        S[(size_t)i * ld + i] = 1.0;
        D[(size_t)i * ld + i] = 0.0;
        Dw[(size_t)i * ld + i] = 0.0;
        return;
    }
    const double *x = cen + (size_t)i * p, *y = cen + (size_t)j * p;
    const double xm = mean[i], ym = mean[j];
    double sum = 0.0;
    for (int k = 0; k < p; k++) sum = __dadd_rn(sum, __dmul_rn(__dsub_rn(x[k], xm), __dsub_rn(y[k], ym)));
    double ans = __ddiv_rn(sum, (double)(p - 1));
    double r;
    if (sdev[i] == 0.0 || sdev[j] == 0.0) r = __longlong_as_double(0x7ff8000000000000LL);
    else {
        r = __ddiv_rn(ans, __dmul_rn(sdev[i], sdev[j]));
        if (r > 1.0) r = 1.0;
        if (r < -1.0) r = -1.0;
    }
    const double d = __dsub_rn(1.0, r);
    S[(size_t)i * ld + j] = S[(size_t)j * ld + i] = r;
    D[(size_t)i * ld + j] = D[(size_t)j * ld + i] = d;
    Dw[(size_t)i * ld + j] = Dw[(size_t)j * ld + i] = d;
}

// problem descriptor + the k-range tweak of R/sMetaC.R:101-119 (one thread)
__global__ void sm_setup_kernel(SmArgs A) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    const int nC = *A.nc_ptr;
    HcProb &P = *A.prob;
    P.n = nC;
    P.ld = A.ld;
    P.D = A.D;
    P.Dw = A.Dw;
    P.ia = A.ia;
    P.ib = A.ib;
    P.crit = A.crit;
    P.Y = A.S;
    P.p = nC;
    P.ldy = A.ld;
    int st = *A.status_in;
    if (st == 0 && nC < 2) st = 24;
    if (st == 0 && nC > A.ld) st = WM_E_TOOMANY;
    P.status = st;
    if (st != 0) P.n = 0;
    HcParamsDev prm = A.prm;
    int minN = prm.min_n, maxN = prm.max_n;
    const int64_t ncells = A.ncells_total;
    const int mm = (int)(ncells / 10000);
    if (ncells < 1000000) {
        int baseN = min(max(mm, 2), 10);
        if (minN == 2 && min(maxN, nC) - baseN >= 3) minN = baseN;
    } else {
        int mm3 = (int)(ncells / 50000), mm2 = (int)(ncells / 5000);
        maxN = max(maxN, mm2);
        minN = max(minN, mm3);
    }
    prm.min_n = minN;
    prm.max_n = maxN;
    *A.prm_out = prm;
}

// "second best if 2 clusters" rule (R/sMetaC.R:139-151) -> tf[nC]; one CTA
__global__ void __launch_bounds__(MT) sm_finish_kernel(SmArgs A, const SweepOut *out) {
    extern __shared__ int fsm[];
    __shared__ int tmp_scan[MT];
    __shared__ int s_level;
    const HcProb &P = *A.prob;
    const SweepOut &O = *out;
    const int n = P.n, tid = threadIdx.x;
    if (tid == 0) *A.status_out = (P.status != 0) ? P.status : O.meta[3];
    if (n <= 0 || P.status != 0 || O.meta[3] != 0) return;
    const HcParamsDev prm = *A.prm_out;
    const int nlev = O.meta[0], kmin = O.meta[4];
    if (tid == 0) {
        int lev = -1;
        if (nlev > 1 && O.meta[1] == 2 && *O.maxsil > prm.sil_thre) {
            /* s1 = second largest value of msil (with multiplicity); s2 = first index holding it (quirk B4) */
            double top = -SHARP_INF, second = -SHARP_INF;
            int ntop = 0;
            for (int i = 0; i < nlev; i++) {
                double v = O.msil[i];
                if (v > top) { second = top; top = v; ntop = 1; }
                else if (v == top) ntop++;
                else if (v > second) second = v;
            }
            double s1 = (ntop >= 2) ? top : second;
            for (int i = 0; i < nlev; i++)
                if (O.msil[i] == s1) { lev = i; break; }
        }
        s_level = lev;
    }
    __syncthreads();
    const int lev = s_level;
    if (lev < 0) {
        for (int i = tid; i < n; i += MT) A.tf[i] = O.f[i];
        return;
    }
    int *step_of = fsm, *link = fsm + A.ld, *rank = fsm + 2 * A.ld, *lab = fsm + 3 * A.ld;
    // labels at k = kmin + lev (same helpers as sweep.cu, inlined here)
    for (int i = tid; i < n; i += MT) { step_of[i] = INT_MAX; link[i] = i; }
    __syncthreads();
    for (int s = tid; s < n - 1; s += MT) { int j2 = P.ib[s] - 1; step_of[j2] = s; link[j2] = P.ia[s] - 1; }
    __syncthreads();
    const int nmerge = n - (kmin + lev);
    for (int i = tid; i < n; i += MT) rank[i] = (step_of[i] >= nmerge) ? 1 : 0;
    __syncthreads();
    block_exclusive_scan<MT>(rank, n, tmp_scan);
    for (int i = tid; i < n; i += MT) {
        int x = i;
        while (step_of[x] < nmerge) x = link[x];
        lab[i] = rank[x];
    }
    __syncthreads();
    for (int i = tid; i < n; i += MT) A.tf[i] = lab[i] + 1;
}

// finalColor = tf[code], optionally scattered to the un-shuffled position
__global__ void __launch_bounds__(MT)
sm_relabel_kernel(int64_t ncells, const int *__restrict__ code, const int *__restrict__ tf, int add,
                  const int64_t *__restrict__ out_row, int *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * MT + threadIdx.x;
    if (i >= ncells) return;
    const int64_t r = out_row ? out_row[i] : i;
    out[r] = tf ? tf[code[i]] : code[i] + add;
}

int launch_sm_codes(sharp_ctx *c, const WmArgs &A, int T, int *coloff, int *nc_out, int *status_out, int *code,
                    int *corder, int *coff) {
    if (A.capU > WM_MAXU) return set_error(SHARP_E_LIMIT, "more than %d clusters per block", WM_MAXU);
    prof_begin(c, KID_SMETAC);
    sm_prefix_kernel<<<1, 32, 0, c->stream>>>(A.ucount, A.status, T, coloff, nc_out, status_out);
    prof_end(c);
    prof_begin(c, KID_SMETAC);
    sm_codes_kernel<<<T, MT, 0, c->stream>>>(A, coloff, code, corder, coff);
    prof_end(c);
    SHARP_CUDA(cudaGetLastError());
    return 0;
}

int launch_sm_centroids(sharp_ctx *c, const double *E1, int p, const int *corder, const int *coff, const int *nc_ptr,
                        int nc_cap, double *cen, int64_t *counts) {
    if (nc_cap <= 0) return 0;
    prof_begin(c, KID_SM_CENTROIDS);
    sm_centroids_kernel<<<nc_cap, MT, 0, c->stream>>>(E1, p, corder, coff, nc_ptr, cen, counts);
    prof_end(c);
    SHARP_CUDA(cudaGetLastError());
    return 0;
}

int launch_sm_similarity(sharp_ctx *c, const SmArgs &A, const double *cen, int p, int nc_cap, double *mean, double *sdev) {
    prof_begin(c, KID_SMETAC);
    sm_stats_kernel<<<(nc_cap + MT - 1) / MT, MT, 0, c->stream>>>(cen, p, A.nc_ptr, mean, sdev);
    prof_end(c);
    size_t pairs = (size_t)nc_cap * nc_cap;
    prof_begin(c, KID_SMETAC);
    sm_cor_kernel<<<(unsigned)((pairs + MT - 1) / MT), MT, 0, c->stream>>>(cen, p, A.nc_ptr, mean, sdev, A.ld, A.S, A.D, A.Dw);
    prof_end(c);
    prof_begin(c, KID_SMETAC);
    sm_setup_kernel<<<1, 32, 0, c->stream>>>(A);
    prof_end(c);
    SHARP_CUDA(cudaGetLastError());
    return 0;
}

int launch_sm_finish(sharp_ctx *c, const SmArgs &A, const SweepOut *out) {
    size_t smem = (size_t)4 * A.ld * 4;
    SHARP_SMEM_OPTIN_ONCE((sm_finish_kernel), c->device);
    prof_begin(c, KID_SMETAC);
    sm_finish_kernel<<<1, MT, smem, c->stream>>>(A, out);
    prof_end(c);
    SHARP_CUDA(cudaGetLastError());
    return 0;
}

int launch_sm_relabel(sharp_ctx *c, int64_t ncells, const int *code, const int *tf, int add, const int64_t *out_row, int *out) {
    if (ncells <= 0) return 0;
    prof_begin(c, KID_SMETAC);
    sm_relabel_kernel<<<(unsigned)((ncells + MT - 1) / MT), MT, 0, c->stream>>>(ncells, code, tf, add, out_row, out);
    prof_end(c);
    SHARP_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace sharp
