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
// resident at once (up to 4 CTAs of 512 threads per SM).
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

template <int THREADS>
__global__ void __launch_bounds__(THREADS) hclust_kernel(HcProb *probs, int method) {
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
    int *nn = reinterpret_cast<int *>(disnn + n);           // [n]
    int *membr = nn + n;                                    // [n]
    int *list = membr + n;                                  // [n]
    unsigned char *flag = reinterpret_cast<unsigned char *>(list + n);  // [n]

    if (method == SHARP_WARD_D2) { /* hclust.f: iOpt 8 works on squared dissimilarities */
        for (size_t idx = tid; idx < (size_t)n * ld; idx += THREADS) D[idx] = __dmul_rn(D[idx], D[idx]);
    }
    for (int i = tid; i < n; i += THREADS) {
        flag[i] = 1;
        membr[i] = 1;
        disnn[i] = SHARP_INF;
        nn[i] = -1;
    }
    __syncthreads();

    // initial nearest neighbours: NN(i) = first minimum over j > i
    for (int i = warp; i < n - 1; i += NW) {
        DI best;
        best.d = SHARP_INF;
        best.i = INT_MAX;
        const double *row = D + (size_t)i * ld;
        for (int j = i + 1 + lane; j < n; j += 32) {
            double d = row[j];
            if (d < best.d) { best.d = d; best.i = j; }
        }
        best = warp_argmin(best);
        if (lane == 0) {
            nn[i] = (best.i == INT_MAX) ? -1 : best.i;
            disnn[i] = best.d;
        }
    }
    __syncthreads();

    for (int step = 0; step < n - 1; ++step) {
        // ---- least dissimilarity over the NN list (first strict minimum over i) ----
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
        const double d12 = D[(size_t)i2 * ld + j2];
        const double mi = (double)membr[i2], mj = (double)membr[j2];
        if (tid == 0) {
            P.ia[step] = i2 + 1;
            P.ib[step] = j2 + 1;
            P.crit[step] = (method == SHARP_WARD_D2) ? sqrt(c.d) : c.d;
        }
        __syncthreads(); /* everybody has read nn[im], membr[] before they change */
        if (tid == 0) flag[j2] = 0;
        __syncthreads();

        // ---- update dissimilarities from the new cluster (kept under index i2) ----
        DI nb;
        nb.d = SHARP_INF;
        nb.i = INT_MAX;
        const double *rowi = D + (size_t)i2 * ld;
        const double *rowj = D + (size_t)j2 * ld;
        for (int k = tid; k < n; k += THREADS) {
            if (!flag[k] || k == i2) continue;
            double r = lance_williams(method, rowi[k], rowj[k], d12, mi, mj, (double)membr[k]);
            D[(size_t)i2 * ld + k] = r;
            D[(size_t)k * ld + i2] = r;
            if (k > i2) {
                if (r < nb.d) { nb.d = r; nb.i = k; }
            } else if (r < disnn[k]) { /* hclust.f "FIX by JB": i2 may become the NN of a smaller index */
                disnn[k] = r;
                nn[k] = i2;
            }
        }
        nb = block_argmin<THREADS>(nb, red);
        if (tid == 0) {
            membr[i2] = membr[i2] + membr[j2];
            disnn[i2] = nb.d;
            nn[i2] = (nb.i == INT_MAX) ? -1 : nb.i;
            s_cnt = 0;
        }
        __syncthreads();

        // ---- redetermine the NN of every i whose NN was i2 or j2 ----
        for (int i = tid; i < n - 1; i += THREADS) {
            if (flag[i]) {
                int q = nn[i];
                if (q == i2 || q == j2) list[atomicAdd(&s_cnt, 1)] = i;
            }
        }
        __syncthreads();
        const int cnt = s_cnt;
        for (int r = warp; r < cnt; r += NW) {
            const int i = list[r];
            DI best;
            best.d = SHARP_INF;
            best.i = INT_MAX;
            const double *row = D + (size_t)i * ld;
            for (int j = i + 1 + lane; j < n; j += 32) {
                if (flag[j]) {
                    double d = row[j];
                    if (d < best.d) { best.d = d; best.i = j; }
                }
            }
            best = warp_argmin(best);
            if (lane == 0) {
                nn[i] = (best.i == INT_MAX) ? -1 : best.i;
                disnn[i] = best.d;
            }
        }
        __syncthreads();
    }
}

static size_t hclust_smem_bytes(int n) { return ((size_t)n * (8 + 4 + 4 + 4 + 1) + 15) & ~(size_t)15; }

int launch_hclust(sharp_ctx *c, HcProb *probs_dev, int nprob, int max_n, int method) {
    if (nprob <= 0) return 0;
    if (method < 1 || method > 8) return set_error(SHARP_E_ARG, "invalid clustering method %d", method);
    size_t smem = hclust_smem_bytes(max_n);
    if (smem > 200 * 1024)
        return set_error(SHARP_E_LIMIT, "hclust: %d objects exceed the shared-memory NN list (max ~9700)", max_n);
    prof_begin(c, max_n > 384 ? KID_HCLUST : KID_HCLUST_SMALL);
    if (max_n > 384) {
        SHARP_CUDA(cudaFuncSetAttribute(hclust_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        hclust_kernel<512><<<nprob, 512, smem, c->stream>>>(probs_dev, method);
    } else {
        SHARP_CUDA(cudaFuncSetAttribute(hclust_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        hclust_kernel<128><<<nprob, 128, smem, c->stream>>>(probs_dev, method);
    }
    prof_end(c);
    SHARP_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace sharp
