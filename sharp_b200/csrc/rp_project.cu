// rp_project.cu -- K1: very-sparse ternary random projection with the normalisation fused into the load.
//
// Replaces  1/sqrt(p) * t(rM[[k]]) %*% log2(E[, tind] + 1)   for all K ensemble members at once
// (R/SHARP.R:567-585, R/RPmat.R:32, R/SHARP_unlimited2.R:388-410) and t(t(x)/colSums(x))*1e6 (R/SHARP.R:113).
//
// Formulation.  The ranM matrices are ternary (+-sqrt(s), density 1/sqrt(m)), and single-cell expression is
// itself sparse (70-95 % zeros), so the product is driven by the NON-ZEROS OF THE EXPRESSION COLUMN: for every
// non-zero gene of a cell, the ~K*p/sqrt(m) (12-18) projection columns that contain this gene receive +-value.
// That is 10-15x fewer operations than gathering along the projection columns, and the expression matrix is
// read exactly once for all K members (HBM traffic = the input as stored + the K*p outputs).
//   * the ternary matrices are re-laid out gene-major (CSR over genes of the concatenated m x K*p matrix,
//     16-bit column + sign per entry, ~1 MB in all): they stay L2/L1-resident and are read directly, 30 B per
//     non-zero of the cell;
//   * one warp owns one cell: its K*p fp64 accumulators live in shared memory (20 KB at K = 5, p = 508), its
//     non-zeros are loaded 32 at a time (coalesced), log-transformed in parallel, staged in shared memory, then
//     applied one gene at a time with the lanes spread over that gene's entries (distinct columns, so no atomics
//     and every accumulator receives its terms in ascending gene order: bit-identical to a sequential sparse
//     product).  The entry loads of 4 genes are in flight before the first add.
//   * bound: the kernel is limited by shared-memory read-modify-write wavefronts of the scattered adds (random
//     columns -> bank conflicts), not by HBM: ~30 k adds per cell against 44 KB of HBM traffic per cell.
//   * a dense tcgen05 contraction is not used: with density 1/sqrt(m) it would execute ~150x more MACs than
//     this kernel does adds (DESIGN.md has the arithmetic).
#include "devutil.cuh"
#include "internal.cuh"

namespace sharp {

__device__ __forceinline__ double rp_transform(double x, double cs, int normalize, double norm_mul, int logkind) {
    double v = x;
    if (normalize) v = __dmul_rn(__ddiv_rn(x, cs), norm_mul);
    if (logkind == 2) v = log2(v + 1.0);
    else if (logkind == 10) v = log10(v + 1.0);
    return v;
}

// base::round(x, digits) in the R >= 4.0 flavour restated by the oracle (closest candidate, ties to even)
__device__ __forceinline__ double rp_round(double x, int digits) {
    if (digits < 0 || x == 0.0 || !isfinite(x)) return x;
    double p10 = 1.0;
    for (int i = 0; i < digits; i++) p10 *= 10.0;
    double xd = __dmul_rn(x, p10);
    double fl = floor(xd), ce = ceil(xd);
    double lo = __ddiv_rn(fl, p10), hi = __ddiv_rn(ce, p10);
    double dl = __dsub_rn(x, lo), dh = __dsub_rn(hi, x);
    if (dl < dh) return lo;
    if (dh < dl) return hi;
    return (fmod(fl, 2.0) == 0.0) ? lo : hi;
}

// ---- column sums (colSums(x), R/SHARP.R:113) : one warp per cell -------------------------------------------
__global__ void __launch_bounds__(256)
colsum_kernel(int m, int64_t n, const double *__restrict__ dense, const int64_t *__restrict__ colptr,
              const double *__restrict__ val, double *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t c = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= n) return;
    double s = 0.0;
    if (dense) {
        const double *x = dense + c * m;
        for (int i = lane; i < m; i += 32) s += x[i];
    } else {
        for (int64_t q = colptr[c] + lane; q < colptr[c + 1]; q += 32) s += val[q];
    }
    s = warp_sum(s);
    if (lane == 0) out[c] = s;
}

int launch_colsum(sharp_ctx *c, const sharp_expr_dev &e, double *colsum) {
    if (e.n <= 0) return 0;
    int64_t blocks = (e.n + 7) / 8;
    prof_begin(c, KID_COLSUM);
    colsum_kernel<<<(unsigned)blocks, 256, 0, c->stream>>>(e.m, e.n, e.dense, e.colptr, e.val, colsum);
    prof_end(c);
    SHARP_CUDA(cudaGetLastError());
    return 0;
}

// ---- projection ------------------------------------------------------------------------------------
struct RpArgs {
    int m;
    int64_t n;
    const double *dense;
    const int64_t *colptr;
    const int32_t *rowidx;
    const double *val;
    const int64_t *cells;   // source column per output cell (or null)
    int64_t ncell;
    const double *colsum;
    int normalize;
    double norm_mul;
    int logkind;
    int round_digits;
    int p, K, KP;
    double scale;           // mag / sqrt(p)
    const uint32_t *rowptr; // [m+1]
    const uint16_t *ent16;
    const uint32_t *ent32;
    double *out;
};

// one staged non-zero of the current cell: where its ranM entries are and the transformed value
struct __align__(16) RpGene {
    uint32_t r0, cnt;
    double v;
};

template <bool ENT16>
__device__ __forceinline__ unsigned rp_load_ent(const RpArgs &A, uint32_t at) {
    return ENT16 ? (unsigned)__ldg(A.ent16 + at) : __ldg(A.ent32 + at);
}

template <bool ENT16>
__device__ __forceinline__ void rp_add(double *acc, unsigned e, double v) {
    const unsigned col = ENT16 ? (e & 0x7fffu) : (e & 0x7fffffffu);
    const unsigned neg = ENT16 ? (e & 0x8000u) : (e & 0x80000000u);
    acc[col] += neg ? -v : v;
}

// Apply up to 32 staged genes (in order) to the warp's accumulators.  Lanes spread over the entries of ONE gene
// (distinct columns: no conflicts, and every accumulator receives its terms in ascending gene order, like the
// reference's sparse product).  The accumulators take almost all of the SM's shared memory, so L1 is tiny and every
// entry load is an L2 round trip: the first 32 entries of ALL staged genes are loaded before the first add (one
// round trip per 32 genes instead of one per gene).
template <bool ENT16>
__device__ __forceinline__ void rp_apply_staged(const RpArgs &A, double *acc, const RpGene *st, int nst, int lane) {
    unsigned e[32];
#pragma unroll
    for (int t = 0; t < 32; t++) {
        e[t] = 0u;
        if (t < nst) {
            const uint32_t r0 = st[t].r0, cnt = st[t].cnt;
            if ((uint32_t)lane < cnt) e[t] = rp_load_ent<ENT16>(A, r0 + lane);
        }
    }
#pragma unroll
    for (int t = 0; t < 32; t++) {
        if (t < nst) { /* nst is warp-uniform */
            const RpGene g = st[t];
            if ((uint32_t)lane < g.cnt) rp_add<ENT16>(acc, e[t], g.v);
            for (uint32_t x = 32 + lane; x < g.cnt; x += 32) rp_add<ENT16>(acc, rp_load_ent<ENT16>(A, g.r0 + x), g.v);
            __syncwarp();
        }
    }
}

// one warp per cell; dynamic shared memory: [W][KP] fp64 accumulators, then [W][32] staged genes
template <bool ENT16>
__global__ void __launch_bounds__(320) rp_project_kernel(RpArgs A, int warps_per_cta) {
    extern __shared__ __align__(16) unsigned char rsm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *acc = reinterpret_cast<double *>(rsm) + (size_t)warp * A.KP;
    RpGene *stage = reinterpret_cast<RpGene *>(reinterpret_cast<double *>(rsm) + (size_t)warps_per_cta * A.KP) + warp * 32;

    const int64_t nwarps = (int64_t)gridDim.x * warps_per_cta;
    for (int64_t pos = (int64_t)blockIdx.x * warps_per_cta + warp; pos < A.ncell; pos += nwarps) {
        const int64_t src = A.cells ? A.cells[pos] : pos;
        const double cs = A.normalize ? A.colsum[src] : 1.0;
        for (int i = lane; i < A.KP; i += 32) acc[i] = 0.0;
        __syncwarp();
        if (A.dense) {
            const double *dcol = A.dense + src * A.m;
            for (int c0 = 0; c0 < A.m; c0 += 32) {
                const int gi = c0 + lane;
                const double x = (gi < A.m) ? dcol[gi] : 0.0;
                const unsigned mask = __ballot_sync(0xffffffffu, x != 0.0);
                if (!mask) continue;
                if (x != 0.0) {
                    RpGene g;
                    g.r0 = __ldg(A.rowptr + gi);
                    g.cnt = __ldg(A.rowptr + gi + 1) - g.r0;
                    g.v = rp_transform(x, cs, A.normalize, A.norm_mul, A.logkind);
                    stage[__popc(mask & ((1u << lane) - 1u))] = g;
                }
                __syncwarp();
                rp_apply_staged<ENT16>(A, acc, stage, __popc(mask), lane);
            }
        } else {
            const int64_t q0 = A.colptr[src], q1 = A.colptr[src + 1];
            for (int64_t q = q0; q < q1; q += 32) {
                const int nst = (int)min((int64_t)32, q1 - q);
                if (lane < nst) {
                    const int gi = A.rowidx[q + lane];
                    const double x = A.val[q + lane];
                    RpGene g;
                    g.r0 = __ldg(A.rowptr + gi);
                    g.cnt = (x != 0.0) ? __ldg(A.rowptr + gi + 1) - g.r0 : 0u; /* explicit zeros contribute nothing */
                    g.v = rp_transform(x, cs, A.normalize, A.norm_mul, A.logkind);
                    stage[lane] = g;
                }
                __syncwarp();
                rp_apply_staged<ENT16>(A, acc, stage, nst, lane);
            }
        }
        for (int i = lane; i < A.KP; i += 32) {
            const int k = i / A.p, j = i - k * A.p;
            double r = __dmul_rn(acc[i], A.scale);
            if (A.round_digits >= 0) r = rp_round(r, A.round_digits);
            A.out[((size_t)k * A.ncell + pos) * A.p + j] = r;
        }
        __syncwarp();
    }
}

int launch_rp_project(sharp_ctx *c, const sharp_expr_dev &e, const int64_t *cells_dev, int64_t ncell,
                      const double *colsum_dev, int normalize, double norm_mul, int logkind, int round_digits,
                      const sharp_rm_dev &rm, double *out) {
    if (ncell <= 0) return 0;
    if (e.m != rm.m) return set_error(SHARP_E_ARG, "rp_project: expression has %d genes but ranM has %d rows", e.m, rm.m);
    RpArgs A;
    A.m = e.m; A.n = e.n; A.dense = e.dense; A.colptr = e.colptr; A.rowidx = e.rowidx; A.val = e.val;
    A.cells = cells_dev; A.ncell = ncell; A.colsum = colsum_dev; A.normalize = normalize ? 1 : 0;
    A.norm_mul = norm_mul; A.logkind = logkind; A.round_digits = round_digits;
    A.p = rm.p; A.K = rm.K; A.KP = rm.K * rm.p;
    A.scale = (1.0 / sqrt((double)rm.p)) * rm.mag;   /* entry of 1/sqrt(p) * t(rM) */
    A.rowptr = rm.rowptr; A.ent16 = rm.ent16; A.ent32 = rm.ent32;
    A.out = out;
    const bool e16 = rm.ent16 != nullptr;
    const size_t budget = 220 * 1024;
    const size_t per_warp = (size_t)A.KP * 8 + 32 * sizeof(RpGene);
    if (per_warp > budget)
        return set_error(SHARP_E_LIMIT, "rp_project: K*p = %d accumulators do not fit in shared memory", A.KP);
    int W = (int)(budget / per_warp);
    if (W > 10) W = 10;
    const size_t smem = per_warp * W;
    const int64_t nbatch = (ncell + W - 1) / W;
    const int grid = (int)std::min<int64_t>(nbatch, (int64_t)c->sm_count);
    prof_begin(c, KID_RP_PROJECT);
    if (e16) {
        SHARP_CUDA(cudaFuncSetAttribute(rp_project_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SHARP_SMEM_OPTIN));
        rp_project_kernel<true><<<grid, W * 32, smem, c->stream>>>(A, W);
    } else {
        SHARP_CUDA(cudaFuncSetAttribute(rp_project_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SHARP_SMEM_OPTIN));
        rp_project_kernel<false><<<grid, W * 32, smem, c->stream>>>(A, W);
    }
    prof_end(c);
    SHARP_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace sharp
