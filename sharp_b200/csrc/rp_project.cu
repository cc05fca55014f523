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
//     16-bit column + sign per entry) and streamed through shared memory in tiles of TILE_GENES genes, shared
//     by all cells a CTA is working on;
//   * one warp owns one cell: its K*p fp64 accumulators live in shared memory, its non-zeros are loaded 32 at a
//     time (coalesced), log-transformed in parallel, then applied one gene at a time with the lanes spread over
//     that gene's entries (distinct columns, so no atomics and a deterministic result);
//   * a dense tcgen05 contraction is not used: with density 1/sqrt(m) it would execute ~150x more MACs than
//     this kernel does adds (DESIGN.md has the arithmetic).
#include "devutil.cuh"
#include "internal.cuh"

namespace sharp {

constexpr int RP_TILE_GENES = 512;

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
    int tile_genes, ntiles, max_tile_entries;
    double *out;
};

template <bool ENT16>
__device__ __forceinline__ void rp_apply(double *acc, const void *s_ent, int r0, int r1, int lane, double vv) {
    for (int e = r0 + lane; e < r1; e += 32) {
        unsigned col, neg;
        if (ENT16) {
            unsigned v = reinterpret_cast<const uint16_t *>(s_ent)[e];
            col = v & 0x7fffu;
            neg = v & 0x8000u;
        } else {
            unsigned v = reinterpret_cast<const uint32_t *>(s_ent)[e];
            col = v & 0x7fffffffu;
            neg = v & 0x80000000u;
        }
        acc[col] += neg ? -vv : vv;
    }
    __syncwarp();
}

template <bool ENT16>
__global__ void __launch_bounds__(256) rp_project_kernel(RpArgs A, int warps_per_cta) {
    extern __shared__ __align__(16) unsigned char rsm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nthreads = warps_per_cta * 32;
    double *acc_all = reinterpret_cast<double *>(rsm);                         // [W][KP]
    uint32_t *s_rowptr = reinterpret_cast<uint32_t *>(acc_all + (size_t)warps_per_cta * A.KP);  // [tile_genes+1]
    unsigned char *s_ent = reinterpret_cast<unsigned char *>(s_rowptr + A.tile_genes + 4);
    double *acc = acc_all + (size_t)warp * A.KP;
    const int esz = ENT16 ? 2 : 4;

    const int64_t nbatch = (A.ncell + warps_per_cta - 1) / warps_per_cta;
    for (int64_t batch = blockIdx.x; batch < nbatch; batch += gridDim.x) {
        const int64_t pos = batch * warps_per_cta + warp;   // output cell of this warp
        const bool active = pos < A.ncell;
        const int64_t src = active ? (A.cells ? A.cells[pos] : pos) : 0;
        const double cs = (A.normalize && active) ? A.colsum[src] : 1.0;
        for (int i = lane; i < A.KP; i += 32) acc[i] = 0.0;
        // CSC cursor
        int64_t q = 0, qend = 0;
        if (!A.dense && active) { q = A.colptr[src]; qend = A.colptr[src + 1]; }
        int chunk_len = 0, consumed = 0;
        int g = INT_MAX;
        double v = 0.0;
        const double *dcol = A.dense ? A.dense + src * A.m : nullptr;

        for (int tile = 0; tile < A.ntiles; tile++) {
            const int g0 = tile * A.tile_genes;
            const int g1 = min(A.m, g0 + A.tile_genes);
            const uint32_t ebase = A.rowptr[g0];
            const uint32_t eend = A.rowptr[g1];
            __syncthreads(); /* previous tile fully consumed */
            for (int i = threadIdx.x; i <= g1 - g0; i += nthreads) s_rowptr[i] = A.rowptr[g0 + i] - ebase;
            {
                /* entries: copy as 4-byte words (ebase*esz may be only 2-byte aligned -> start from the aligned word) */
                const size_t b0 = (size_t)ebase * esz, b1 = (size_t)eend * esz;
                const size_t w0 = b0 & ~(size_t)3;
                const uint32_t *srcw = reinterpret_cast<const uint32_t *>(
                    (ENT16 ? reinterpret_cast<const unsigned char *>(A.ent16) : reinterpret_cast<const unsigned char *>(A.ent32)) + w0);
                uint32_t *dstw = reinterpret_cast<uint32_t *>(s_ent);
                const size_t nw = (b1 - w0 + 3) >> 2;
                for (size_t i = threadIdx.x; i < nw; i += nthreads) dstw[i] = srcw[i];
            }
            __syncthreads();
            const unsigned char *ent = s_ent + (((size_t)ebase * esz) & 3); /* skip the alignment slack */
            if (!active) continue;
            if (dcol) {
                for (int c0 = g0; c0 < g1; c0 += 32) {
                    const int gi = c0 + lane;
                    double x = (gi < g1) ? dcol[gi] : 0.0;
                    unsigned mask = __ballot_sync(0xffffffffu, x != 0.0);
                    if (!mask) continue;
                    double tv = (x != 0.0) ? rp_transform(x, cs, A.normalize, A.norm_mul, A.logkind) : 0.0;
                    while (mask) {
                        const int t = __ffs(mask) - 1;
                        mask &= mask - 1;
                        const double vv = __shfl_sync(0xffffffffu, tv, t);
                        const int gl = c0 + t - g0;
                        rp_apply<ENT16>(acc, ent, (int)s_rowptr[gl], (int)s_rowptr[gl + 1], lane, vv);
                    }
                }
            } else {
                while (true) {
                    if (consumed == chunk_len) {
                        q += chunk_len;
                        chunk_len = (int)min((int64_t)32, qend - q);
                        consumed = 0;
                        if (chunk_len <= 0) { chunk_len = 0; break; }
                        if (lane < chunk_len) {
                            g = A.rowidx[q + lane];
                            v = rp_transform(A.val[q + lane], cs, A.normalize, A.norm_mul, A.logkind);
                        } else {
                            g = INT_MAX;
                            v = 0.0;
                        }
                    }
                    unsigned mask = __ballot_sync(0xffffffffu, lane >= consumed && lane < chunk_len && g < g1);
                    const int cnt = __popc(mask);
                    if (cnt == 0) break;
                    for (int t = consumed; t < consumed + cnt; t++) {
                        const int gl = __shfl_sync(0xffffffffu, g, t) - g0;
                        const double vv = __shfl_sync(0xffffffffu, v, t);
                        if (vv != 0.0) rp_apply<ENT16>(acc, ent, (int)s_rowptr[gl], (int)s_rowptr[gl + 1], lane, vv);
                    }
                    consumed += cnt;
                    if (consumed < chunk_len) break;
                }
            }
        }
        __syncwarp();
        if (active) {
            for (int i = lane; i < A.KP; i += 32) {
                const int k = i / A.p, j = i - k * A.p;
                double r = __dmul_rn(acc[i], A.scale);
                if (A.round_digits >= 0) r = rp_round(r, A.round_digits);
                A.out[((size_t)k * A.ncell + pos) * A.p + j] = r;
            }
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
    A.tile_genes = rm.tile_genes; A.ntiles = rm.ntiles; A.max_tile_entries = rm.max_tile_entries;
    A.out = out;
    const bool e16 = rm.ent16 != nullptr;
    const size_t tile_bytes = (size_t)(rm.tile_genes + 4) * 4 + (size_t)rm.max_tile_entries * (e16 ? 2 : 4) + 16;
    const size_t budget = 200 * 1024;
    const size_t per_warp = (size_t)A.KP * 8;
    if (tile_bytes + per_warp > budget)
        return set_error(SHARP_E_LIMIT, "rp_project: K*p = %d accumulators do not fit in shared memory", A.KP);
    int W = (int)((budget - tile_bytes) / per_warp);
    if (W > 8) W = 8;
    size_t smem = tile_bytes + per_warp * W;
    smem = (smem + 15) & ~(size_t)15;
    int64_t nbatch = (ncell + W - 1) / W;
    int grid = (int)std::min<int64_t>(nbatch, (int64_t)c->sm_count * (smem <= 100 * 1024 ? 2 : 1));
    prof_begin(c, KID_RP_PROJECT);
    if (e16) {
        SHARP_CUDA(cudaFuncSetAttribute(rp_project_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        rp_project_kernel<true><<<grid, W * 32, smem, c->stream>>>(A, W);
    } else {
        SHARP_CUDA(cudaFuncSetAttribute(rp_project_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        rp_project_kernel<false><<<grid, W * 32, smem, c->stream>>>(A, W);
    }
    prof_end(c);
    SHARP_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace sharp
