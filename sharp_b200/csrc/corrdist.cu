// corrdist.cu -- K2: per-block pairwise correlation distance in the reduced space.
//
// Replaces  mat <- t(scale(t(mat))); d = as.dist(1 - cor(t(mat)))   (R/get_opt_hclust.R:71-72).
// Pearson correlation between two rows equals the dot product of the centred, unit-length rows, so
//   (1) unit_rows_kernel turns every projected cell (row of n x p) into a unit vector (fp64), zero-padded to a
//       multiple of 16 columns so that the contraction needs no tail handling and rows are 16-byte aligned;
//   (2) corrdist_kernel computes D = 1 - clamp(U U^T) for every (ensemble member, cell block) problem of a wave
//       with fp64 tensor-core MMAs (mma.sync m8n8k4 -- tcgen05 has no fp64 kind), 128x128 tiles, cp.async
//       double buffering; only the upper-triangular tiles are computed and each is mirrored on store.
// The distance matrix is written twice (D for the silhouette sweep, Dw as the agglomeration's working copy).
// This kernel carries most of the arithmetic of the whole path (4 GFLOP per 2000 x 508 block); it is bound by
// the fp64 pipe, not by HBM.
#include "devutil.cuh"
#include "internal.cuh"

namespace sharp {

// ---- (1) unit rows ---------------------------------------------------------------------------------
// one warp per row; X row-major rows x p; U row-major rows x ldu (ldu >= p, multiple of 16, zero padded)
__global__ void __launch_bounds__(256) unit_rows_kernel(const double *__restrict__ X, int64_t rows, int p, int ldu,
                                                        double *__restrict__ U) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const double *x = X + row * p;
    double s = 0.0;
    for (int j = lane; j < p; j += 32) s += x[j];
    s = warp_sum(s);
    const double mean = s / (double)p;
    double ss = 0.0;
    for (int j = lane; j < p; j += 32) {
        double c = x[j] - mean;
        ss = fma(c, c, ss);
    }
    ss = warp_sum(ss);
    const double inv = 1.0 / sqrt(ss); /* constant row -> NaN, like scale() producing NaN in the reference */
    double *u = U + row * ldu;
    for (int j = lane; j < ldu; j += 32) u[j] = (j < p) ? (x[j] - mean) * inv : 0.0;
}

int launch_unit_rows(sharp_ctx *c, const double *X, int64_t rows, int p, int ldu, double *U) {
    if (rows <= 0) return 0;
    const int wpb = 8;
    int64_t blocks = (rows + wpb - 1) / wpb;
    prof_begin(c, KID_UNIT_ROWS);
    unit_rows_kernel<<<(unsigned)blocks, wpb * 32, 0, c->stream>>>(X, rows, p, ldu, U);
    prof_end(c);
    SHARP_CUDA(cudaGetLastError());
    return 0;
}

// ---- (2) D = 1 - clamp(U U^T) ------------------------------------------------------------------------
constexpr int GT = 128;       // CTA tile (rows and columns)
constexpr int GK = 16;        // k-tile
constexpr int GLD = GK + 4;   // smem row stride in doubles: == 4 (mod 16) -> conflict-free fragment loads
constexpr int G_THREADS = 256;

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src, bool valid) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    int sz = valid ? 16 : 0; /* src-size 0 => zero fill */
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem_src), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma8x8x4(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// tile list: for every problem the upper-triangular (bi <= bj) tiles; a prefix table maps blockIdx.x -> problem
__global__ void __launch_bounds__(G_THREADS)
corrdist_kernel(const GemmProb *__restrict__ probs, const int *__restrict__ tile_prefix, int nprob, int ldu) {
    extern __shared__ __align__(16) double gsm[];
    // locate the problem of this CTA (binary search over the prefix sums of tile counts)
    int lo = 0, hi = nprob - 1;
    const int bid = blockIdx.x;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (tile_prefix[mid + 1] > bid) hi = mid;
        else lo = mid + 1;
    }
    const GemmProb P = probs[lo];
    const int n = P.n;
    const int nt = (n + GT - 1) / GT;
    int t = bid - tile_prefix[lo];
    // unrank t -> (bi, bj) with bi <= bj, row-major over the upper triangle
    int bi = 0;
    while (t >= nt - bi) { t -= nt - bi; bi++; }
    const int bj = bi + t;
    const int a0 = bi * GT, b0 = bj * GT;

    double *As = gsm;                       // [2][GT][GLD]
    double *Bs = gsm + 2 * GT * GLD;        // [2][GT][GLD]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wr = warp >> 1, wc = warp & 1;   // warp tile: rows wr*32 .. +32, cols wc*64 .. +64
    const double *__restrict__ U = P.U;

    // each thread copies 4 x 16 bytes per operand per k-tile: tile = 128 rows x 16 doubles = 1024 x 16B chunks
    auto load_tile = [&](int stage, int k0) {
        double *as = As + stage * GT * GLD;
        double *bs = Bs + stage * GT * GLD;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            int chunk = tid + i * G_THREADS;   // 0..1023
            int r = chunk >> 3, cq = (chunk & 7) * 2;  // row, first double of the 16B chunk
            int ra = a0 + r, rb = b0 + r;
            cp_async16(as + r * GLD + cq, U + (size_t)(ra < n ? ra : 0) * ldu + k0 + cq, ra < n);
            cp_async16(bs + r * GLD + cq, U + (size_t)(rb < n ? rb : 0) * ldu + k0 + cq, rb < n);
        }
    };

    double acc[4][8][2];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

    const int nk = ldu / GK;
    load_tile(0, 0);
    cp_async_commit();
    for (int kt = 0; kt < nk; kt++) {
        if (kt + 1 < nk) load_tile((kt + 1) & 1, (kt + 1) * GK);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        const double *as = As + (kt & 1) * GT * GLD + (wr * 32 + (lane >> 2)) * GLD + (lane & 3);
        const double *bs = Bs + (kt & 1) * GT * GLD + (wc * 64 + (lane >> 2)) * GLD + (lane & 3);
#pragma unroll
        for (int k4 = 0; k4 < GK; k4 += 4) {
            double a[4], b[8];
#pragma unroll
            for (int i = 0; i < 4; i++) a[i] = as[i * 8 * GLD + k4];
#pragma unroll
            for (int j = 0; j < 8; j++) b[j] = bs[j * 8 * GLD + k4];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 8; j++) dmma8x8x4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        __syncthreads();
    }

    // epilogue: d = 1 - clamp(r); exact zero on the diagonal; mirror the tile
    double *__restrict__ D = P.D;
    double *__restrict__ Dw = P.Dw;
    const int ld = P.ld;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int r = a0 + wr * 32 + i * 8 + (lane >> 2);
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int cc = b0 + wc * 64 + j * 8 + (lane & 3) * 2;
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int col = cc + e;
                if (r < n && col < n) {
                    double v = acc[i][j][e];
                    v = fmin(1.0, fmax(-1.0, v));
                    double d = (r == col) ? 0.0 : 1.0 - v;
                    if (bi != bj || col >= r) {
                        D[(size_t)r * ld + col] = d;
                        if (Dw) Dw[(size_t)r * ld + col] = d;
                        if (col != r) {
                            D[(size_t)col * ld + r] = d;
                            if (Dw) Dw[(size_t)col * ld + r] = d;
                        }
                    }
                }
            }
        }
    }
}

int launch_corrdist_batched(sharp_ctx *c, const GemmProb *probs_dev, const int *tile_prefix_dev, int nprob,
                            int total_tiles, int ldu) {
    if (nprob <= 0 || total_tiles <= 0) return 0;
    size_t smem = (size_t)4 * GT * GLD * sizeof(double);
    SHARP_SMEM_OPTIN_ONCE((corrdist_kernel), c->device);
    prof_begin(c, KID_CORRDIST);
    corrdist_kernel<<<total_tiles, G_THREADS, smem, c->stream>>>(probs_dev, tile_prefix_dev, nprob, ldu);
    prof_end(c);
    SHARP_CUDA(cudaGetLastError());
    return 0;
}

int corrdist_tiles(int n) {
    int nt = (n + GT - 1) / GT;
    return nt * (nt + 1) / 2;
}

}  // namespace sharp
