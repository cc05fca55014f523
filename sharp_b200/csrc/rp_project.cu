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
#include "rp_common.cuh"

namespace sharp {

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

// =====================================================================================================
// K1, fixed-point variant (the one the pipeline uses whenever K*p <= 32700)
// =====================================================================================================
// The scatter above is bound by its dependent shared-memory read-modify-write chain: one warp per cell, at most ten
// cells per SM (their fp64 accumulators fill shared memory), every gene's add waiting for the previous one.  This
// variant makes the accumulation ORDER-INDEPENDENT, which removes the chain and the one-warp-per-cell limit:
//   * every transformed value v is converted to a 64-bit fixed-point integer q = rint(v * 2^f); f is chosen per cell
//     from max |v| so that |q| < 2^(62 - cb) where 2^cb exceeds the largest number of terms any output can receive
//     (the longest ranM column) -- sums cannot overflow, and the quantum 2^-f is below the fp64 rounding unit of the
//     cell's largest value (the integer sum is EXACT; the only roundings are v -> q and the final int64 -> double,
//     so the result is at least as accurate as the reference's fp64 running sum and bit-reproducible);
//   * the accumulators are two 32-bit limbs per output in shared memory, updated with native shared-memory integer
//     atomics: the atomic add on the low limb returns the old value, from which the carry into the high limb follows;
//   * COUNT CLASSES.  Single-cell counts are small integers: most non-zeros of a cell are 1, 2, 3 or 4, and all
//     non-zeros with the same raw value share the same transformed value.  For those the kernel does not add q at
//     all: it COUNTS, per output, how many +entries and how many -entries each class contributed (8-bit fields, four
//     per 32-bit word, one non-returning-cost atomic per entry instead of two dependent ones and no log per non-zero),
//     and the epilogue adds q_class * (n+ - n-) to the limbs -- the same exact integer sum, so the result is
//     bit-identical to adding every term.  Fields cannot overflow: a field counts entries of ONE ranM column, and the
//     longest column has < 2^FB entries (FB = 16 with two classes when a column is longer than 255);
//   * lanes map to GENES (32 non-zeros of the cell at a time): a lane fetches its gene's padded entry list with
//     16-byte loads and issues its adds without waiting for anybody; eight warps share one cell, five cells per SM.
//     Padding entries point at 32 dummy outputs behind the real ones, so the inner loop has no per-entry branch, and
//     all shared-memory addresses are 32-bit shared-window addresses computed once (atom.shared / red.shared in PTX).
struct RpFxArgs {
    RpArgs a;
    const uint32_t *vecptr;
    const uint4 *entvec;
    int cb;     // bits of headroom for the number of terms per output
    int kpd;    // words per accumulator array: K*p rounded up to 32, plus 32 dummy outputs for the padding entries
};

constexpr int RPF_THREADS = 256;
constexpr int RPF_WARPS = RPF_THREADS / 32;
constexpr int RPF_STAGE = 64;  // per-warp compaction buffer of the dense path
constexpr int RPF_MAXCLS = 4;

// what one lane adds for its gene: `base` = shared address of the array its first atomic goes to (a class counter array
// or the low limbs), apos / aneg = the addend for a + / - entry; generic lanes also carry the high words
struct RpLane {
    uint32_t base, apos, aneg, hpos, hneg, gen;
};

// four entries (two packed words): all first atomics are issued before the first carry is consumed
// PH2: 0 = no lane carries high words (class lanes only), 1 = some may (warp-uniform `anygen`), 2 = generic lanes only
template <int PH2>
__device__ __forceinline__ void rpf_add4(const RpLane &L, uint32_t w0, uint32_t w1, uint32_t hoff, bool anygen) {
    uint32_t ad[4], av[4], old[4];
    ad[0] = L.base + ((w0 & 0x7fffu) << 2);
    ad[1] = L.base + ((w0 >> 14) & 0x1fffcu);
    ad[2] = L.base + ((w1 & 0x7fffu) << 2);
    ad[3] = L.base + ((w1 >> 14) & 0x1fffcu);
    av[0] = (w0 & 0x8000u) ? L.aneg : L.apos;
    av[1] = ((int32_t)w0 < 0) ? L.aneg : L.apos;
    av[2] = (w1 & 0x8000u) ? L.aneg : L.apos;
    av[3] = ((int32_t)w1 < 0) ? L.aneg : L.apos;
#pragma unroll
    for (int e = 0; e < 4; e++) old[e] = atoms_add(ad[e], av[e]);
    if (PH2 == 2 || (PH2 == 1 && anygen)) { /* warp-uniform */
        const uint32_t hv[4] = {(w0 & 0x8000u) ? L.hneg : L.hpos, ((int32_t)w0 < 0) ? L.hneg : L.hpos,
                                (w1 & 0x8000u) ? L.hneg : L.hpos, ((int32_t)w1 < 0) ? L.hneg : L.hpos};
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const uint32_t carry = ((uint32_t)(old[e] + av[e]) < av[e]) ? 1u : 0u;
            reds_add_if(ad[e] + hoff, hv[e] + carry, L.gen);
        }
    }
}

// count class of a raw value: 0 .. NCLS-1 for the integers 1 .. NCLS, else -1 (generic)
template <int FB>
__device__ __forceinline__ int rpf_class(double x, bool valid) {
    constexpr int NCLS = (FB == 8) ? 4 : 2;
    if (!valid) return -1;
    const int xi = __double2int_rz(x);
    return (xi >= 1 && xi <= NCLS && (double)xi == x) ? xi - 1 : -1;
}

// MODE 0: class and generic lanes mixed (dense input); 1: class lanes only (lanes with cls < 0 idle); 2: generic lanes only
template <int VEC, int FB, int MODE>
__device__ __forceinline__ void rpf_chunk(const RpFxArgs &A, uint32_t s_base, int gi, double x, bool valid, int cls, double cs,
                                          double qscale, int okmask, int *bad) {
    constexpr int CPW = 16 / FB;              /* classes per 32-bit word (two fields each) */
    constexpr int PH2 = MODE == 1 ? 0 : (MODE == 2 ? 2 : 1);
    const uint32_t arr = (uint32_t)A.kpd * 4u;
    if (MODE == 2) cls = -1;
    if (MODE == 1) valid = valid && cls >= 0;
    RpLane L;
    L.gen = (valid && cls < 0) ? 1u : 0u;
    long long q = 0;
    if (L.gen) {
        double v = rp_transform(x, cs, A.a.normalize, A.a.norm_mul, A.a.logkind);
        if (!isfinite(v)) { *bad = 1; v = 0.0; } /* the whole row becomes NaN, like the reference's NaN propagation */
        q = __double2ll_rn(v * qscale);
    } else if (cls >= 0 && !((okmask >> cls) & 1)) *bad = 1;
    const unsigned long long nq = 0ull - (unsigned long long)q;
    L.base = s_base + 2u * arr; /* low limbs */
    L.apos = (uint32_t)q;
    L.hpos = (uint32_t)((unsigned long long)q >> 32);
    L.aneg = (uint32_t)nq;
    L.hneg = (uint32_t)(nq >> 32);
    if (cls >= 0) {
        L.base = s_base + (uint32_t)(cls / CPW) * arr;
        L.apos = 1u << ((cls % CPW) * 2 * FB);
        L.aneg = L.apos << FB;
    }
    uint32_t v0 = 0, nv = 0;
    if (cls >= 0 || q != 0) {
        v0 = __ldg(A.vecptr + gi);
        nv = __ldg(A.vecptr + gi + 1) - v0;
    }
    uint4 w[VEC];
#pragma unroll
    for (int u = 0; u < VEC; u++)
        if ((uint32_t)u < nv) w[u] = __ldg(A.entvec + v0 + u);
    const bool anygen = (MODE == 0) ? __any_sync(0xffffffffu, L.gen && nv > 0) : (MODE == 2);
#pragma unroll
    for (int u = 0; u < VEC; u++)
        if ((uint32_t)u < nv) {
            rpf_add4<PH2>(L, w[u].x, w[u].y, arr, anygen);
            rpf_add4<PH2>(L, w[u].z, w[u].w, arr, anygen);
        }
    if (nv > VEC) { /* longer lists: the next vector is requested before the current one is applied */
        uint4 t = __ldg(A.entvec + v0 + VEC);
        for (uint32_t u = VEC; u < nv; u++) {
            uint4 nx = t;
            if (u + 1 < nv) nx = __ldg(A.entvec + v0 + u + 1);
            rpf_add4<PH2>(L, t.x, t.y, arr, anygen);
            rpf_add4<PH2>(L, t.z, t.w, arr, anygen);
            t = nx;
        }
    }
}

__device__ __forceinline__ double rpf_block_max(double v, double *red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = red[0];
#pragma unroll
    for (int w = 1; w < RPF_WARPS; w++) r = fmax(r, red[w]);
    __syncthreads();
    return r;
}

// dynamic shared memory: [4][kpd] words (class counters 0, class counters 1, low limbs, high limbs), then, for dense
// input only, the per-warp compaction buffers
template <int VEC, int FB, bool DENSE>
__global__ void __launch_bounds__(RPF_THREADS, 5) rp_project_fx_kernel(RpFxArgs A) {
    extern __shared__ __align__(16) uint32_t fsm[];
    __shared__ double red[RPF_WARPS];
    __shared__ long long s_qc[RPF_MAXCLS];
    __shared__ int s_bad, s_ok;
    constexpr int NCLS = (FB == 8) ? 4 : 2;
    constexpr int CPW = 16 / FB;
    constexpr uint32_t FMASK = (1u << FB) - 1u;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int KP = A.a.KP, kpd = A.kpd;
    const uint32_t s_base = (uint32_t)__cvta_generic_to_shared(fsm);
    uint32_t *cw = fsm, *lo = fsm + 2 * kpd, *hi = fsm + 3 * kpd;
    int *sgi = reinterpret_cast<int *>(fsm + 4 * kpd) + warp * RPF_STAGE;
    double *sx = reinterpret_cast<double *>(fsm + 4 * kpd + RPF_WARPS * RPF_STAGE) + warp * RPF_STAGE;
    for (int i = tid; i < 4 * kpd; i += RPF_THREADS) fsm[i] = 0u;
    if (tid == 0) s_bad = 0;
    __syncthreads();

    for (int64_t pos = blockIdx.x; pos < A.a.ncell; pos += gridDim.x) {
        const int64_t src = A.a.cells ? A.a.cells[pos] : pos;
        const double cs = A.a.normalize ? A.a.colsum[src] : 1.0;
        const double *dcol = DENSE ? A.a.dense + src * A.a.m : nullptr;
        int64_t q0 = 0, q1 = 0;
        if (!DENSE) { q0 = A.a.colptr[src]; q1 = A.a.colptr[src + 1]; }
        // ---- pass 1: range of the raw values -> bound on |v| (the transform is monotone) -> fixed-point scale ----
        double xmax = -SHARP_INF, xmin = SHARP_INF;
        if (DENSE) {
            for (int g = tid; g < A.a.m; g += RPF_THREADS) { const double x = dcol[g]; xmax = fmax(xmax, x); xmin = fmin(xmin, x); }
        } else {
            for (int64_t q = q0 + tid; q < q1; q += RPF_THREADS) { const double x = A.a.val[q]; xmax = fmax(xmax, x); xmin = fmin(xmin, x); }
        }
        xmax = rpf_block_max(xmax, red);
        xmin = -rpf_block_max(-xmin, red);
        double bound = 0.0;
        if (xmax >= xmin) { /* at least one value */
            const double b1 = fabs(rp_transform(xmax, cs, A.a.normalize, A.a.norm_mul, A.a.logkind));
            const double b2 = fabs(rp_transform(xmin, cs, A.a.normalize, A.a.norm_mul, A.a.logkind));
            bound = fmax(isfinite(b1) ? b1 : 0.0, isfinite(b2) ? b2 : 0.0);
        }
        const int e = (bound > 0.0) ? ilogb(bound) + 1 : 0;
        const int fb = 62 - A.cb - e;
        const double qscale = ldexp(1.0, fb);
        if (tid == 0) s_ok = 0;
        __syncthreads();
        if (tid < NCLS) { /* the fixed-point value of every count class (a class that does not occur is multiplied by 0) */
            const double v = rp_transform((double)(tid + 1), cs, A.a.normalize, A.a.norm_mul, A.a.logkind);
            const bool ok = isfinite(v) && fabs(v) <= ldexp(bound, 3);
            s_qc[tid] = ok ? __double2ll_rn(v * qscale) : 0ll;
            if (ok) atomicOr(&s_ok, 1 << tid);
        }
        __syncthreads();
        const int okmask = s_ok;
        // ---- pass 2: scatter ----
        if (DENSE) {
            int cnt = 0;
            for (int c0 = warp * 32; c0 < A.a.m; c0 += RPF_THREADS) {
                const int gi = c0 + lane;
                const double x = (gi < A.a.m) ? dcol[gi] : 0.0;
                const unsigned mask = __ballot_sync(0xffffffffu, x != 0.0);
                if (x != 0.0) {
                    const int at = cnt + __popc(mask & ((1u << lane) - 1u));
                    sgi[at] = gi;
                    sx[at] = x;
                }
                cnt += __popc(mask);
                __syncwarp();
                if (cnt >= 32) {
                    rpf_chunk<VEC, FB, 0>(A, s_base, sgi[lane], sx[lane], true, rpf_class<FB>(sx[lane], true), cs, qscale, okmask, &s_bad);
                    __syncwarp();
                    int tg = 0;
                    double tx = 0.0;
                    if (lane < cnt - 32) { tg = sgi[32 + lane]; tx = sx[32 + lane]; }
                    __syncwarp();
                    if (lane < cnt - 32) { sgi[lane] = tg; sx[lane] = tx; }
                    cnt -= 32;
                    __syncwarp();
                }
            }
            if (cnt > 0) {
                const double tx = lane < cnt ? sx[lane] : 0.0;
                rpf_chunk<VEC, FB, 0>(A, s_base, lane < cnt ? sgi[lane] : 0, tx, lane < cnt, rpf_class<FB>(tx, lane < cnt), cs, qscale, okmask, &s_bad);
            }
        } else {
            /* Class lanes need one atomic per entry, generic lanes two and a carry: a warp that mixes them pays for both.
               Generic non-zeros are therefore set aside (their position in the column, 4 bytes each, in a per-warp list)
               and handled 32 at a time by all-generic passes; the rest of the warp's work is class-only. */
            uint32_t *stq = reinterpret_cast<uint32_t *>(fsm + 4 * kpd) + warp * RPF_STAGE;
            int scnt = 0;
            int64_t q = q0 + warp * 32;
            bool valid = q + lane < q1;
            int gi = valid ? A.a.rowidx[q + lane] : 0;
            double x = valid ? A.a.val[q + lane] : 0.0;
            while (q < q1) { /* the next 32 non-zeros are requested before this chunk's atomics are issued */
                const int64_t qn = q + RPF_THREADS;
                const bool nvalid = qn + lane < q1;
                const int ngi = nvalid ? A.a.rowidx[qn + lane] : 0;
                const double nx = nvalid ? A.a.val[qn + lane] : 0.0;
                const int cls = rpf_class<FB>(x, valid);
                const bool gen = valid && cls < 0 && x != 0.0; /* explicit zeros contribute nothing */
                const unsigned gmask = __ballot_sync(0xffffffffu, gen);
                if (gen) stq[scnt + __popc(gmask & ((1u << lane) - 1u))] = (uint32_t)(q + lane - q0);
                scnt += __popc(gmask);
                rpf_chunk<VEC, FB, 1>(A, s_base, gi, x, valid, cls, cs, qscale, okmask, &s_bad);
                __syncwarp();
                if (scnt >= 32) {
                    const int64_t at = q0 + stq[lane];
                    rpf_chunk<VEC, FB, 2>(A, s_base, A.a.rowidx[at], A.a.val[at], true, -1, cs, qscale, okmask, &s_bad);
                    __syncwarp();
                    const uint32_t t = (lane < scnt - 32) ? stq[32 + lane] : 0u;
                    __syncwarp();
                    if (lane < scnt - 32) stq[lane] = t;
                    scnt -= 32;
                    __syncwarp();
                }
                q = qn; valid = nvalid; gi = ngi; x = nx;
            }
            if (scnt > 0) {
                const bool v = lane < scnt;
                const int64_t at = q0 + (v ? stq[lane] : 0u);
                rpf_chunk<VEC, FB, 2>(A, s_base, v ? A.a.rowidx[at] : 0, v ? A.a.val[at] : 0.0, v, -1, cs, qscale, okmask, &s_bad);
            }
        }
        __syncthreads();
        // ---- output: the exact integer sum -> double (one rounding), times the common factor; reset the words ----
        const bool bad = s_bad != 0;
        const double unscale = ldexp(1.0, -fb);
        long long qc[NCLS];
#pragma unroll
        for (int c = 0; c < NCLS; c++) qc[c] = s_qc[c];
        for (int k = 0, i0 = 0; k < A.a.K; k++, i0 += A.a.p)
        for (int j = tid; j < A.a.p; j += RPF_THREADS) { /* (member, column) without an integer division per output */
            const int i = i0 + j;
            long long tot = (long long)(((unsigned long long)hi[i] << 32) | (unsigned long long)lo[i]);
            lo[i] = 0u;
            hi[i] = 0u;
#pragma unroll
            for (int a = 0; a < NCLS / CPW; a++) {
                const uint32_t w = cw[a * kpd + i];
                cw[a * kpd + i] = 0u;
#pragma unroll
                for (int f = 0; f < CPW; f++) {
                    const int np = (int)((w >> (f * 2 * FB)) & FMASK), nn = (int)((w >> (f * 2 * FB + FB)) & FMASK);
                    tot += qc[a * CPW + f] * (long long)(np - nn);
                }
            }
            double r = __dmul_rn(__dmul_rn((double)tot, unscale), A.a.scale);
            if (A.a.round_digits >= 0) r = rp_round(r, A.a.round_digits);
            if (bad) r = __longlong_as_double(0x7ff8000000000000LL);
            A.a.out[((size_t)k * A.a.ncell + pos) * A.a.p + j] = r;
        }
        __syncthreads();
        if (tid == 0) s_bad = 0;
        /* the dummy outputs collect the padding entries; they are never read and only have to be cleared often enough
           not to matter -- a wrapped dummy word is harmless */
    }
}

template <int VEC, int FB, bool DENSE>
static int launch_fx2(sharp_ctx *c, const RpFxArgs &A, int64_t ncell, size_t smem) {
    SHARP_SMEM_OPTIN_ONCE((rp_project_fx_kernel<VEC, FB, DENSE>), c->device);
    /* a persistent grid: exactly the CTAs that are resident at once (no tail wave); the occupancy query is a driver call,
       made once per shared-memory size and host thread */
    static thread_local size_t q_smem = 0;
    static thread_local int q_per_sm = 0;
    if (q_per_sm == 0 || q_smem != smem) {
        SHARP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q_per_sm, rp_project_fx_kernel<VEC, FB, DENSE>, RPF_THREADS, smem));
        q_smem = smem;
    }
    const int per_sm = q_per_sm;
    const int grid = (int)std::min<int64_t>(ncell, (int64_t)c->sm_count * std::max(1, per_sm));
    rp_project_fx_kernel<VEC, FB, DENSE><<<grid, RPF_THREADS, smem, c->stream>>>(A);
    return 0;
}
template <int VEC, int FB>
static int launch_fx(sharp_ctx *c, const RpFxArgs &A, int64_t ncell, size_t smem) {
    return A.a.dense ? launch_fx2<VEC, FB, true>(c, A, ncell, smem) : launch_fx2<VEC, FB, false>(c, A, ncell, smem);
}
template <int FB>
static int launch_fx_vec(sharp_ctx *c, const RpFxArgs &A, int vec_per_gene, int64_t grid, size_t smem) {
    /* two vectors (16 entries) are fetched up front, the tail of longer lists one vector ahead: more in registers
       spills at the 48 registers that keep five cells per SM resident */
    (void)vec_per_gene;
    return launch_fx<2, FB>(c, A, grid, smem);
}

struct RpArgs;
int launch_rp_project_v3(sharp_ctx *c, const RpArgs &Ain, const sharp_rm_dev &rm, bool staged, double *colsum_out, void *info_ws);
size_t rp_cellinfo_bytes(int64_t n);

int launch_rp_project(sharp_ctx *c, const sharp_expr_dev &e, const int64_t *cells_dev, int64_t ncell,
                      double *colsum_dev, bool colsum_ready, int normalize, double norm_mul, int logkind, int round_digits,
                      const sharp_rm_dev &rm, double *out) {
    if (ncell <= 0) return 0;
    if (e.m != rm.m) return set_error(SHARP_E_ARG, "rp_project: expression has %d genes but ranM has %d rows", e.m, rm.m);
    if (normalize && !colsum_dev) return set_error(SHARP_E_ARG, "rp_project: normalisation needs the column sums");
    RpArgs A;
    A.m = e.m; A.n = e.n; A.dense = e.dense; A.colptr = e.colptr; A.rowidx = e.rowidx; A.val = e.val;
    A.cells = cells_dev; A.ncell = ncell; A.colsum = colsum_dev; A.normalize = normalize ? 1 : 0;
    A.norm_mul = norm_mul; A.logkind = logkind; A.round_digits = round_digits;
    A.p = rm.p; A.K = rm.K; A.KP = rm.K * rm.p;
    A.scale = (1.0 / sqrt((double)rm.p)) * rm.mag;   /* entry of 1/sqrt(p) * t(rM) */
    A.rowptr = rm.rowptr; A.ent16 = rm.ent16; A.ent32 = rm.ent32;
    A.out = out;
    if ((c->rp_variant == 0 || c->rp_variant == 3) && !e.dense && rm.rec) { /* record-gather variant (CSC input) */
        SHARP_TRY(c->ws[WS_CELLINFO_SLOT].reserve(rp_cellinfo_bytes(e.n)));
        RpArgs B = A;
        B.normalize = !normalize ? 0 : (colsum_ready ? 1 : 2); /* 2: the pre-pass computes (and stores) the sums */
        const int rc = launch_rp_project_v3(c, B, rm, c->rp_variant == 3, (normalize && !colsum_ready) ? colsum_dev : nullptr,
                                            c->ws[WS_CELLINFO_SLOT].ptr);
        if (rc <= 0) return rc;
        /* rc == 1: does not apply (very long ranM columns, huge K*p): the kernels below */
    }
    if (normalize && !colsum_ready) SHARP_TRY(launch_colsum(c, e, colsum_dev));
    if (rm.entvec && c->rp_variant != 1 && rm.max_col_nnz < 65536) { /* fixed-point variant */
        RpFxArgs F;
        F.a = A;
        F.vecptr = rm.vecptr;
        F.entvec = rm.entvec;
        F.cb = 1;
        while ((1 << F.cb) <= rm.max_col_nnz) F.cb++;
        F.kpd = rm.kpd;
        const size_t fsmem = (size_t)4 * F.kpd * 4 + (size_t)RPF_WARPS * RPF_STAGE * (e.dense ? 12 : 4);
        if (fsmem <= 200 * 1024) {
            prof_begin(c, KID_RP_PROJECT);
            const int rc = rm.max_col_nnz <= 255 ? launch_fx_vec<8>(c, F, rm.vec_per_gene, ncell, fsmem)
                                                 : launch_fx_vec<16>(c, F, rm.vec_per_gene, ncell, fsmem);
            prof_end(c);
            SHARP_TRY(rc);
            SHARP_CUDA(cudaGetLastError());
            return 0;
        }
    }
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
        SHARP_SMEM_OPTIN_ONCE((rp_project_kernel<true>), c->device);
        rp_project_kernel<true><<<grid, W * 32, smem, c->stream>>>(A, W);
    } else {
        SHARP_SMEM_OPTIN_ONCE((rp_project_kernel<false>), c->device);
        rp_project_kernel<false><<<grid, W * 32, smem, c->stream>>>(A, W);
    }
    prof_end(c);
    SHARP_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace sharp
