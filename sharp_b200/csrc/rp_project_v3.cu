// rp_project_v3.cu -- K1, record-gather variant of the fixed-point projection scatter (CSC input; the default).
//
// Replaces  1/sqrt(p) * t(rM[[k]]) %*% log2(E[, tind] + 1)  for all K members at once, with the CPM normalisation
// fused into the load (R/SHARP.R:567-585, R/RPmat.R:32, R/SHARP_unlimited2.R:388-410, R/SHARP.R:113) -- the same exact
// 64-bit fixed-point sums as rp_project_fx_kernel (rp_project.cu explains the arithmetic: every transformed value is
// rounded ONCE to an integer multiple of a per-cell quantum, integer sums are exact and order-independent, non-zeros
// equal to 1..4 are COUNTED per output instead of added).  What changed is how the ranM entries reach the lanes -- the
// r1 kernel was bound by its gathers, not by HBM and not by the shared-memory atomics (ncu: 15 sectors per load request,
// the L1 data pipe the busiest unit, two dependent L2 round trips per gene):
//   * RECORDS.  Every gene owns one fixed-size, 16*RV-byte aligned record of RV 16-byte vectors (no pointer array, no
//     dependent lookup: the address follows from the gene index).  A vector is {count, 7 entries}; the gene's entries are
//     dealt round-robin over its vectors, so the RV lanes that fetch one record -- ONE coalesced request per gene, all
//     lanes of a gene in the same line -- each get an equal share of the adds.  An entry is a byte offset into the PAIR
//     of count arrays [+ counts | - counts]: col * 4 for a +entry, col * 4 + (bytes of one array) for a -entry, so the
//     sign costs no instruction on the count path: address = base + entry, addend = 1 << (8 * class).
//     Records are sized so that < 0.2 % of the genes overflow; those are served from the gene-major CSR lists.
//   * COUNT WORDS: one word per output and sign, four 8-bit fields (classes 1..4); a field counts the entries of one
//     ranM column with that sign that met that class in this cell, so it cannot overflow while the longest column has
//     <= 255 entries (m up to ~65 000 genes; longer columns take the r1 kernel's 16-bit path).
//   * PER-CELL RECORD.  colSums(x), the value range, the fixed-point scale and the four class values are produced by
//     one streaming pre-pass (cellprep_kernel, which replaces colsum_kernel on this path), so the scatter kernel makes a
//     single pass over a cell's non-zeros and needs two barriers per cell instead of nine.
//   * STAGED variant (TMA = true): the cell's (rowidx, val) segments are copied into shared memory by cp.async.bulk
//     (1D TMA, completion on an mbarrier) ONE CELL AHEAD of the scatter, double buffered; the chunk loop then reads
//     shared memory only.  Selected with sharp_ctx_set_rp_variant(ctx, 3); profiles/ holds the ncu comparison.
#include "rp_common.cuh"

namespace sharp {

struct __align__(16) RpCellInfo {
    double cs;          // divisor of the normalisation (1 when none)
    double qscale;      // 2^f
    double unscale;     // 2^-f
    long long qc[4];    // fixed-point value of the count classes 1..4
    int okmask;         // bit c: class c + 1 has a finite, in-range value
    int pad;
};

struct RpV3Args {
    RpArgs a;
    const uint4 *rec;         // [m * RV]
    const RpCellInfo *info;   // [n], per SOURCE column
    int kpd;                  // words per accumulator array
    uint32_t dummy;           // byte offset of a dummy word (what the unused slots of a record point at)
    int cb;                   // bits of headroom for the number of terms per output
};

// ---- pre-pass: column sum, value range -> fixed-point scale and class values (one warp per cell) ---------------
__global__ void __launch_bounds__(256)
cellprep_kernel(int64_t n, const int64_t *__restrict__ colptr, const double *__restrict__ val, const double *__restrict__ colsum_in,
                int normalize, double norm_mul, int logkind, int cb, double *__restrict__ colsum_out, RpCellInfo *__restrict__ info) {
    const int lane = threadIdx.x & 31;
    const int64_t c = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= n) return;
    double s = 0.0, xmax = -SHARP_INF, xmin = SHARP_INF;
    for (int64_t q = colptr[c] + lane; q < colptr[c + 1]; q += 32) {
        const double x = val[q];
        s += x;
        xmax = fmax(xmax, x);
        xmin = fmin(xmin, x);
    }
    s = warp_sum(s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        xmax = fmax(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
        xmin = fmin(xmin, __shfl_xor_sync(0xffffffffu, xmin, o));
    }
    if (colsum_out && lane == 0) colsum_out[c] = s;
    const double cs = normalize == 2 ? s : (normalize == 1 ? colsum_in[c] : 1.0);
    double bound = 0.0;
    if (xmax >= xmin) { /* at least one value; the transform is monotone */
        const double b1 = fabs(rp_transform(xmax, cs, normalize, norm_mul, logkind));
        const double b2 = fabs(rp_transform(xmin, cs, normalize, norm_mul, logkind));
        bound = fmax(isfinite(b1) ? b1 : 0.0, isfinite(b2) ? b2 : 0.0);
    }
    const int e = (bound > 0.0) ? ilogb(bound) + 1 : 0;
    const int fb = 62 - cb - e;
    const double qscale = ldexp(1.0, fb);
    long long qc = 0;
    bool ok = false;
    if (lane < 4) {
        const double v = rp_transform((double)(lane + 1), cs, normalize, norm_mul, logkind);
        ok = isfinite(v) && fabs(v) <= ldexp(bound, 3);
        qc = ok ? __double2ll_rn(v * qscale) : 0ll;
    }
    const unsigned okm = __ballot_sync(0xffffffffu, ok) & 15u;
    const long long q1 = __shfl_sync(0xffffffffu, qc, 1), q2 = __shfl_sync(0xffffffffu, qc, 2), q3 = __shfl_sync(0xffffffffu, qc, 3);
    if (lane == 0) {
        RpCellInfo I;
        I.cs = cs; I.qscale = qscale; I.unscale = ldexp(1.0, -fb);
        I.qc[0] = qc; I.qc[1] = q1; I.qc[2] = q2; I.qc[3] = q3;
        I.okmask = (int)okm; I.pad = 0;
        info[c] = I;
    }
}

// ---- scatter helpers --------------------------------------------------------------------------------------
__device__ __forceinline__ void reds_add(uint32_t addr, uint32_t v) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// the seven entries of one vector (the first half-word is the count)
__device__ __forceinline__ void rpv_entries(const uint4 &w, uint32_t e[7]) {
    e[0] = w.x >> 16;
    e[1] = w.y & 0xffffu; e[2] = w.y >> 16;
    e[3] = w.z & 0xffffu; e[4] = w.z >> 16;
    e[5] = w.w & 0xffffu; e[6] = w.w >> 16;
}

// count-class lanes: one non-returning atomic per entry; the entry IS the byte offset into [+ counts | - counts].
// UNC (the default): the unused slots of a record point at dummy words behind the + counts (spread over the banks), so a
// lane issues its seven atomics in a straight line -- 16 instructions per vector instead of ~44 with a branch per slot
// (ptxas turns predicated shared atomics into branches); measured 71 vs 80 ms per 1.3 M-cell step on B200.
template <bool UNC>
__device__ __forceinline__ void rpv_scatter_class(const uint4 &w, uint32_t n, uint32_t base, uint32_t addc) {
    uint32_t e[7];
    rpv_entries(w, e);
    if (UNC) { /* the unused slots of a record point at the dummy words behind the + counts: seven straight-line atomics */
#pragma unroll
        for (int s = 0; s < 7; s++) reds_add(base + e[s], addc);
        return;
    }
#pragma unroll
    for (int s = 0; s < 7; s++)
        if ((uint32_t)s < n) reds_add(base + e[s], addc);
}

// generic lanes: 64-bit add as two 32-bit limbs (the returning atomic on the low limb yields the carry)
__device__ __forceinline__ void rpv_scatter_generic(const uint4 &w, uint32_t n, uint32_t lo_base, uint32_t arr,
                                                    uint32_t apos, uint32_t hpos, uint32_t aneg, uint32_t hneg) {
    uint32_t e[7];
    rpv_entries(w, e);
#pragma unroll
    for (int s = 0; s < 7; s++) {
        if ((uint32_t)s < n) {
            const bool neg = e[s] >= arr;
            const uint32_t addr = lo_base + (neg ? e[s] - arr : e[s]);
            const uint32_t av = neg ? aneg : apos;
            const uint32_t old = atoms_add(addr, av);
            const uint32_t hv = (neg ? hneg : hpos) + (((uint32_t)(old + av) < av) ? 1u : 0u);
            if (hv) reds_add(addr + arr, hv);
        }
    }
}

// genes whose entry list does not fit a record (vector 0 carries the count 0xffff): the gene-major CSR list, the lanes of
// the gene striding over it
template <int RV>
__device__ __forceinline__ void rpv_overflow(const RpV3Args &A, uint32_t g, int v, bool generic, uint32_t base, uint32_t arr,
                                             uint32_t apos, uint32_t hpos, uint32_t aneg, uint32_t hneg) {
    const uint32_t r0 = __ldg(A.a.rowptr + g), r1 = __ldg(A.a.rowptr + g + 1);
    for (uint32_t q = r0 + v; q < r1; q += RV) {
        const unsigned ent = __ldg(A.a.ent16 + q);
        const uint32_t addr = base + ((ent & 0x7fffu) << 2);
        const bool neg = ent & 0x8000u;
        if (!generic) reds_add(addr + (neg ? arr : 0u), apos); /* base = the + count array, apos = 1 << (8 * class) */
        else {
            const uint32_t av = neg ? aneg : apos;
            const uint32_t old = atoms_add(addr, av);
            const uint32_t hv = (neg ? hneg : hpos) + (((uint32_t)(old + av) < av) ? 1u : 0u);
            if (hv) reds_add(addr + arr, hv);
        }
    }
}

// 32 class non-zeros (lane l: gene `gi`, class `cls` in 0..3, or invalid) -> RV lanes per gene, 32 / RV genes per pass
template <int RV, bool UNC>
__device__ __forceinline__ void rpv_class_chunk(const RpV3Args &A, uint32_t s_base, uint32_t arr, int gi, int cls, bool valid, int lane) {
    constexpr int G = 32 / RV;
    constexpr int SB = RV < 2 ? RV : 2; /* record vectors in flight per lane */
    const unsigned packed = valid ? ((unsigned)gi | ((unsigned)cls << 28) | 0x80000000u) : 0u;
    const int v = lane & (RV - 1), grp = lane / RV;
#pragma unroll 1
    for (int u0 = 0; u0 < RV; u0 += SB) {
        uint4 w[SB];
        unsigned pk[SB];
#pragma unroll
        for (int b = 0; b < SB; b++) {
            pk[b] = __shfl_sync(0xffffffffu, packed, (u0 + b) * G + grp);
            w[b] = make_uint4(0u, 0u, 0u, 0u);
            if (pk[b]) w[b] = __ldg(A.rec + (size_t)(pk[b] & 0x0fffffffu) * RV + v);
            else if (UNC) { /* idle lane: its own dummy word */
                const uint32_t d = A.dummy + 4u * (uint32_t)lane;
                w[b] = make_uint4(d << 16, d * 0x10001u, d * 0x10001u, d * 0x10001u);
            }
        }
#pragma unroll
        for (int b = 0; b < SB; b++) {
            uint32_t n = w[b].x & 0xffffu;
            const bool ovf = n == 0xffffu;
            if (ovf) n = 0;
            const uint32_t addc = 1u << ((pk[b] >> 25) & 24u); /* 1 << (8 * class): the class sits in bits 28..29 */
            rpv_scatter_class<UNC>(w[b], n, s_base, addc);
            if (__any_sync(0xffffffffu, ovf)) {
                /* vector 0 of the record carries the marker: tell the gene's other lanes */
                const bool govf = __shfl_sync(0xffffffffu, ovf ? 1 : 0, lane & ~(RV - 1)) != 0;
                if (govf && pk[b]) rpv_overflow<RV>(A, pk[b] & 0x0fffffffu, v, false, s_base, arr, addc, 0u, 0u, 0u);
            }
        }
    }
}

// 32 generic non-zeros (lane l: gene `gi`, fixed-point value q, or invalid)
template <int RV>
__device__ __forceinline__ void rpv_generic_chunk(const RpV3Args &A, uint32_t s_base, uint32_t arr, int gi, long long q, bool valid, int lane) {
    constexpr int G = 32 / RV;
    const unsigned packed = (valid && q != 0) ? ((unsigned)gi | 0x80000000u) : 0u;
    const uint32_t qlo = (uint32_t)q, qhi = (uint32_t)((unsigned long long)q >> 32);
    const int v = lane & (RV - 1), grp = lane / RV;
    const uint32_t lo_base = s_base + 2u * arr;
#pragma unroll 1
    for (int u = 0; u < RV; u++) {
        const int sl = u * G + grp;
        const unsigned pk = __shfl_sync(0xffffffffu, packed, sl);
        const uint32_t apos = __shfl_sync(0xffffffffu, qlo, sl), hpos = __shfl_sync(0xffffffffu, qhi, sl);
        uint4 w = make_uint4(0u, 0u, 0u, 0u);
        if (pk) w = __ldg(A.rec + (size_t)(pk & 0x0fffffffu) * RV + v);
        const unsigned long long nq = 0ull - (((unsigned long long)hpos << 32) | apos);
        const uint32_t aneg = (uint32_t)nq, hneg = (uint32_t)(nq >> 32);
        uint32_t n = w.x & 0xffffu;
        const bool ovf = n == 0xffffu;
        if (ovf) n = 0;
        rpv_scatter_generic(w, n, lo_base, arr, apos, hpos, aneg, hneg);
        if (__any_sync(0xffffffffu, ovf)) {
            const bool govf = __shfl_sync(0xffffffffu, ovf ? 1 : 0, lane & ~(RV - 1)) != 0;
            if (govf && pk) rpv_overflow<RV>(A, pk & 0x0fffffffu, v, true, lo_base, arr, apos, hpos, aneg, hneg);
        }
    }
}

// count class of a raw value: 0..3 for the integers 1..4, else -1
__device__ __forceinline__ int rpv_class(double x) {
    const int xi = __double2int_rz(x);
    return (xi >= 1 && xi <= 4 && (double)xi == x) ? xi - 1 : -1;
}

// ---- mbarrier / bulk-copy helpers of the staged variant ----------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra WAIT_%=;\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

constexpr int RPV_LIST = 64;       // per-warp list of generic non-zeros waiting for an all-generic pass
constexpr int RPV_STAGE_NNZ = 2048; // staged variant: non-zeros of a cell held per buffer (the rest is read from global)

// dynamic shared memory: [4][kpd] words (+ counts, - counts, low limbs, high limbs); per-warp generic
// lists (position in the column); staged variant: 2 buffers of {rowidx[STAGE + 8], val[STAGE + 4]}
template <int RV, int NT, int MINB, bool TMA, bool UNC>
__global__ void __launch_bounds__(NT, MINB) rp_project_v3_kernel(RpV3Args A) {
    extern __shared__ __align__(16) uint32_t vsm[];
    __shared__ int s_badv[2];   /* alternating per cell: the flag of the NEXT cell is cleared inside this cell's epilogue */
    __shared__ __align__(8) unsigned long long s_bar[2];
    constexpr int NWARP = NT / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int kpd = A.kpd;
    const uint32_t arr = (uint32_t)kpd * 4u;
    const uint32_t s_base = (uint32_t)__cvta_generic_to_shared(vsm);
    uint32_t *cw = vsm, *lo = vsm + 2 * kpd, *hi = vsm + 3 * kpd;
    uint32_t *glist = vsm + 4 * kpd + warp * RPV_LIST;
    // staged buffers (16-byte aligned: kpd is a multiple of 32)
    unsigned char *stage = reinterpret_cast<unsigned char *>(vsm + 4 * kpd + NWARP * RPV_LIST);
    constexpr int ST_IDX_BYTES = (RPV_STAGE_NNZ + 8) * 4, ST_VAL_BYTES = (RPV_STAGE_NNZ + 4) * 8;
    constexpr int ST_BYTES = ST_IDX_BYTES + ST_VAL_BYTES;
    for (int i = tid; i < 4 * kpd; i += NT) vsm[i] = 0u;
    if (tid == 0) {
        s_badv[0] = s_badv[1] = 0;
        if (TMA) {
            mbar_init((uint32_t)__cvta_generic_to_shared(&s_bar[0]), 1);
            mbar_init((uint32_t)__cvta_generic_to_shared(&s_bar[1]), 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    __syncthreads();

    // staged variant: issue the copies of one cell (thread 0).  Sources are widened to 16-byte boundaries; the head
    // offsets are recomputed by the consumers from the same colptr values.
    auto issue = [&](int64_t pos, int buf) {
        const int64_t src = A.a.cells ? A.a.cells[pos] : pos;
        const int64_t q0 = A.a.colptr[src], q1 = A.a.colptr[src + 1];
        const int64_t ns = min((int64_t)RPV_STAGE_NNZ, q1 - q0);
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_bar[buf]);
        const int64_t i0 = q0 & ~(int64_t)3, i1 = (q0 + ns + 3) & ~(int64_t)3;   /* rowidx: 4 per 16 bytes */
        const int64_t v0 = q0 & ~(int64_t)1, v1 = (q0 + ns + 1) & ~(int64_t)1;   /* val: 2 per 16 bytes */
        const uint32_t bi = (uint32_t)(i1 - i0) * 4u, bv = (uint32_t)(v1 - v0) * 8u;
        mbar_expect_tx(bar, bi + bv);
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(stage + (size_t)buf * ST_BYTES);
        if (bi) bulk_g2s(dst, A.a.rowidx + i0, bi, bar);
        if (bv) bulk_g2s(dst + ST_IDX_BYTES, A.a.val + v0, bv, bar);
    };
    int it = 0;
    if (TMA && tid == 0 && (int64_t)blockIdx.x < A.a.ncell) issue(blockIdx.x, 0);

    for (int64_t pos = blockIdx.x; pos < A.a.ncell; pos += gridDim.x, it++) {
        const int64_t src = A.a.cells ? A.a.cells[pos] : pos;
        const int64_t q0 = A.a.colptr[src], q1 = A.a.colptr[src + 1];
        const RpCellInfo *I = A.info + src;
        const double cs = I->cs, qscale = I->qscale;
        const int okmask = I->okmask;
        const int buf = it & 1;
        int &s_bad = s_badv[it & 1];
        const int32_t *sidx = nullptr;
        const double *sval = nullptr;
        int64_t nstaged = 0;
        if (TMA) {
            if (tid == 0 && pos + gridDim.x < A.a.ncell) issue(pos + gridDim.x, buf ^ 1); /* the other buffer was released by the barrier that ended the previous cell */
            mbar_wait((uint32_t)__cvta_generic_to_shared(&s_bar[buf]), (uint32_t)((it >> 1) & 1));
            nstaged = min((int64_t)RPV_STAGE_NNZ, q1 - q0);
            sidx = reinterpret_cast<const int32_t *>(stage + (size_t)buf * ST_BYTES) + (q0 & 3);
            sval = reinterpret_cast<const double *>(stage + (size_t)buf * ST_BYTES + ST_IDX_BYTES) + (q0 & 1);
        }
        auto fetch = [&](int64_t q, int &gi, double &x) -> bool { /* non-zero q of the column (absolute index) */
            if (q >= q1) { gi = 0; x = 0.0; return false; }
            if (TMA && q - q0 < nstaged) { gi = sidx[q - q0]; x = sval[q - q0]; }
            else { gi = A.a.rowidx[q]; x = A.a.val[q]; }
            return true;
        };
        // ---- scatter: every warp takes chunks of 32 non-zeros; generic ones wait in the warp's list ----
        int scnt = 0;
        int64_t q = q0 + warp * 32;
        int gi;
        double x;
        bool valid = fetch(q + lane, gi, x);
        while (q < q1) { /* the next chunk is requested before this chunk's atomics are issued */
            const int64_t qn = q + NT;
            int ngi;
            double nx;
            const bool nvalid = fetch(qn + lane, ngi, nx);
            const int cls = valid ? rpv_class(x) : -1;
            const bool gen = valid && cls < 0 && x != 0.0; /* explicit zeros contribute nothing */
            const unsigned gmask = __ballot_sync(0xffffffffu, gen);
            if (gen) glist[scnt + __popc(gmask & ((1u << lane) - 1u))] = (uint32_t)(q + lane - q0);
            scnt += __popc(gmask);
            if (cls >= 0 && !((okmask >> cls) & 1)) s_bad = 1;
            rpv_class_chunk<RV, UNC>(A, s_base, arr, gi, cls, cls >= 0, lane);
            __syncwarp();
            if (scnt >= 32) {
                int g2;
                double x2;
                fetch(q0 + glist[lane], g2, x2);
                double v = rp_transform(x2, cs, A.a.normalize, A.a.norm_mul, A.a.logkind);
                if (!isfinite(v)) { s_bad = 1; v = 0.0; } /* the whole row becomes NaN, like the reference's NaN propagation */
                rpv_generic_chunk<RV>(A, s_base, arr, g2, __double2ll_rn(v * qscale), true, lane);
                __syncwarp();
                const uint32_t t = (lane < scnt - 32) ? glist[32 + lane] : 0u;
                __syncwarp();
                if (lane < scnt - 32) glist[lane] = t;
                scnt -= 32;
                __syncwarp();
            }
            q = qn; valid = nvalid; gi = ngi; x = nx;
        }
        if (scnt > 0) {
            const bool vv = lane < scnt;
            int g2 = 0;
            double x2 = 0.0;
            if (vv) fetch(q0 + glist[lane], g2, x2);
            double v = vv ? rp_transform(x2, cs, A.a.normalize, A.a.norm_mul, A.a.logkind) : 0.0;
            if (!isfinite(v)) { s_bad = 1; v = 0.0; }
            rpv_generic_chunk<RV>(A, s_base, arr, g2, __double2ll_rn(v * qscale), vv, lane);
        }
        __syncthreads();
        // ---- output: the exact integer sum -> double (one rounding), times the common factor; reset the words ----
        const bool bad = s_bad != 0;
        const double unscale = I->unscale;
        long long qc[4];
#pragma unroll
        for (int c = 0; c < 4; c++) qc[c] = I->qc[c];
        for (int k = 0, i0 = 0; k < A.a.K; k++, i0 += A.a.p)
            for (int j = tid; j < A.a.p; j += NT) { /* (member, column) without an integer division per output */
                const int i = i0 + j;
                long long tot = (long long)(((unsigned long long)hi[i] << 32) | (unsigned long long)lo[i]);
                lo[i] = 0u;
                hi[i] = 0u;
                const uint32_t wp = cw[i], wn = cw[kpd + i];   /* + counts, - counts: four 8-bit class fields each */
                cw[i] = 0u;
                cw[kpd + i] = 0u;
#pragma unroll
                for (int c = 0; c < 4; c++)
                    tot += qc[c] * (long long)((int)((wp >> (8 * c)) & 255u) - (int)((wn >> (8 * c)) & 255u));
                double r = __dmul_rn(__dmul_rn((double)tot, unscale), A.a.scale);
                if (A.a.round_digits >= 0) r = rp_round(r, A.a.round_digits);
                if (bad) r = __longlong_as_double(0x7ff8000000000000LL);
                A.a.out[((size_t)k * A.a.ncell + pos) * A.a.p + j] = r;
            }
        if (tid == 0) s_badv[(it + 1) & 1] = 0; /* last read in the previous cell's epilogue, first written after the barrier below */
        __syncthreads();
    }
}

template <int RV, int NT, int MINB, bool TMA, bool UNC>
static int launch_v3(sharp_ctx *c, const RpV3Args &A, int64_t ncell, size_t smem) {
    SHARP_SMEM_OPTIN_ONCE((rp_project_v3_kernel<RV, NT, MINB, TMA, UNC>), c->device);
    static thread_local size_t q_smem = 0;
    static thread_local int q_per_sm = 0;
    if (q_per_sm == 0 || q_smem != smem) {
        SHARP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q_per_sm, rp_project_v3_kernel<RV, NT, MINB, TMA, UNC>, NT, smem));
        q_smem = smem;
    }
    const int grid = (int)std::min<int64_t>(ncell, (int64_t)c->sm_count * std::max(1, q_per_sm));
    rp_project_v3_kernel<RV, NT, MINB, TMA, UNC><<<grid, NT, smem, c->stream>>>(A);
    return 0;
}

template <int RV>
static int launch_v3_rv(sharp_ctx *c, const RpV3Args &A, int64_t ncell, bool staged) {
    static const bool unc = getenv("SHARP_RP_COND") == nullptr; /* development switch: set to get the predicated class scatter */
    const size_t acc = (size_t)4 * A.kpd * 4;
    if (staged) {
        constexpr int NT = 512;
        const size_t smem = acc + (size_t)(NT / 32) * RPV_LIST * 4 + 2 * ((size_t)(RPV_STAGE_NNZ + 8) * 4 + (size_t)(RPV_STAGE_NNZ + 4) * 8);
        if (smem <= (size_t)SHARP_SMEM_OPTIN) return launch_v3<RV, NT, 2, true, false>(c, A, ncell, smem);
    }
    constexpr int NT = 256;
    const size_t smem = acc + (size_t)(NT / 32) * RPV_LIST * 4;
    if (unc) return launch_v3<RV, NT, 4, false, true>(c, A, ncell, smem);
    return launch_v3<RV, NT, 4, false, false>(c, A, ncell, smem);
}

// CSC input only.  Returns 1 when this variant does not apply (the caller falls back to rp_project_fx_kernel).
int launch_rp_project_v3(sharp_ctx *c, const RpArgs &Ain, const sharp_rm_dev &rm, bool staged, double *colsum_out, void *info_ws) {
    if (!rm.rec || Ain.dense || !rm.ent16 || rm.max_col_nnz > 255) return 1;
    RpV3Args A;
    A.a = Ain;
    A.a.normalize = Ain.normalize ? 1 : 0; /* the scatter kernel divides by the record's cs */
    A.rec = rm.rec;
    A.info = reinterpret_cast<const RpCellInfo *>(info_ws);
    A.kpd = rm.kpd;
    A.dummy = (uint32_t)(rm.kpd - 32) * 4u;
    A.cb = 1;
    while ((1 << A.cb) <= rm.max_col_nnz) A.cb++;
    if ((size_t)4 * A.kpd * 4 + 8 * RPV_LIST * 4 > (size_t)200 * 1024) return 1;
    prof_begin(c, KID_COLSUM);
    cellprep_kernel<<<(unsigned)((Ain.n + 7) / 8), 256, 0, c->stream>>>(Ain.n, Ain.colptr, Ain.val, Ain.colsum, Ain.normalize, Ain.norm_mul,
                                                                      Ain.logkind, A.cb, colsum_out, reinterpret_cast<RpCellInfo *>(info_ws));
    prof_end(c);
    prof_begin(c, KID_RP_PROJECT);
    int rc;
    switch (rm.rv) {
    case 2: rc = launch_v3_rv<2>(c, A, Ain.ncell, staged); break;
    case 4: rc = launch_v3_rv<4>(c, A, Ain.ncell, staged); break;
    case 8: rc = launch_v3_rv<8>(c, A, Ain.ncell, staged); break;
    case 16: rc = launch_v3_rv<16>(c, A, Ain.ncell, staged); break;
    default: rc = set_error(SHARP_E_ARG, "rp_project: unsupported record size %d", rm.rv);
    }
    prof_end(c);
    SHARP_TRY(rc);
    SHARP_CUDA(cudaGetLastError());
    return 0;
}

size_t rp_cellinfo_bytes(int64_t n) { return (size_t)std::max<int64_t>(n, 1) * sizeof(RpCellInfo); }

}  // namespace sharp
