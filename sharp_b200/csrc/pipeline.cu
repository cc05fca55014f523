// pipeline.cu -- host side of libsharpb200: contexts, workspaces, the C ABI of include/sharp_b200.h and the
// fused SHARP_small / SHARP_large device pipeline.  No CPU fallback anywhere: every entry point enqueues CUDA
// kernels of this library on the context's stream and fails with SHARP_E_CUDA when there is no device.
#include <algorithm>
#include <atomic>
#include <mutex>
#include <chrono>
#include <cstdlib>
#include <cmath>
#include <cstdarg>
#include <cstring>
#include <numeric>
#include <thread>

#include <nvtx3/nvToolsExt.h>

#include "internal.cuh"
#include "metac.cuh"

namespace sharp {

// ---- errors ------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";
static std::atomic<int> g_live_ctx{0};

int set_error(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

const char *status_message(int st) {
    switch (st) {
    case SWEEP_E_HCLUST_NA: return "hclust: NA/NaN/Inf in foreign function call (dissimilarities are not finite; a cell with a constant projection?)";
    case SWEEP_E_KRANGE: return "cutree: elements of 'k' must be between 1 and n";
    case SWEEP_E_NAN: return "get_opt_hclust: the median silhouette is NaN";
    case SWEEP_E_CHNAN: return "get_opt_hclust: which.max(CHind) is empty (all CH indices are NaN)";
    case SWEEP_E_UNSORTED: return "cutree: the 'height' component of 'tree' is not sorted (increasingly)";
    case WM_E_ONECLUSTER: return "wMetaC: missing value where TRUE/FALSE needed (one-cluster fallback, R/wMetaC.R:152)";
    case SWEEP_E_FEWPOINTS: return "get_opt_hclust: fewer than minN.cluster+1 objects (silhouette() returns NA; R: incorrect number of dimensions)";
    case SWEEP_E_OIND: return "get_opt_hclust: subscript out of bounds (v[, oind], R/get_opt_hclust.R:228)";
    case WM_E_FEWCLUSTERS: return "wMetaC: combn(allC, 2) needs at least two clusters";
    case SM_E_FEWCLUSTERS: return "sMetaC: combn(nC, 2) needs at least two clusters";
    case WM_E_TOOMANY: return "more clusters than this implementation's capacity";
    default: return "unknown status";
    }
}

static int rstop(int st, const char *where) {
    int code = (st == WM_E_TOOMANY) ? SHARP_E_LIMIT : SHARP_E_RSTOP;
    return set_error(code, "%s: %s", where, status_message(st));
}

// ---- buffers -------------------------------------------------------------------------------------
int DevBuf::reserve(size_t bytes) {
    if (bytes <= cap) return 0;
    if (ptr) {
        cudaError_t e = cudaFree(ptr);
        ptr = nullptr;
        cap = 0;
        if (e != cudaSuccess) return set_error(SHARP_E_CUDA, "cudaFree: %s", cudaGetErrorString(e));
    }
    /* 1/8 headroom: the parts of one job differ a little in size, and a regrow is a cudaFree, i.e. a device-wide
       synchronisation that stalls the contexts working on other streams */
    size_t want = (bytes + bytes / 8 + (1 << 20) - 1) & ~(size_t)((1 << 20) - 1);
    cudaError_t e = cudaMalloc(&ptr, want);
    if (e != cudaSuccess) {
        ptr = nullptr;
        return set_error(SHARP_E_NOMEM, "cudaMalloc(%zu bytes) failed: %s", want, cudaGetErrorString(e));
    }
    cap = want;
    return 0;
}

void DevBuf::release() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
}

}  // namespace sharp

int sharp_ctx::reserve_pinned(size_t bytes) {
    bytes = (std::max<size_t>(bytes, 8) + 255) & ~(size_t)255;
    if (!arena || 2 * bytes > arena_cap) {
        if (arena) {
            cudaStreamSynchronize(stream);
            cudaFreeHost(arena);
        }
        arena = nullptr;
        arena_cap = arena_off = 0;
        pinned = nullptr;
        half_used[0] = half_used[1] = false;
        half_cur = 0;
        size_t want = std::max<size_t>((size_t)32 << 20, 4 * bytes);
        cudaError_t e = cudaMallocHost((void **)&arena, want);
        if (e != cudaSuccess) return sharp::set_error(SHARP_E_NOMEM, "cudaMallocHost(%zu) failed: %s", want, cudaGetErrorString(e));
        arena_cap = want;
        for (int h = 0; h < 2; h++)
            if (!ev_half[h]) cudaEventCreateWithFlags(&ev_half[h], cudaEventDisableTiming);
    }
    /* The arena is used as two halves.  Leaving a half records an event behind every copy staged in it; coming back to
       it waits for THAT event only -- recorded half an arena ago, normally long complete -- instead of draining the
       stream (a drain in the middle of a group run stalls the host, and with it every other stream's queue). */
    const size_t halfcap = arena_cap / 2;
    const int cur = half_cur;
    if (arena_off + bytes > (size_t)(cur + 1) * halfcap) {
        const int nxt = cur ^ 1;
        cudaEventRecord(ev_half[cur], stream);
        half_used[cur] = true;
        if (half_used[nxt]) {
            cudaError_t e = cudaEventSynchronize(ev_half[nxt]);
            if (e != cudaSuccess) return sharp::set_error(SHARP_E_CUDA, "cudaEventSynchronize: %s", cudaGetErrorString(e));
        }
        arena_off = (size_t)nxt * halfcap;
        half_cur = nxt;
    }
    pinned = arena + arena_off;
    pinned_cap = bytes;
    arena_off += bytes;
    return 0;
}

namespace sharp {

// ---- per-kernel device-time profile -------------------------------------------------------------------
static cudaEvent_t prof_event(sharp_ctx *c) {
    if (!c->prof_pool.empty()) {
        cudaEvent_t e = c->prof_pool.back();
        c->prof_pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}
void prof_begin(sharp_ctx *c, int kid) {
    if (!c->prof_on) return;
    ProfPending p;
    p.kid = kid;
    p.a = prof_event(c);
    p.b = prof_event(c);
    cudaEventRecord(p.a, c->stream);
    c->prof_pending.push_back(p);
    c->prof_open = (int)c->prof_pending.size() - 1;
}
void prof_end(sharp_ctx *c) {
    c->launches++;
    if (!c->prof_on || c->prof_open < 0) return;
    cudaEventRecord(c->prof_pending[c->prof_open].b, c->stream);
    c->prof_open = -1;
    if (c->prof_pending.size() >= 65536) prof_collect(c); /* collecting waits for the events: not inside a run */
}
void prof_collect(sharp_ctx *c) {
    for (auto &p : c->prof_pending) {
        float ms = 0.f;
        if (cudaEventSynchronize(p.b) == cudaSuccess && cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
            c->prof_ms[p.kid] += ms;
            c->prof_n[p.kid]++;
        }
        c->prof_pool.push_back(p.a);
        c->prof_pool.push_back(p.b);
    }
    c->prof_pending.clear();
    cudaGetLastError();
}

// NVTX ranges around the host-side stages of a run (front / blocks / back / complete, per part and group): they cost
// nothing without a profiler attached and give nsys / ncu timelines the structure of the pipeline
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

// workspace slots
enum Slot {
    WS_SRC = 0, WS_COLSUM, WS_PROJ, WS_U, WS_D, WS_DW, WS_HC_INT, WS_HC_DBL, WS_DESC, WS_SWEEP_SCRATCH, WS_ENRP, WS_E1,
    WS_WM_INT, WS_WM_DBL, WS_WM_S, WS_WM_DESC, WS_WM_SCRATCH, WS_SM_INT, WS_SM_DBL, WS_SM_S, WS_SM_SCRATCH, WS_VIEU,
    WS_LABELS, WS_X0, WS_TMP0, WS_TMP1, WS_TMP2, WS_TMP3, WS_START, WS_CEN, WS_EX_A, WS_EX_B, WS_EX_C, WS_GDESC, WS_HC_E, WS_COUNT
};
static_assert(WS_COUNT <= WS_COMM_SLOT, "the last two workspace slots belong to comm.cu and the projection kernel");

// bump allocator over a byte region
struct Bump {
    unsigned char *base;
    size_t off = 0, cap;
    Bump(void *b, size_t c) : base(reinterpret_cast<unsigned char *>(b)), cap(c) {}
    template <class T> T *take(size_t count) {
        off = (off + 255) & ~(size_t)255;
        T *r = reinterpret_cast<T *>(base + off);
        off += count * sizeof(T);
        return r;
    }
};
static size_t bump_size(std::initializer_list<size_t> bytes) {
    size_t t = 256;
    for (size_t b : bytes) t += b + 256;
    return t;
}

// trace aid (SHARP_B200_TRACE): report any host-side call of the enqueue path that blocks for more than 5 ms
struct SlowCall {
    const char *what;
    std::chrono::steady_clock::time_point t0;
    explicit SlowCall(const char *w) : what(w), t0(std::chrono::steady_clock::now()) {}
    ~SlowCall() {
        static const bool trace = getenv("SHARP_B200_TRACE") != nullptr;
        if (!trace) return;
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (ms > 5.0) fprintf(stderr, "[sharp trace slow] %s blocked the host for %.2f ms\n", what, ms);
    }
};
#define SLOW(name, stmt) do { SlowCall _sc(name); stmt; } while (0)

static int ld_of(int n) { return (n + 3) & ~3; }
static int ldu_of(int p) { return (p + 15) & ~15; }

// ---- small utility kernels ---------------------------------------------------------------------------
__global__ void sym_to_dist_kernel(int n, const double *__restrict__ S, int lds, double *D, double *Dw, int ld) {
    /* d = as.dist(1 - mat): as.dist keeps the lower triangle (row > col) of the R matrix */
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)n * n) return;
    const int i = (int)(idx / n), j = (int)(idx % n);
    const int hi = i > j ? i : j, lo = i > j ? j : i;
    const double d = (i == j) ? 0.0 : __dsub_rn(1.0, S[(size_t)hi * lds + lo]);
    D[(size_t)i * ld + j] = d;
    if (Dw) Dw[(size_t)i * ld + j] = d;
}

__global__ void pad_copy_kernel(int n, int ncol, const double *__restrict__ src, int lds, double *dst, int ldd) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)n * ncol) return;
    const int i = (int)(idx / ncol), j = (int)(idx % ncol);
    dst[(size_t)i * ldd + j] = src[(size_t)i * lds + j];
}

// enE/K in member order: E1[i][j] = (((0 + P_1) + P_2) + ...)/K   (R/SHARP.R:629-635, 750)
__global__ void ene_kernel(const double *__restrict__ proj, int64_t np, int K, double *__restrict__ E1) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= np) return;
    double s = 0.0;
    for (int k = 0; k < K; k++) s = __dadd_rn(s, proj[(size_t)k * np + idx]);
    E1[idx] = __ddiv_rn(s, (double)K);
}

// colour index: cluster ids above 40 wrap (R/getrowColor.R:59-68)
__global__ void colour_wrap_kernel(int32_t *lab, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int c = lab[i];
    if (c > 40) { c = c % 40; if (c == 0) c = 40; lab[i] = c; }
}

// scatter rows: dst[row_of[i]] = src[i]  (viE[reind, ] = viE)
__global__ void scatter_rows_kernel(const double *__restrict__ src, int64_t n, int p, const int64_t *__restrict__ row_of,
                                    double *__restrict__ dst) {
    const int64_t i = blockIdx.x;
    if (i >= n) return;
    const int64_t r = row_of ? row_of[i] : i;
    for (int j = threadIdx.x; j < p; j += blockDim.x) dst[(size_t)r * p + j] = src[(size_t)i * p + j];
}

static int grid1d(size_t n, int threads) { return (int)((n + threads - 1) / threads); }

// ---- upload helpers -------------------------------------------------------------------------------
// staged = true: the device copy lives in the context's grow-only workspace (no cudaMalloc / cudaFree per call --
// cudaFree synchronises the whole device, which would serialise contexts working on other streams) and is valid
// until the next staged upload on this context; staged = false: the caller owns it (sharp_expr_upload).
static int upload_expr(sharp_ctx *c, int m, int64_t n, const double *dense, const int64_t *colptr, const int32_t *rowidx,
                       const double *val, sharp_expr_dev *e, bool staged = false, cudaStream_t on = nullptr) {
    const cudaStream_t st = on ? on : c->stream;
    e->device = c->device;
    e->m = m;
    e->n = n;
    e->owned = !staged;
    if (m <= 0 || n < 0) return set_error(SHARP_E_ARG, "expression matrix: bad dimensions %d x %lld", m, (long long)n);
    auto get = [&](int slot, void **ptr, size_t bytes) -> int {
        bytes = std::max<size_t>(bytes, 8) + 64; /* slack: the staged projection kernel widens its bulk copies to 16-byte boundaries */
        if (staged) {
            SHARP_TRY(c->ws[slot].reserve(bytes));
            *ptr = c->ws[slot].ptr;
        } else SHARP_CUDA(cudaMalloc(ptr, bytes));
        return 0;
    };
    if (dense) {
        size_t bytes = (size_t)m * n * sizeof(double);
        SHARP_TRY(get(WS_EX_A, (void **)&e->dense, bytes));
        SHARP_CUDA(cudaMemcpyAsync(e->dense, dense, bytes, cudaMemcpyHostToDevice, st));
    } else {
        if (!colptr || (!rowidx && colptr[n] > 0) || (!val && colptr[n] > 0))
            return set_error(SHARP_E_ARG, "expression matrix: neither dense nor complete CSC slots given");
        e->nnz = colptr[n];
        SHARP_TRY(get(WS_EX_A, (void **)&e->colptr, (size_t)(n + 1) * 8));
        SHARP_TRY(get(WS_EX_B, (void **)&e->rowidx, (size_t)e->nnz * 4));
        SHARP_TRY(get(WS_EX_C, (void **)&e->val, (size_t)e->nnz * 8));
        SHARP_CUDA(cudaMemcpyAsync(e->colptr, colptr, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, st));
        SHARP_CUDA(cudaMemcpyAsync(e->rowidx, rowidx, (size_t)e->nnz * 4, cudaMemcpyHostToDevice, st));
        SHARP_CUDA(cudaMemcpyAsync(e->val, val, (size_t)e->nnz * 8, cudaMemcpyHostToDevice, st));
    }
    return 0;
}

static void free_expr(sharp_expr_dev *e) {
    if (!e) return;
    if (!e->owned) {
        e->dense = nullptr;
        e->colptr = nullptr;
        e->rowidx = nullptr;
        e->val = nullptr;
        return;
    }
    if (e->dense) cudaFree(e->dense);
    if (e->colptr) cudaFree(e->colptr);
    if (e->rowidx) cudaFree(e->rowidx);
    if (e->val) cudaFree(e->val);
    e->dense = nullptr;
    e->colptr = nullptr;
    e->rowidx = nullptr;
    e->val = nullptr;
}

static int h2d(sharp_ctx *c, void *dst, const void *src, size_t bytes) {
    if (bytes == 0) return 0;
    SHARP_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
    return 0;
}
// Host -> device copy of a chunk of the context's PINNED arena (descriptor tables, index vectors: bytes to a few hundred
// KB) done by a kernel that reads the mapped host memory, not by the copy engine: the copy engine serves the streams'
// copies in order, so behind the multi-GB expression uploads of a group run a 4 KB descriptor copy would wait for
// every part queued before it -- and the kernels behind that copy with it.
__global__ void stage_copy_kernel(uint32_t *__restrict__ dst, const uint32_t *__restrict__ src, size_t nwords) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
// the same copy on an explicit stream, any size (multiple of 4 bytes, 4-byte aligned): `src` is page-locked host memory
int h2d_by_kernel(cudaStream_t st, void *dst, const void *src, size_t bytes) {
    if (bytes == 0) return 0;
    if ((bytes & 3) || ((uintptr_t)dst & 3) || ((uintptr_t)src & 3)) return set_error(SHARP_E_ARG, "h2d_by_kernel: unaligned");
    const size_t nw = bytes / 4;
    const int grid = (int)std::min<size_t>(148, (nw + 255) / 256);
    stage_copy_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<uint32_t *>(dst), reinterpret_cast<const uint32_t *>(src), nw);
    SHARP_CUDA(cudaGetLastError());
    return 0;
}
static int h2d_staged(sharp_ctx *c, void *dst, const void *src, size_t bytes) {
    if (bytes == 0) return 0;
    if ((bytes & 3) || ((uintptr_t)dst & 3) || ((uintptr_t)src & 3) || bytes > ((size_t)16 << 20)) return h2d(c, dst, src, bytes);
    const size_t nw = bytes / 4;
    const int grid = (int)std::min<size_t>(64, (nw + 255) / 256);
    stage_copy_kernel<<<grid, 256, 0, c->stream>>>(reinterpret_cast<uint32_t *>(dst), reinterpret_cast<const uint32_t *>(src), nw);
    SHARP_CUDA(cudaGetLastError());
    return 0;
}
static int d2h(sharp_ctx *c, void *dst, const void *src, size_t bytes) {
    if (bytes == 0) return 0;
    SHARP_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
    return 0;
}
static int sync(sharp_ctx *c) {
    SHARP_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}
static int use(sharp_ctx *c) {
    if (!c) return set_error(SHARP_E_ARG, "null context");
    SHARP_CUDA(cudaSetDevice(c->device));
    return 0;
}

// =====================================================================================================
// get_opt_hclust on one matrix already on the device
// =====================================================================================================
struct OptResult {  // device pointers into workspaces, valid until the next call on the context
    int *f;
    int *v;
    double *msil, *chind, *crit;
    int *meta;
    double *maxsil;
    int maxlev;
};

// mat_dev: row-major nrow x ncol on the device.  exact: 1 exact sweep, 0 nested (feature only).
static int opt_hclust_dev(sharp_ctx *c, int nrow, int ncol, const double *mat_dev, int symmetric, int exact,
                          const sharp_hc_params &prm_in, bool want_v, OptResult *R) {
    HcParamsDev prm = to_dev(prm_in);
    const int n = nrow;
    if (n < 2) return set_error(SHARP_E_RSTOP, "hclust: must have n >= 2 objects to cluster");
    if (symmetric && nrow != ncol) return set_error(SHARP_E_ARG, "a similarity matrix must be square");
    if (prm.n_cluster != 0 && prm.n_cluster < 2)
        return set_error(SHARP_E_RSTOP, "The given N.cluster is less than 2, which is not suitable for clustering!");
    const int ld = ld_of(n);
    int maxlev = prm.n_cluster ? 1 : std::max(1, prm.max_n - prm.min_n + 1);
    int kcap = prm.n_cluster ? prm.n_cluster : prm.max_n;
    if (!symmetric && !exact && std::min(kcap, n - 1) > NESTED_MAXK_HOST) exact = 1;
    if (symmetric) exact = 1;
    SHARP_TRY(c->ws[WS_D].reserve((size_t)n * ld * 8));
    SHARP_TRY(c->ws[WS_DW].reserve((size_t)n * ld * 8));
    double *D = c->ws[WS_D].as<double>(), *Dw = c->ws[WS_DW].as<double>();
    const double *Y;
    int yp, ldy;
    if (symmetric) {
        prof_begin(c, KID_MISC);
        sym_to_dist_kernel<<<grid1d((size_t)n * n, 256), 256, 0, c->stream>>>(n, mat_dev, ncol, D, Dw, ld);
        prof_end(c);
        Y = mat_dev;
        yp = n;
        ldy = ncol;
    } else {
        const int ldu = ldu_of(ncol);
        SHARP_TRY(c->ws[WS_U].reserve((size_t)n * ldu * 8));
        double *U = c->ws[WS_U].as<double>();
        SHARP_TRY(launch_unit_rows(c, mat_dev, n, ncol, ldu, U));
        SHARP_TRY(c->reserve_pinned(4096));
        SHARP_TRY(c->ws[WS_DESC].reserve(4096));
        Bump hb(c->pinned, 4096), db(c->ws[WS_DESC].ptr, 4096);
        GemmProb *gp = hb.take<GemmProb>(1);
        int *tp = hb.take<int>(2);
        gp->U = U; gp->n = n; gp->ld = ld; gp->D = D; gp->Dw = Dw;
        tp[0] = 0; tp[1] = corrdist_tiles(n);
        GemmProb *gpd = db.take<GemmProb>(1);
        int *tpd = db.take<int>(2);
        SHARP_TRY(h2d_staged(c, c->ws[WS_DESC].ptr, c->pinned, hb.off));
        SHARP_TRY(launch_corrdist_batched(c, gpd, tpd, 1, tp[1], ldu));
        Y = U;
        yp = ncol;
        ldy = ldu;
    }
    // hc problem + outputs
    size_t ibytes = bump_size({(size_t)n * 4, (size_t)n * 4, (size_t)n * 4, (size_t)n * maxlev * 4, 64});
    size_t dbytes = bump_size({(size_t)n * 8, (size_t)maxlev * 8, (size_t)maxlev * 8, 64});
    SHARP_TRY(c->ws[WS_HC_INT].reserve(ibytes));
    SHARP_TRY(c->ws[WS_HC_DBL].reserve(dbytes));
    Bump bi(c->ws[WS_HC_INT].ptr, ibytes), bd(c->ws[WS_HC_DBL].ptr, dbytes);
    int *ia = bi.take<int>(n), *ib = bi.take<int>(n);
    R->f = bi.take<int>(n);
    R->v = want_v ? bi.take<int>((size_t)n * maxlev) : nullptr;
    R->meta = bi.take<int>(8);
    R->crit = bd.take<double>(n);
    R->msil = bd.take<double>(maxlev);
    R->chind = bd.take<double>(maxlev);
    R->maxsil = bd.take<double>(1);
    R->maxlev = maxlev;
    SHARP_CUDA(cudaMemsetAsync(R->meta, 0, 8 * 4, c->stream));
    SHARP_CUDA(cudaMemsetAsync(R->msil, 0, (size_t)maxlev * 8, c->stream));
    SHARP_CUDA(cudaMemsetAsync(R->chind, 0, (size_t)maxlev * 8, c->stream));
    SHARP_TRY(c->reserve_pinned(4096));
    SHARP_TRY(c->ws[WS_DESC].reserve(4096));
    Bump hb(c->pinned, 4096), db(c->ws[WS_DESC].ptr, 4096);
    HcProb *hp = hb.take<HcProb>(1);
    SweepOut *so = hb.take<SweepOut>(1);
    HcParamsDev *pp = hb.take<HcParamsDev>(1);
    hp->n = n; hp->ld = ld; hp->D = D; hp->Dw = Dw; hp->ia = ia; hp->ib = ib; hp->crit = R->crit;
    hp->Y = Y; hp->p = yp; hp->ldy = ldy; hp->status = 0; hp->fallback = 0;
    hp->E = nullptr; hp->ecap = 0;
    if (!symmetric && hclust_fast_ok(n, prm.hmethod)) {
        SHARP_TRY(c->ws[WS_HC_E].reserve((size_t)n * ld * 8));
        hp->E = c->ws[WS_HC_E].as<double>();
        hp->ecap = (long long)n * ld;
    }
    so->f = R->f; so->v = R->v; so->msil = R->msil; so->chind = R->chind; so->meta = R->meta; so->maxsil = R->maxsil;
    *pp = prm;
    HcProb *hpd = db.take<HcProb>(1);
    SweepOut *sod = db.take<SweepOut>(1);
    HcParamsDev *ppd = db.take<HcParamsDev>(1);
    SHARP_TRY(h2d_staged(c, c->ws[WS_DESC].ptr, c->pinned, hb.off));
    SHARP_TRY(launch_hclust(c, hpd, 1, n, prm.hmethod, symmetric ? 0 : 1));
    if (exact) {
        size_t sb = sweep_exact_scratch_bytes(n, yp);
        SHARP_TRY(c->ws[WS_SWEEP_SCRATCH].reserve(sb));
        SHARP_TRY(launch_sweep_exact(c, hpd, sod, 1, n, yp, ppd, maxlev, std::max(kcap, 2), c->ws[WS_SWEEP_SCRATCH].as<double>(), sb));
    } else {
        size_t sb = sweep_nested_scratch_bytes(n, yp, prm);
        SHARP_TRY(c->ws[WS_SWEEP_SCRATCH].reserve(sb));
        SHARP_TRY(launch_sweep_nested(c, hpd, sod, 1, n, yp, prm, c->ws[WS_SWEEP_SCRATCH].as<double>(), sb));
    }
    return 0;
}

// =====================================================================================================
// wMetaC on labels already on the device: [K][ncells] codes, T blocks given by start (host copy too)
// =====================================================================================================
struct WmBuffers {
    WmArgs A;
    SweepOut *outs_dev;
    HcParamsDev *prm_dev;
    int T, max_block_n;
};

static int wmetac_dev(sharp_ctx *c, const int32_t *labels_dev, int64_t ncells, int K, int T, const int64_t *start_host,
                      const int64_t *start_dev, int max_label, const sharp_hc_params &prm_in, WmBuffers *W) {
    HcParamsDev prm = to_dev(prm_in);
    if (prm.n_cluster != 0 && prm.n_cluster < 2)
        return set_error(SHARP_E_RSTOP, "The given N.cluster is less than 2, which is not suitable for clustering!");
    if (K > WM_MAXK) return set_error(SHARP_E_LIMIT, "wMetaC: at most %d clustering solutions (got %d)", WM_MAXK, K);
    if (max_label > WM_MAXL) return set_error(SHARP_E_LIMIT, "wMetaC: at most %d clusters per solution (got %d)", WM_MAXL, max_label);
    int max_block_n = 0;
    for (int t = 0; t < T; t++) max_block_n = std::max<int>(max_block_n, (int)(start_host[t + 1] - start_host[t]));
    const int capC = std::max(2, K * std::min(max_label, max_block_n));
    const int kcap = prm.n_cluster ? prm.n_cluster : prm.max_n;
    const int capU = std::max(2, std::min(std::max(kcap, 2), capC));
    if (capU > WM_MAXU) return set_error(SHARP_E_LIMIT, "wMetaC: at most %d meta-clusters (got %d)", WM_MAXU, capU);
    const int maxlev = prm.n_cluster ? 1 : std::max(1, prm.max_n - prm.min_n + 1);
    WmArgs &A = W->A;
    A.labels = labels_dev;
    A.ncells = ncells;
    A.K = K;
    A.start = start_dev;
    A.capC = capC;
    A.capU = capU;
    size_t ibytes = bump_size({(size_t)K * ncells * 4, (size_t)K * ncells * 4, (size_t)T * (capC + 1) * 4, (size_t)T * capC * 4,
                               (size_t)T * 4, (size_t)T * capC * 4, (size_t)T * capC * 4, (size_t)ncells * 4,
                               (size_t)ncells * 4, (size_t)T * 4, (size_t)T * capU * 4, (size_t)T * 4,
                               (size_t)T * capC * 4, (size_t)T * 8 * 4});
    size_t dbytes = bump_size({(size_t)ncells * 8, (size_t)T * capC * 8, (size_t)T * maxlev * 8, (size_t)T * maxlev * 8, (size_t)T * 8});
    SHARP_TRY(c->ws[WS_WM_INT].reserve(ibytes));
    SHARP_TRY(c->ws[WS_WM_DBL].reserve(dbytes));
    SHARP_TRY(c->ws[WS_WM_S].reserve((size_t)3 * T * capC * capC * 8));
    Bump bi(c->ws[WS_WM_INT].ptr, ibytes), bd(c->ws[WS_WM_DBL].ptr, dbytes);
    A.gid = bi.take<int>((size_t)K * ncells);
    A.members = bi.take<int>((size_t)K * ncells);
    A.moff = bi.take<int>((size_t)T * (capC + 1));
    A.col_of = bi.take<int>((size_t)T * capC);
    A.allc = bi.take<int>(T);
    A.ia = bi.take<int>((size_t)T * capC);
    A.ib = bi.take<int>((size_t)T * capC);
    A.finalc = bi.take<int>(ncells);
    A.fcode = bi.take<int>(ncells);
    A.ucount = bi.take<int>(T);
    A.ulist = bi.take<int>((size_t)T * capU);
    A.status = bi.take<int>(T);
    int *meta_f = bi.take<int>((size_t)T * capC);
    int *meta_meta = bi.take<int>((size_t)T * 8);
    A.w1 = bd.take<double>(ncells);
    A.crit = bd.take<double>((size_t)T * capC);
    double *msil = bd.take<double>((size_t)T * maxlev);
    double *chind = bd.take<double>((size_t)T * maxlev);
    double *maxsil = bd.take<double>(T);
    A.S = c->ws[WS_WM_S].as<double>();
    A.D = A.S + (size_t)T * capC * capC;
    A.Dw = A.D + (size_t)T * capC * capC;
    SHARP_CUDA(cudaMemsetAsync(meta_meta, 0, (size_t)T * 8 * 4, c->stream));
    SHARP_CUDA(cudaMemsetAsync(A.status, 0, (size_t)T * 4, c->stream));
    SHARP_CUDA(cudaMemsetAsync(A.ucount, 0, (size_t)T * 4, c->stream));
    // descriptors
    size_t desc = bump_size({(size_t)T * sizeof(HcProb), (size_t)T * sizeof(SweepOut), (size_t)T * sizeof(HcParamsDev)});
    SHARP_TRY(c->reserve_pinned(desc));
    SHARP_TRY(c->ws[WS_WM_DESC].reserve(desc));
    Bump hb(c->pinned, desc), db(c->ws[WS_WM_DESC].ptr, desc);
    (void)hb.take<HcProb>(T); /* filled on the device by wm_similarity */
    SweepOut *so = hb.take<SweepOut>(T);
    HcParamsDev *pp = hb.take<HcParamsDev>(T);
    for (int t = 0; t < T; t++) {
        so[t].f = meta_f + (size_t)t * capC;
        so[t].v = nullptr;
        so[t].msil = msil + (size_t)t * maxlev;
        so[t].chind = chind + (size_t)t * maxlev;
        so[t].meta = meta_meta + (size_t)t * 8;
        so[t].maxsil = maxsil + t;
        pp[t] = prm;
    }
    A.probs = db.take<HcProb>(T);
    W->outs_dev = db.take<SweepOut>(T);
    W->prm_dev = db.take<HcParamsDev>(T);
    SHARP_TRY(h2d_staged(c, c->ws[WS_WM_DESC].ptr, c->pinned, hb.off));
    W->T = T;
    W->max_block_n = max_block_n;
    SHARP_TRY(launch_wmetac_front(c, A, T, max_block_n));
    SHARP_TRY(launch_hclust(c, A.probs, T, capC, prm.hmethod));
    size_t sb = sweep_exact_scratch_bytes(capC, capC);
    SHARP_TRY(c->ws[WS_WM_SCRATCH].reserve(sb * T));
    SHARP_TRY(launch_sweep_exact(c, A.probs, W->outs_dev, T, capC, capC, W->prm_dev, maxlev, std::max(kcap, 2),
                                 c->ws[WS_WM_SCRATCH].as<double>(), sb));
    SHARP_TRY(launch_wmetac_vote(c, A, W->outs_dev, T));
    return 0;
}

// =====================================================================================================
// sMetaC core on the device: cluster member lists (corder/coff) and E1 given
// =====================================================================================================
struct SmBuffers {
    SmArgs A;
    SweepOut *out_dev;
    int *tf;        // [capS]
    int *status;    // [1]
    double *cen;    // [capS][p]
};

// nc_ptr/status_in: device ints.  cen_in: optional precomputed centroids (device) -- else computed from E1.
static int smetac_dev(sharp_ctx *c, int capS, int p, const int *nc_ptr, const int *status_in, const double *E1,
                      const int *corder, const int *coff, const double *cen_in, int64_t ncells_total,
                      const sharp_hc_params &prm_in, SmBuffers *B) {
    HcParamsDev prm = to_dev(prm_in);
    if (prm.n_cluster != 0 && prm.n_cluster < 2)
        return set_error(SHARP_E_RSTOP, "The given N.cluster is less than 2, which is not suitable for clustering!");
    // host-side upper bounds of the tweaked range (R/sMetaC.R:101-119)
    int maxN = prm.max_n, minN = prm.min_n;
    if (ncells_total >= 1000000) {
        maxN = std::max(maxN, (int)(ncells_total / 5000));
        minN = std::max(minN, (int)(ncells_total / 50000));
    }
    const int maxlev = prm.n_cluster ? 1 : std::max(1, maxN - std::min(minN, prm.min_n) + 1);
    const int kcap = std::max(2, prm.n_cluster ? prm.n_cluster : maxN);
    const int ld = ld_of(capS);
    size_t ibytes = bump_size({(size_t)ld * 4, (size_t)ld * 4, (size_t)ld * 4, (size_t)ld * 4, 64, 64});
    size_t dbytes = bump_size({(size_t)ld * 8, (size_t)maxlev * 8, (size_t)maxlev * 8, 64, (size_t)capS * 8, (size_t)capS * 8,
                               (size_t)capS * p * 8});
    SHARP_TRY(c->ws[WS_SM_INT].reserve(ibytes));
    SHARP_TRY(c->ws[WS_SM_DBL].reserve(dbytes));
    SHARP_TRY(c->ws[WS_SM_S].reserve((size_t)3 * ld * ld * 8));
    Bump bi(c->ws[WS_SM_INT].ptr, ibytes), bd(c->ws[WS_SM_DBL].ptr, dbytes);
    SmArgs &A = B->A;
    A.ia = bi.take<int>(ld);
    A.ib = bi.take<int>(ld);
    int *f = bi.take<int>(ld);
    B->tf = bi.take<int>(ld);
    int *meta = bi.take<int>(8);
    B->status = bi.take<int>(1);
    A.crit = bd.take<double>(ld);
    double *msil = bd.take<double>(maxlev);
    double *chind = bd.take<double>(maxlev);
    double *maxsil = bd.take<double>(1);
    double *mean = bd.take<double>(capS);
    double *sdev = bd.take<double>(capS);
    B->cen = bd.take<double>((size_t)capS * p);
    A.S = c->ws[WS_SM_S].as<double>();
    A.D = A.S + (size_t)ld * ld;
    A.Dw = A.D + (size_t)ld * ld;
    A.ld = ld;
    A.nc_ptr = nc_ptr;
    A.status_in = status_in;
    A.prm = prm;
    A.ncells_total = ncells_total;
    A.tf = B->tf;
    A.status_out = B->status;
    SHARP_CUDA(cudaMemsetAsync(meta, 0, 8 * 4, c->stream));
    size_t desc = bump_size({sizeof(HcProb), sizeof(SweepOut), sizeof(HcParamsDev)});
    SHARP_TRY(c->reserve_pinned(desc));
    SHARP_TRY(c->ws[WS_DESC].reserve(desc));
    Bump hb(c->pinned, desc), db(c->ws[WS_DESC].ptr, desc);
    (void)hb.take<HcProb>(1);
    SweepOut *so = hb.take<SweepOut>(1);
    (void)hb.take<HcParamsDev>(1);
    so->f = f; so->v = nullptr; so->msil = msil; so->chind = chind; so->meta = meta; so->maxsil = maxsil;
    A.prob = db.take<HcProb>(1);
    B->out_dev = db.take<SweepOut>(1);
    A.prm_out = db.take<HcParamsDev>(1);
    SHARP_TRY(h2d_staged(c, c->ws[WS_DESC].ptr, c->pinned, hb.off));
    const double *cen = cen_in;
    if (!cen) {
        SHARP_TRY(launch_sm_centroids(c, E1, p, corder, coff, nc_ptr, capS, B->cen, nullptr));
        cen = B->cen;
    }
    SHARP_TRY(launch_sm_similarity(c, A, cen, p, capS, mean, sdev));
    SHARP_TRY(launch_hclust(c, A.prob, 1, ld, prm.hmethod));
    size_t sb = sweep_exact_scratch_bytes(ld, ld);
    SHARP_TRY(c->ws[WS_SM_SCRATCH].reserve(sb));
    SHARP_TRY(launch_sweep_exact(c, A.prob, B->out_dev, 1, ld, ld, A.prm_out, maxlev, kcap, c->ws[WS_SM_SCRATCH].as<double>(), sb));
    SHARP_TRY(launch_sm_finish(c, A, B->out_dev));
    return 0;
}

// SHARP_B200_TRACE=1: host wall-clock per phase of run_core on stderr (development aid)
struct Trace {
    bool on;
    std::chrono::steady_clock::time_point t0, last;
    std::string line;
    sharp_ctx *c;
    explicit Trace(sharp_ctx *ctx) : c(ctx) {
        static const bool env = getenv("SHARP_B200_TRACE") != nullptr;
        on = env;
        if (on) t0 = last = std::chrono::steady_clock::now();
    }
    void mark(const char *what, bool dev_sync = false) {
        if (!on) return;
        if (dev_sync) cudaStreamSynchronize(c->stream);
        auto now = std::chrono::steady_clock::now();
        char buf[96];
        snprintf(buf, sizeof buf, " %s=%.2f", what, std::chrono::duration<double, std::milli>(now - last).count());
        line += buf;
        last = now;
    }
    ~Trace() {
        if (on) fprintf(stderr, "[sharp trace dev%d ctx%p]%s total=%.2f ms\n", c->device, (void *)c, line.c_str(),
                        std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    }
};

// R/SHARP.R:513-536: block boundaries in the (shuffled) cell order
static void make_blocks(int64_t n, int large, int ng, std::vector<int64_t> &start) {
    start.clear();
    if (!large || ng <= 0 || n <= ng) {
        start = {0, n};
        return;
    }
    int64_t T = (n + ng - 1) / ng;
    int64_t nt = n - (T - 2) * ng;
    start.push_back(0);
    for (int64_t t = 0; t < T - 2; t++) start.push_back(start.back() + ng);
    start.push_back(start.back() + nt / 2);
    start.push_back(n);
}

// =====================================================================================================
// the fused SHARP_small / SHARP_large pipeline, in phases
// =====================================================================================================
// One matrix (or one part of SHARP_unlimited) goes through
//   part_front   colSums, K1 projection, unit rows, per-problem output tables          (stream of the part's context)
//   run_blocks   K2 distances, K3 agglomeration, K4-K6 sweep of ALL (member, block) problems of one or SEVERAL parts
//                in shared launches                                                     (stream of the group context)
//   part_back    colour wrap, enE/K, wMetaC per block, sMetaC over blocks, relabel, un-shuffle; statuses and labels are
//                copied to pinned mirrors                                               (stream of the part's context)
//   part_finish  the only host synchronisation: statuses -> R-style errors, x0 / viE outputs
// Nothing between part_front and part_finish waits for the device, so the host can enqueue several parts (and the
// next group) while the GPU works: the agglomeration kernel is latency-bound with one CTA per problem, and several
// parts' problems in ONE launch are what fills the SMs (3-4 resident CTAs each instead of 0.85).
struct PartRun {
    sharp_ctx *c = nullptr;
    sharp_expr_dev e;
    const sharp_rm_dev *rm = nullptr;
    sharp_run_params Q;
    int64_t n = 0;
    int p = 0, K = 0, T = 0, max_bn = 0, nprob = 0, ldu = 0, maxlev = 0, kcap_ind = 0;
    bool shuffle = false, nested = true;
    std::vector<int64_t> start;
    int64_t *src_dev = nullptr, *start_dev = nullptr;
    double *proj = nullptr, *U = nullptr, *E1 = nullptr, *vieu = nullptr;
    int32_t *enrp = nullptr;
    int *ia_all = nullptr, *ib_all = nullptr, *meta_all = nullptr;
    double *crit_all = nullptr, *msil_all = nullptr, *chind_all = nullptr, *maxsil_all = nullptr;
    HcParamsDev ind;
    // back
    WmBuffers W;
    SmBuffers B;
    int *labels_dev = nullptr, *coloff_dev = nullptr, *nc_dev = nullptr;
    int *h_meta = nullptr, *h_wst = nullptr, *h_sst = nullptr, *h_nu = nullptr;  // pinned mirrors
    bool skipped_smetac = false;
    // block-sharded run (sharp_run_params.shard / sharp_part.sharded with a communicator on the context): this rank
    // clusters blocks [t0, t0 + T) of the T_all blocks; n / start / enrp / proj ... above are LOCAL (its own cells, in
    // shuffled order), na / start_all are the whole matrix.  The per-block wMetaC results and the rows of enE / K are
    // allgathered, after which every rank runs the cross-block stage (sMetaC, relabel, un-shuffle) on the whole matrix.
    bool sharded = false;
    int64_t na = 0, pos0 = 0;
    int t0 = 0, T_all = 0;
    std::vector<int64_t> start_all;
    int64_t *src_all_dev = nullptr, *start_all_dev = nullptr;
    double *E1_all = nullptr;
};

// blocks [t0, t1) of T for one rank: as even as possible, the larger shares rotated by `rotate` (successive sharded
// matrices give their extra blocks to different ranks)
static void shard_range(int T, int rank, int world, int rotate, int *t0, int *t1) {
    const int r = ((rank + rotate) % world + world) % world;
    const int base = T / world, extra = T % world;
    *t0 = r * base + std::min(r, extra);
    *t1 = *t0 + base + (r < extra ? 1 : 0);
}

static int part_front(PartRun &R, sharp_ctx *c, const sharp_expr_dev &e, const double *colsum_host,
                      const sharp_rm_dev &rm, const int64_t *reind, const sharp_run_params &Q) {
    R.c = c;
    R.e = e;
    R.rm = &rm;
    R.Q = Q;
    const int64_t na = R.na = e.n_total ? e.n_total : e.n;   /* cells of the whole matrix (e may hold a column slice of it) */
    const int p = R.p = rm.p, K = R.K = rm.K;
    if (na < 2) return set_error(SHARP_E_ARG, "need at least 2 cells");
    if (na > 2000000000LL / std::max(1, K)) return set_error(SHARP_E_LIMIT, "too many cells for one call (%lld); split into parts", (long long)na);
    make_blocks(na, Q.large, Q.partition_ncells, R.start_all);
    R.T_all = (int)R.start_all.size() - 1;
    R.shuffle = Q.large && reind && na < 100000;
    R.sharded = Q.large && Q.shard && c->comm && c->comm_world > 1;
    R.t0 = 0;
    int t1 = R.T_all;
    if (R.sharded) {
        if (R.T_all < c->comm_world) return set_error(SHARP_E_ARG, "sharded run: %d blocks cannot be dealt over %d ranks", R.T_all, c->comm_world);
        shard_range(R.T_all, c->comm_rank, c->comm_world, Q.shard_rotate, &R.t0, &t1);
    } else if (e.n_total && e.n_total != e.n) return set_error(SHARP_E_ARG, "a column slice needs a sharded run");
    const int T = R.T = t1 - R.t0;
    R.pos0 = R.start_all[R.t0];
    R.start.resize((size_t)T + 1);
    for (int t = 0; t <= T; t++) R.start[t] = R.start_all[R.t0 + t] - R.pos0;
    const int64_t n = R.n = R.start[T];
    R.max_bn = 0;
    for (int t = 0; t < T; t++) R.max_bn = std::max<int>(R.max_bn, (int)(R.start[t + 1] - R.start[t]));
    if (R.shuffle && e.col0 != 0) return set_error(SHARP_E_ARG, "a shuffled matrix cannot be given as a column slice");
    if (!R.shuffle && R.sharded && (e.col0 > R.pos0 || e.col0 + e.n < R.pos0 + n))
        return set_error(SHARP_E_ARG, "the column slice does not cover this rank's blocks");

    // ---- inputs: source column per position, block boundaries, column sums ----
    R.src_dev = nullptr;
    R.src_all_dev = nullptr;
    if (R.shuffle) {
        SHARP_TRY(c->ws[WS_SRC].reserve((size_t)na * 8));
        R.src_all_dev = c->ws[WS_SRC].as<int64_t>();
        SHARP_TRY(c->reserve_pinned((size_t)na * 8));
        int64_t *hp = reinterpret_cast<int64_t *>(c->pinned);
        for (int64_t i = 0; i < na; i++) {
            if (reind[i] < 1 || reind[i] > na) return set_error(SHARP_E_ARG, "reind is not a permutation of 1..n");
            hp[i] = reind[i] - 1;
        }
        SHARP_TRY(h2d_staged(c, R.src_all_dev, hp, (size_t)na * 8));
        R.src_dev = e.compact ? nullptr : R.src_all_dev + R.pos0;   /* this rank's cells: positions pos0 .. pos0 + n of the shuffled order */
    } else if (R.sharded && R.pos0 != e.col0) { /* un-shuffled, not starting at the matrix's first column: explicit source list */
        SHARP_TRY(c->ws[WS_SRC].reserve((size_t)n * 8));
        R.src_dev = c->ws[WS_SRC].as<int64_t>();
        SHARP_TRY(c->reserve_pinned((size_t)n * 8));
        int64_t *hp = reinterpret_cast<int64_t *>(c->pinned);
        for (int64_t i = 0; i < n; i++) hp[i] = R.pos0 - e.col0 + i;
        SHARP_TRY(h2d_staged(c, R.src_dev, hp, (size_t)n * 8));
    }
    SHARP_TRY(c->ws[WS_START].reserve((size_t)(T + 1) * 8));
    R.start_dev = c->ws[WS_START].as<int64_t>();
    SHARP_TRY(c->reserve_pinned((size_t)(T + 1) * 8));
    memcpy(c->pinned, R.start.data(), (size_t)(T + 1) * 8);
    SHARP_TRY(h2d_staged(c, R.start_dev, c->pinned, (size_t)(T + 1) * 8));
    double *colsum_dev = nullptr;
    if (Q.normalize) {
        SHARP_TRY(c->ws[WS_COLSUM].reserve((size_t)e.n * 8));
        colsum_dev = c->ws[WS_COLSUM].as<double>();
        if (Q.normalize == 1) {
            if (!colsum_host) return set_error(SHARP_E_ARG, "normalize = 1 needs the column sums");
            SHARP_TRY(h2d(c, colsum_dev, colsum_host + e.col0, (size_t)e.n * 8));
        } /* normalize = 2: the projection launcher computes them */
    }
    const int logkind = Q.logflag ? (Q.logkind ? Q.logkind : 2) : 0;

    // ---- K1: projection of every cell for all K members ----
    const size_t np = (size_t)n * p;
    SHARP_TRY(c->ws[WS_PROJ].reserve((size_t)K * np * 8));
    R.proj = c->ws[WS_PROJ].as<double>();
    SHARP_TRY(launch_rp_project(c, e, R.src_dev, n, colsum_dev, Q.normalize != 2, Q.normalize, Q.norm_mul, logkind, Q.round_digits, rm, R.proj));

    // ---- K2 input: unit rows ----
    const int ldu = R.ldu = ldu_of(p);
    SHARP_TRY(c->ws[WS_U].reserve((size_t)K * n * ldu * 8));
    R.U = c->ws[WS_U].as<double>();
    SHARP_TRY(launch_unit_rows(c, R.proj, (int64_t)K * n, p, ldu, R.U));

    // ---- per-(member, block) problems: parameters and small output tables ----
    sharp_hc_params indp = Q.hc;
    indp.n_cluster = Q.ind_n_cluster;
    if (Q.block_max_n > 0) indp.max_n = Q.block_max_n; /* SHARP_fpart: `maxN.cluster = 40` inside the worker (R/SHARP_unlimited2.R:421) */
    R.ind = to_dev(indp);
    if (R.ind.n_cluster != 0 && R.ind.n_cluster < 2)
        return set_error(SHARP_E_RSTOP, "The given N.cluster is less than 2, which is not suitable for clustering!");
    R.kcap_ind = R.ind.n_cluster ? R.ind.n_cluster : R.ind.max_n;
    R.nested = std::min(R.kcap_ind, R.max_bn - 1) <= NESTED_MAXK_HOST;
    const int maxlev = R.maxlev = R.ind.n_cluster ? 1 : std::max(1, R.ind.max_n - R.ind.min_n + 1);
    const int nprob = R.nprob = K * T;
    SHARP_TRY(c->ws[WS_ENRP].reserve((size_t)K * n * 4));
    R.enrp = c->ws[WS_ENRP].as<int32_t>();
    size_t ibytes = bump_size({(size_t)K * n * 4, (size_t)K * n * 4, (size_t)nprob * 8 * 4});
    size_t dbytes = bump_size({(size_t)K * n * 8, (size_t)nprob * maxlev * 8, (size_t)nprob * maxlev * 8, (size_t)nprob * 8});
    SHARP_TRY(c->ws[WS_HC_INT].reserve(ibytes));
    SHARP_TRY(c->ws[WS_HC_DBL].reserve(dbytes));
    Bump bi(c->ws[WS_HC_INT].ptr, ibytes), bd(c->ws[WS_HC_DBL].ptr, dbytes);
    R.ia_all = bi.take<int>((size_t)K * n);
    R.ib_all = bi.take<int>((size_t)K * n);
    R.meta_all = bi.take<int>((size_t)nprob * 8);
    R.crit_all = bd.take<double>((size_t)K * n);
    R.msil_all = bd.take<double>((size_t)nprob * maxlev);
    R.chind_all = bd.take<double>((size_t)nprob * maxlev);
    R.maxsil_all = bd.take<double>(nprob);
    SHARP_CUDA(cudaMemsetAsync(R.meta_all, 0, (size_t)nprob * 8 * 4, c->stream));
    return 0;
}

// K2-K6 for every (member, block) problem of the given parts, on the stream of `g`.  The distance matrices live in
// g's workspace; the problems are processed in waves sized from the memory budget (normally one wave).
static int run_blocks(sharp_ctx *g, PartRun *const *parts, int np_parts) {
    if (np_parts <= 0) return 0;
    const PartRun &R0 = *parts[0];
    int nprob = 0, max_bn = 0;
    for (int i = 0; i < np_parts; i++) {
        const PartRun &R = *parts[i];
        if (R.ldu != R0.ldu || R.p != R0.p || R.nested != R0.nested || R.maxlev != R0.maxlev || R.kcap_ind != R0.kcap_ind ||
            R.ind.hmethod != R0.ind.hmethod)
            return set_error(SHARP_E_ARG, "run_blocks: the parts of a group must share the ranM matrices and parameters");
        nprob += R.nprob;
        max_bn = std::max(max_bn, R.max_bn);
    }
    const int ldu = R0.ldu, p = R0.p;
    const HcParamsDev ind = R0.ind;
    // wave size from the distance-matrix budget.  The free-memory query is a driver call that can block for a long time
    // while other streams are busy (SHARP_B200_TRACE showed it): it is only made when the workspaces this context already
    // owns are smaller than its cap, i.e. while they are still growing.
    const size_t have = g->ws[WS_D].cap + g->ws[WS_DW].cap + g->ws[WS_HC_E].cap;
    const size_t cap_b = (size_t)g->block_budget_gb << 30;
    size_t budget = cap_b;
    if (have < cap_b && g->ws_budget_seen != have + 1) {
        size_t free_b = 0, total_b = 0;
        SLOW("cudaMemGetInfo", SHARP_CUDA(cudaMemGetInfo(&free_b, &total_b)));
        /* top-level contexts on other streams of the same device run concurrently: share the free memory between them */
        const int live = std::max(1, g_live_ctx.load());
        budget = std::min<size_t>(cap_b, (size_t)(free_b * 0.6 / live) + have);
        g->ws_budget = budget;
    } else if (have < cap_b) budget = g->ws_budget;
    const size_t per_prob = (size_t)max_bn * ld_of(max_bn) * 8;
    /* round-parallel agglomeration: a third (smaller) matrix per problem, and the distance kernel writes D only */
    const bool fast = hclust_fast_ok(max_bn, ind.hmethod);
    const size_t ecap = fast ? (((size_t)(0.8 * max_bn * ld_of(max_bn)) + 31) & ~(size_t)31) : 0;
    const size_t per_all = 2 * per_prob + ecap * 8;
    int wave_probs = (int)std::min<size_t>((size_t)nprob, std::max<size_t>(1, budget / per_all));
    if (fast) {
        /* One problem per SM and wave (125 of the benchmark's 2000-cell blocks on 148 SMs), evened out over the waves.
           The agglomeration kernel could keep two problems per SM resident, but a lone problem already streams at the
           per-SM share of HBM bandwidth (ncu: 0.13 ms per problem at 125 per launch, 0.15 ms at 297), and the shorter
           launches let the other lane's stages slot in between them: the whole job ran 4 % faster on B200 this way. */
        static const int per_sm = getenv("SHARP_WAVE_PER_SM") ? std::max(1, atoi(getenv("SHARP_WAVE_PER_SM"))) : 1; /* development switch */
        wave_probs = std::min(wave_probs, std::max(1, g->sm_count * per_sm));
        const int nw = (nprob + wave_probs - 1) / wave_probs;
        wave_probs = (nprob + nw - 1) / nw;
    }
    SLOW("reserve WS_D", SHARP_TRY(g->ws[WS_D].reserve(per_prob * wave_probs)));
    SLOW("reserve WS_DW", SHARP_TRY(g->ws[WS_DW].reserve(per_prob * wave_probs)));
    if (ecap) SHARP_TRY(g->ws[WS_HC_E].reserve(ecap * 8 * wave_probs));
    /* the budget computed above stays valid as long as the workspaces do not change */
    g->ws_budget_seen = g->ws[WS_D].cap + g->ws[WS_DW].cap + g->ws[WS_HC_E].cap + 1;
    double *Dall = g->ws[WS_D].as<double>(), *Dwall = g->ws[WS_DW].as<double>(), *Eall = g->ws[WS_HC_E].as<double>();
    const size_t sweep_scr = R0.nested ? sweep_nested_scratch_bytes(max_bn, ldu, ind) : sweep_exact_scratch_bytes(max_bn, ldu);
    SLOW("reserve WS_SWEEP_SCRATCH", SHARP_TRY(g->ws[WS_SWEEP_SCRATCH].reserve(sweep_scr * wave_probs)));
    // descriptors of ALL problems (part-major, then block, then member), built once
    const int nwaves = (nprob + wave_probs - 1) / wave_probs;
    size_t desc = bump_size({(size_t)nprob * sizeof(GemmProb), (size_t)(nprob + nwaves + 2) * 4, (size_t)nprob * sizeof(HcProb),
                             (size_t)nprob * sizeof(SweepOut), (size_t)nprob * sizeof(HcParamsDev)});
    SLOW("blocks reserve_pinned", SHARP_TRY(g->reserve_pinned(desc)));
    SLOW("reserve WS_GDESC", SHARP_TRY(g->ws[WS_GDESC].reserve(desc)));
    Bump hb(g->pinned, desc), db(g->ws[WS_GDESC].ptr, desc);
    GemmProb *gp = hb.take<GemmProb>(nprob);
    int *tp = hb.take<int>(nprob + nwaves + 2);
    HcProb *hp = hb.take<HcProb>(nprob);
    SweepOut *so = hb.take<SweepOut>(nprob);
    HcParamsDev *pp = hb.take<HcParamsDev>(nprob);
    GemmProb *gpd = db.take<GemmProb>(nprob);
    int *tpd = db.take<int>(nprob + nwaves + 2);
    HcProb *hpd = db.take<HcProb>(nprob);
    SweepOut *sod = db.take<SweepOut>(nprob);
    HcParamsDev *ppd = db.take<HcParamsDev>(nprob);
    struct Wave { int q0, nq, tiles, tp_off; };
    std::vector<Wave> waves;
    {
        int q = 0;
        for (int i = 0; i < np_parts; i++) {
            const PartRun &R = *parts[i];
            for (int t = 0; t < R.T; t++)
                for (int k = 0; k < R.K; k++) {
                    const int nq = (int)(R.start[t + 1] - R.start[t]);
                    const int ld = ld_of(nq);
                    const int slot = q % wave_probs;
                    const size_t row0 = (size_t)k * R.n + R.start[t];
                    const int ql = t * R.K + k; /* index of the problem inside its part */
                    gp[q].U = R.U + row0 * ldu;
                    gp[q].n = nq;
                    gp[q].ld = ld;
                    gp[q].D = Dall + (size_t)slot * (per_prob / 8);
                    gp[q].Dw = fast ? nullptr : Dwall + (size_t)slot * (per_prob / 8);
                    hp[q].n = nq; hp[q].ld = ld; hp[q].D = gp[q].D; hp[q].Dw = Dwall + (size_t)slot * (per_prob / 8);
                    hp[q].E = ecap ? Eall + (size_t)slot * ecap : nullptr; hp[q].ecap = (long long)ecap;
                    hp[q].ia = R.ia_all + row0; hp[q].ib = R.ib_all + row0; hp[q].crit = R.crit_all + row0;
                    hp[q].Y = gp[q].U; hp[q].p = p; hp[q].ldy = ldu; hp[q].status = 0; hp[q].fallback = 0;
                    so[q].f = R.enrp + row0; so[q].v = nullptr;
                    so[q].msil = R.msil_all + (size_t)ql * R.maxlev; so[q].chind = R.chind_all + (size_t)ql * R.maxlev;
                    so[q].meta = R.meta_all + (size_t)ql * 8; so[q].maxsil = R.maxsil_all + ql;
                    pp[q] = ind;
                    q++;
                }
        }
        int tpo = 0;
        for (int q0 = 0; q0 < nprob; q0 += wave_probs) {
            Wave w;
            w.q0 = q0;
            w.nq = std::min(wave_probs, nprob - q0);
            w.tp_off = tpo;
            int tiles = 0;
            tp[tpo] = 0;
            for (int j = 0; j < w.nq; j++) {
                tiles += corrdist_tiles(gp[q0 + j].n);
                tp[tpo + j + 1] = tiles;
            }
            w.tiles = tiles;
            tpo += w.nq + 1;
            waves.push_back(w);
        }
    }
    SLOW("blocks desc copy", SHARP_TRY(h2d_staged(g, g->ws[WS_GDESC].ptr, g->pinned, hb.off)));
    for (const Wave &w : waves) {
        SLOW("launch corrdist", SHARP_TRY(launch_corrdist_batched(g, gpd + w.q0, tpd + w.tp_off, w.nq, w.tiles, ldu)));
        SLOW("launch hclust", SHARP_TRY(launch_hclust(g, hpd + w.q0, w.nq, max_bn, ind.hmethod, 1)));
        if (R0.nested)
            SLOW("launch sweep", SHARP_TRY(launch_sweep_nested(g, hpd + w.q0, sod + w.q0, w.nq, max_bn, ldu, ind, g->ws[WS_SWEEP_SCRATCH].as<double>(), sweep_scr)));
        else
            SHARP_TRY(launch_sweep_exact(g, hpd + w.q0, sod + w.q0, w.nq, max_bn, p, ppd + w.q0, R0.maxlev, std::max(R0.kcap_ind, 2),
                                         g->ws[WS_SWEEP_SCRATCH].as<double>(), sweep_scr));
    }
    return 0;
}

// everything after the block clusterings, without waiting for the device
static int part_back(PartRun &R) {
    sharp_ctx *c = R.c;
    const int64_t n = R.n;
    const int p = R.p, K = R.K, T = R.T, nprob = R.nprob;
    const size_t np = (size_t)n * p;
    const sharp_run_params &Q = R.Q;
    prof_begin(c, KID_MISC);
    colour_wrap_kernel<<<grid1d((size_t)K * n, 256), 256, 0, c->stream>>>(R.enrp, (int64_t)K * n);
    prof_end(c);
    // ---- enE / K ----
    SHARP_TRY(c->ws[WS_E1].reserve(np * 8));
    R.E1 = c->ws[WS_E1].as<double>();
    prof_begin(c, KID_ENE);
    ene_kernel<<<grid1d(np, 256), 256, 0, c->stream>>>(R.proj, (int64_t)np, K, R.E1);
    prof_end(c);
    // pinned mirrors of every status word of this run (read by part_finish)
    SHARP_TRY(c->reserve_pinned(((size_t)nprob * 8 + T + 8) * 4));
    R.h_meta = reinterpret_cast<int *>(c->pinned);
    R.h_wst = R.h_meta + (size_t)nprob * 8;
    R.h_sst = R.h_wst + T;
    R.h_nu = R.h_sst + 1;
    R.h_sst[0] = 0;
    R.h_nu[0] = 0;
    SHARP_TRY(d2h(c, R.h_meta, R.meta_all, (size_t)nprob * 8 * 4));
    // ---- wMetaC per block ----
    sharp_hc_params wp = Q.hc;
    wp.n_cluster = Q.large ? Q.enp_n_cluster : Q.n_cluster;
    SLOW("back wmetac_dev", SHARP_TRY(wmetac_dev(c, R.enrp, n, K, T, R.start.data(), R.start_dev, 40, wp, &R.W)));
    SHARP_TRY(d2h(c, R.h_wst, R.W.A.status, (size_t)T * 4));
    // ---- sharded run: the cross-block stage sees the whole matrix on every rank ----
    const int64_t na = R.na;
    const int T_all = R.T_all;
    WmArgs Afull = R.W.A;            /* what launch_sm_codes reads: start, ucount, status, fcode */
    const double *E1x = R.E1;
    if (R.sharded) {
        const int W = c->comm_world;
        std::vector<int64_t> bcell((size_t)W), bblk((size_t)W), brow((size_t)W);
        for (int r = 0; r < W; r++) { /* the same deterministic split on every rank */
            int a0, a1;
            shard_range(T_all, r, W, Q.shard_rotate, &a0, &a1);
            bcell[r] = (R.start_all[a1] - R.start_all[a0]) * 4;
            bblk[r] = (int64_t)(a1 - a0) * 4;
            brow[r] = (R.start_all[a1] - R.start_all[a0]) * (int64_t)p * 8;
        }
        /* ranks own contiguous block ranges, but a rotated split starts at another rank than 0: segments are gathered in
           RANK order into staging and copied to their BLOCK position */
        size_t ib3 = bump_size({(size_t)na * 4, (size_t)T_all * 4, (size_t)T_all * 4, (size_t)(T_all + 1) * 8, (size_t)na * 4,
                                (size_t)T_all * 4, (size_t)T_all * 4});
        SHARP_TRY(c->ws[WS_TMP2].reserve(ib3));
        Bump b3(c->ws[WS_TMP2].ptr, ib3);
        int *fcode_all = b3.take<int>(na), *ucount_all = b3.take<int>(T_all), *status_all = b3.take<int>(T_all);
        R.start_all_dev = b3.take<int64_t>(T_all + 1);
        int *g_fcode = b3.take<int>(na), *g_ucount = b3.take<int>(T_all), *g_status = b3.take<int>(T_all);
        SHARP_TRY(c->ws[WS_TMP3].reserve((size_t)na * p * 8 * 2));
        R.E1_all = c->ws[WS_TMP3].as<double>();
        double *g_rows = R.E1_all + (size_t)na * p;
        SHARP_TRY(c->reserve_pinned((size_t)(T_all + 1) * 8));
        memcpy(c->pinned, R.start_all.data(), (size_t)(T_all + 1) * 8);
        SHARP_TRY(h2d_staged(c, R.start_all_dev, c->pinned, (size_t)(T_all + 1) * 8));
        prof_begin(c, KID_COMM);
        int rc = comm_allgatherv_dev(c, R.W.A.fcode, g_fcode, bcell.data(), c->stream);
        if (!rc) rc = comm_allgatherv_dev(c, R.W.A.ucount, g_ucount, bblk.data(), c->stream);
        if (!rc) rc = comm_allgatherv_dev(c, R.W.A.status, g_status, bblk.data(), c->stream);
        if (!rc) rc = comm_allgatherv_dev(c, R.E1, g_rows, brow.data(), c->stream);
        prof_end(c);
        c->launches--; /* collectives, not kernels of this library */
        SHARP_TRY(rc);
        int64_t oc = 0, ob = 0;
        for (int r = 0; r < W; r++) { /* rank order -> block order */
            int a0, a1;
            shard_range(T_all, r, W, Q.shard_rotate, &a0, &a1);
            const int64_t cells = R.start_all[a1] - R.start_all[a0];
            SHARP_CUDA(cudaMemcpyAsync(fcode_all + R.start_all[a0], g_fcode + oc, (size_t)cells * 4, cudaMemcpyDeviceToDevice, c->stream));
            SHARP_CUDA(cudaMemcpyAsync(R.E1_all + (size_t)R.start_all[a0] * p, g_rows + (size_t)oc * p, (size_t)cells * p * 8, cudaMemcpyDeviceToDevice, c->stream));
            SHARP_CUDA(cudaMemcpyAsync(ucount_all + a0, g_ucount + ob, (size_t)(a1 - a0) * 4, cudaMemcpyDeviceToDevice, c->stream));
            SHARP_CUDA(cudaMemcpyAsync(status_all + a0, g_status + ob, (size_t)(a1 - a0) * 4, cudaMemcpyDeviceToDevice, c->stream));
            oc += cells;
            ob += a1 - a0;
        }
        Afull.start = R.start_all_dev;
        Afull.ucount = ucount_all;
        Afull.status = status_all;
        Afull.fcode = fcode_all;
        Afull.ncells = na;
        E1x = R.E1_all;
    }
    const int Tx = R.sharded ? T_all : T;
    const int64_t *out_row = R.sharded ? R.src_all_dev : R.src_dev;
    // ---- labels / sMetaC ----
    SHARP_TRY(c->ws[WS_LABELS].reserve((size_t)na * 4));
    R.labels_dev = c->ws[WS_LABELS].as<int>();
    R.coloff_dev = nullptr;
    R.nc_dev = nullptr;
    if (Tx == 1) {
        if (!Q.large) SHARP_TRY(launch_sm_relabel(c, n, R.W.A.finalc, nullptr, 0, R.src_dev, R.labels_dev)); /* finalC ids */
        else SHARP_TRY(launch_sm_relabel(c, n, R.W.A.fcode, nullptr, 1, R.src_dev, R.labels_dev)); /* position in unique(fColor) */
        SHARP_TRY(d2h(c, R.h_nu, R.W.A.ucount, 4));
    } else {
        const int capS = Tx * R.W.A.capU;
        size_t ib2 = bump_size({(size_t)(Tx + 1) * 4, 64, 64, (size_t)na * 4, (size_t)na * 4, (size_t)(capS + 1) * 4});
        SHARP_TRY(c->ws[WS_TMP0].reserve(ib2));
        Bump b2(c->ws[WS_TMP0].ptr, ib2);
        R.coloff_dev = b2.take<int>(Tx + 1);
        R.nc_dev = b2.take<int>(1);
        int *st_dev = b2.take<int>(1);
        int *code = b2.take<int>(na);
        int *corder = b2.take<int>(na);
        int *coff = b2.take<int>(capS + 1);
        SHARP_TRY(launch_sm_codes(c, Afull, Tx, R.coloff_dev, R.nc_dev, st_dev, code, corder, coff));
        if (Q.skip_smetac) {
            /* SHARP_fpart (R/SHARP_unlimited2.R:477-531): the part ends after the per-block wMetaC; fColor = "<finalC>en<t>"
               is returned as the position of the (block, meta-cluster) pair in the blocks' cluster lists -- an injective
               code, which is all paste() / unique() downstream look at -- un-shuffled like fColor[reind] = fColor */
            R.skipped_smetac = true;
            SHARP_TRY(launch_sm_relabel(c, na, code, nullptr, 1, out_row, R.labels_dev));
        } else {
            sharp_hc_params sp = Q.hc;
            sp.n_cluster = Q.n_cluster;
            SLOW("back smetac_dev", SHARP_TRY(smetac_dev(c, capS, p, R.nc_dev, st_dev, E1x, corder, coff, nullptr, na, sp, &R.B)));
            SHARP_TRY(d2h(c, R.h_sst, R.B.status, 4));
            SHARP_TRY(launch_sm_relabel(c, na, code, R.B.tf, 0, out_row, R.labels_dev));
        }
    }
    // viE = enE/K, un-shuffled; kept on the device for sharp_centroids
    SHARP_TRY(c->ws[WS_VIEU].reserve((size_t)na * p * 8));
    R.vieu = c->ws[WS_VIEU].as<double>();
    prof_begin(c, KID_ENE);
    scatter_rows_kernel<<<(unsigned)na, 128, 0, c->stream>>>(E1x, na, p, out_row, R.vieu);
    prof_end(c);
    c->last_n = na;
    c->last_p = p;
    c->last_K = K;
    return 0;
}

// statuses -> errors (after the caller synchronised the part's stream)
static int part_status(const PartRun &R) {
    for (int q = 0; q < R.nprob; q++)
        if (R.h_meta[(size_t)q * 8 + 3] != 0) return rstop(R.h_meta[(size_t)q * 8 + 3], "getrowColor (block clustering)");
    for (int t = 0; t < R.T; t++)
        if (R.h_wst[t] != 0) return rstop(R.h_wst[t], "wMetaC");
    if (R.h_sst[0] != 0) return rstop(R.h_sst[0], "sMetaC");
    return 0;
}

static int part_finish(PartRun &R, int32_t *labels_out, double *vie_out, double *x0_out, int *x0_cols, int max_x0_cols) {
    sharp_ctx *c = R.c;
    const int64_t n = R.na;
    const int p = R.p, T = R.sharded ? R.T_all : R.T;
    if (R.sharded && x0_out) return set_error(SHARP_E_ARG, "a sharded run does not produce x0 (call with forview = FALSE)");
    SHARP_TRY(d2h(c, labels_out, R.labels_dev, (size_t)n * 4));
    SHARP_TRY(sync(c));
    SHARP_TRY(part_status(R));
    int ncol_x0 = 0;
    std::vector<int> colmap_host;
    if (T == 1) ncol_x0 = R.h_nu[0];
    else if (R.skipped_smetac) { /* SHARP_fpart returns no x0 (its sx0 stays local, R/SHARP_unlimited2.R:501-514) */
        if (x0_out) return set_error(SHARP_E_ARG, "skip_smetac: the block-level run has no x0 output");
        SHARP_TRY(d2h(c, &ncol_x0, R.nc_dev, 4));
        SHARP_TRY(sync(c));
    } else if (x0_out || x0_cols) {
        int nC = 0;
        SHARP_TRY(d2h(c, &nC, R.nc_dev, 4));
        SHARP_TRY(sync(c));
        colmap_host.resize(nC);
        SHARP_TRY(d2h(c, colmap_host.data(), R.B.tf, (size_t)nC * 4));
        SHARP_TRY(sync(c));
        for (int v : colmap_host) ncol_x0 = std::max(ncol_x0, v);
    }
    if (x0_cols) *x0_cols = ncol_x0;
    if (x0_out) {
        if (ncol_x0 > max_x0_cols) return set_error(SHARP_E_NOMEM, "x0 needs %d columns but the buffer has %d", ncol_x0, max_x0_cols);
        SHARP_TRY(c->ws[WS_X0].reserve((size_t)n * ncol_x0 * 8));
        double *x0d = c->ws[WS_X0].as<double>();
        SHARP_CUDA(cudaMemsetAsync(x0d, 0, (size_t)n * ncol_x0 * 8, c->stream));
        if (T == 1) {
            SHARP_TRY(launch_wmetac_x0(c, R.W.A, R.W.outs_dev, T, R.W.max_block_n, nullptr, nullptr, R.src_dev, x0d, ncol_x0));
        } else {
            /* colmap = tf - 1 (0-based output column of every block-level cluster) */
            SHARP_TRY(c->ws[WS_TMP1].reserve(colmap_host.size() * 4 + 64));
            int *cm = c->ws[WS_TMP1].as<int>();
            for (int &v : colmap_host) v -= 1;
            SHARP_TRY(c->reserve_pinned(colmap_host.size() * 4 + 64));
            memcpy(c->pinned, colmap_host.data(), colmap_host.size() * 4);
            SHARP_TRY(h2d_staged(c, cm, c->pinned, colmap_host.size() * 4));
            SHARP_TRY(launch_wmetac_x0(c, R.W.A, R.W.outs_dev, T, R.W.max_block_n, R.coloff_dev, cm, R.src_dev, x0d, ncol_x0));
        }
        SHARP_TRY(d2h(c, x0_out, x0d, (size_t)n * ncol_x0 * 8));
    }
    if (vie_out) SHARP_TRY(d2h(c, vie_out, R.vieu, (size_t)n * p * 8));
    SHARP_TRY(sync(c));
    return 0;
}

static int run_core(sharp_ctx *c, const sharp_expr_dev &e, const double *colsum_host, const sharp_rm_dev &rm,
                    const int64_t *reind, const sharp_run_params &Q, int32_t *labels_out, double *vie_out, double *x0_out,
                    int *x0_cols, int max_x0_cols) {
    Trace tr(c);
    PartRun R;
    { NvtxRange r("sharp front (colSums, projection, unit rows)"); SHARP_TRY(part_front(R, c, e, colsum_host, rm, reind, Q)); }
    tr.mark("front", true);
    PartRun *one = &R;
    { NvtxRange r("sharp blocks (distances, agglomeration, sweep)"); SHARP_TRY(run_blocks(c, &one, 1)); }
    tr.mark("blocks", true);
    { NvtxRange r("sharp back (wMetaC, sMetaC, relabel)"); SHARP_TRY(part_back(R)); }
    tr.mark("back", true);
    NvtxRange rf("sharp finish (outputs)");
    SHARP_TRY(part_finish(R, labels_out, vie_out, x0_out, x0_cols, max_x0_cols));
    tr.mark("finish");
    return 0;
}

// =====================================================================================================
// SHARP_unlimited's loop over parts as ONE call: groups of parts share the block-clustering launches
// =====================================================================================================
static int make_child(sharp_ctx *parent, sharp_ctx **out) {
    sharp_ctx *c = new sharp_ctx();
    c->device = parent->device;
    c->sm_count = parent->sm_count;
    c->parent = parent;
    c->rp_variant = parent->rp_variant;
    c->comm = parent->comm; c->comm_rank = parent->comm_rank; c->comm_world = parent->comm_world;
    c->block_budget_gb = parent->block_budget_gb;
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_ready, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_blocks, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_done, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev1);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_up, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        delete c;
        return set_error(SHARP_E_CUDA, "creating a sub-context: %s", cudaGetErrorString(e));
    }
    c->prof_on = parent->prof_on;
    *out = c;
    return 0;
}

void destroy_ctx_resources(sharp_ctx *c) {
    for (sharp_ctx *s : c->subs) {
        destroy_ctx_resources(s);
        delete s;
    }
    c->subs.clear();
    cudaStreamSynchronize(c->stream);
    prof_collect(c);
    for (cudaEvent_t e : c->prof_pool) cudaEventDestroy(e);
    c->prof_pool.clear();
    for (auto &b : c->ws) b.release();
    if (c->arena) cudaFreeHost(c->arena);
    for (int h = 0; h < 2; h++)
        if (c->ev_half[h]) { cudaEventDestroy(c->ev_half[h]); c->ev_half[h] = nullptr; }
    if (c->h_labels) cudaFreeHost(c->h_labels);
    if (c->h_gather) cudaFreeHost(c->h_gather);
    c->h_gather = nullptr;
    c->arena = nullptr;
    c->h_labels = nullptr;
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->ev_ready) cudaEventDestroy(c->ev_ready);
    if (c->ev_blocks) cudaEventDestroy(c->ev_blocks);
    if (c->ev_done) cudaEventDestroy(c->ev_done);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_up) cudaEventDestroy(c->ev_up);
    if (c->up_stream) { cudaStreamSynchronize(c->up_stream); cudaStreamDestroy(c->up_stream); }
    if (c->rm_stream) { cudaStreamSynchronize(c->rm_stream); cudaStreamDestroy(c->rm_stream); }
    if (c->rm_stage) cudaFreeHost(c->rm_stage);
    c->rm_stage = nullptr;
    cudaStreamDestroy(c->stream);
}

// R/SHARP.R:816-843 on the labels of one part: clusters with fewer than `thre` cells are merged into the smallest such
// id, then clusterID = match(y, unique(y)).  Returns the number of clusters; lab is rewritten in place (1-based).
static int host_merge_relabel(int32_t *lab, int64_t n, int thre, std::vector<int> &coff, std::vector<int> &corder) {
    int mx = 0;
    for (int64_t i = 0; i < n; i++) mx = std::max(mx, lab[i]);
    std::vector<int64_t> cnt((size_t)mx + 1, 0);
    for (int64_t i = 0; i < n; i++) cnt[lab[i]]++;
    if (thre > 0) {
        int smallest = 0;
        for (int v = 1; v <= mx; v++)
            if (cnt[v] > 0 && cnt[v] < thre) { smallest = v; break; }
        if (smallest)
            for (int64_t i = 0; i < n; i++)
                if (cnt[lab[i]] < thre) lab[i] = smallest;
    }
    std::vector<int> code((size_t)mx + 1, 0);
    int next = 0;
    for (int64_t i = 0; i < n; i++) {
        int &cd = code[lab[i]];
        if (!cd) cd = ++next;
        lab[i] = cd;
    }
    coff.assign((size_t)next + 1, 0);
    for (int64_t i = 0; i < n; i++) coff[lab[i]]++;
    for (int q = 0; q < next; q++) coff[q + 1] += coff[q];
    corder.resize(n);
    std::vector<int> fill(coff.begin(), coff.end() - 1);
    for (int64_t i = 0; i < n; i++) corder[fill[lab[i] - 1]++] = (int)i;
    return next;
}

struct GroupRun {
    std::vector<int> idx;          // part indices
    std::vector<PartRun> runs;
    sharp_ctx *blocks = nullptr;
    cudaStream_t up = nullptr;     // upload stream of the run (null: copies go on each part's own stream)
    std::vector<sharp_ctx *> subs;
    // centroid staging per part
    std::vector<int> nclust;
    std::vector<double *> h_cen;
    std::vector<int64_t *> h_cnt;
};

static int group_issue(GroupRun &G, sharp_part *parts, int m, const sharp_rm_dev &rm, const sharp_run_params &Q) {
    static const bool trace = getenv("SHARP_B200_TRACE") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    double t_up = 0, t_front = 0, t_blocks = 0, t_back = 0;
    auto lap = [&](double &acc, std::chrono::steady_clock::time_point &from) {
        const auto now = std::chrono::steady_clock::now();
        acc += std::chrono::duration<double, std::milli>(now - from).count();
        from = now;
    };
    auto tl = t0;
    const int np = (int)G.idx.size();
    G.runs.assign(np, PartRun());
    for (int j = 0; j < np; j++) {
        sharp_part &P = parts[G.idx[j]];
        sharp_ctx *s = G.subs[j];
        sharp_expr_dev e;
        if (P.dev) {
            e = *P.dev;
            e.owned = false;
        } else if (s->pf_part == G.idx[j] && s->pf_src == (P.dense ? (const void *)P.dense : (const void *)P.val)) {
            /* copied in by group_prefetch while the previous group was running (or by sharp_parts_prefetch) */
            s->pf_part = -1;
            e.device = s->device; e.m = m; e.n = P.n; e.owned = false;
            if (P.dense) e.dense = s->ws[WS_EX_A].as<double>();
            else {
                e.nnz = P.colptr[P.n];
                e.colptr = s->ws[WS_EX_A].as<int64_t>();
                e.rowidx = s->ws[WS_EX_B].as<int32_t>();
                e.val = s->ws[WS_EX_C].as<double>();
            }
            SHARP_CUDA(cudaStreamWaitEvent(s->stream, s->ev_up, 0));
        } else {
            /* the sub-context's buffers are about to be overwritten: a look-ahead copy that was never used (the caller
               changed the grouping between sharp_parts_prefetch and this call) must not be matched by a later group */
            s->pf_part = -1;
            bool compacted = false;
            if (P.sharded && s->comm && s->comm_world > 1 && Q.large && P.colptr && P.reind && P.n < 100000) {
                /* a block-sharded, SHUFFLED part on host buffers: this rank needs the columns reind[pos0 .. pos0 + n_loc) only
                   -- 1 / world of the part.  They are gathered on the host (a few threads) into pinned staging, in shuffled
                   order, and only those bytes cross PCIe. */
                std::vector<int64_t> st;
                make_blocks(P.n, 1, Q.partition_ncells, st);
                const int T = (int)st.size() - 1;
                if (T >= s->comm_world) {
                    int t0, t1;
                    shard_range(T, s->comm_rank, s->comm_world, G.idx[j], &t0, &t1);
                    const int64_t pos0 = st[t0], nloc = st[t1] - st[t0];
                    int64_t nnz = 0;
                    for (int64_t i = 0; i < nloc; i++) {
                        const int64_t src = P.reind[pos0 + i] - 1;
                        if (src < 0 || src >= P.n) return set_error(SHARP_E_ARG, "reind is not a permutation of 1..n");
                        nnz += P.colptr[src + 1] - P.colptr[src];
                    }
                    const size_t o_ri = (((size_t)(nloc + 1) * 8) + 63) & ~(size_t)63, o_v = (o_ri + (size_t)nnz * 4 + 63) & ~(size_t)63;
                    const size_t bytes = o_v + (size_t)nnz * 8 + 64;
                    if (bytes > s->h_gather_cap) {
                        if (s->h_gather) { SHARP_CUDA(cudaStreamSynchronize(G.up ? G.up : s->stream)); cudaFreeHost(s->h_gather); }
                        s->h_gather = nullptr;
                        s->h_gather_cap = 0;
                        const size_t want = bytes + bytes / 4;
                        SHARP_CUDA(cudaMallocHost((void **)&s->h_gather, want));
                        s->h_gather_cap = want;
                    } else if (s->ev_up) SHARP_CUDA(cudaEventSynchronize(s->ev_up)); /* the previous copy out of this staging is done */
                    int64_t *cp = reinterpret_cast<int64_t *>(s->h_gather);
                    int32_t *ri = reinterpret_cast<int32_t *>(s->h_gather + o_ri);
                    double *xv = reinterpret_cast<double *>(s->h_gather + o_v);
                    cp[0] = 0;
                    for (int64_t i = 0; i < nloc; i++) {
                        const int64_t src = P.reind[pos0 + i] - 1;
                        cp[i + 1] = cp[i] + (P.colptr[src + 1] - P.colptr[src]);
                    }
                    const int nth = (int)std::min<int64_t>(8, std::max<int64_t>(1, nloc / 256));
                    std::vector<std::thread> th;
                    for (int t = 0; t < nth; t++)
                        th.emplace_back([&, t]() {
                            for (int64_t i = nloc * t / nth; i < nloc * (t + 1) / nth; i++) {
                                const int64_t src = P.reind[pos0 + i] - 1, q0 = P.colptr[src], len = P.colptr[src + 1] - q0;
                                memcpy(ri + cp[i], P.rowidx + q0, (size_t)len * 4);
                                memcpy(xv + cp[i], P.val + q0, (size_t)len * 8);
                            }
                        });
                    for (auto &x : th) x.join();
                    SLOW("gathered upload_expr", SHARP_TRY(upload_expr(s, m, nloc, nullptr, cp, ri, xv, &e, true, G.up)));
                    if (G.up) {
                        SHARP_CUDA(cudaEventRecord(s->ev_up, G.up));
                        SHARP_CUDA(cudaStreamWaitEvent(s->stream, s->ev_up, 0));
                    } else SHARP_CUDA(cudaEventRecord(s->ev_up, s->stream));
                    e.compact = true;
                    e.col0 = 0;
                    e.n_total = P.n;
                    compacted = true;
                }
            }
            if (compacted) {
            } else if (G.up) { /* all uploads of a group run go through ONE stream: the copy engine shares the link between
                           streams, and the first group must not wait for the bytes of the groups behind it */
                SLOW("inline upload_expr", SHARP_TRY(upload_expr(s, m, P.n, P.dense, P.colptr, P.rowidx, P.val, &e, true, G.up)));
                SHARP_CUDA(cudaEventRecord(s->ev_up, G.up));
                SHARP_CUDA(cudaStreamWaitEvent(s->stream, s->ev_up, 0));
            } else {
                prof_begin(s, KID_H2D);
                int urc = upload_expr(s, m, P.n, P.dense, P.colptr, P.rowidx, P.val, &e, true);
                prof_end(s);
                s->launches--; /* a copy, not a kernel */
                SHARP_TRY(urc);
            }
        }
        if (e.m != rm.m) return set_error(SHARP_E_ARG, "run_parts: part %d has %d genes but ranM has %d rows", G.idx[j], e.m, rm.m);
        lap(t_up, tl);
        sharp_run_params Qp = Q;
        Qp.shard = P.sharded ? 1 : 0;          /* the blocks of a sharded part are dealt over the ranks of the communicator */
        Qp.shard_rotate = G.idx[j];
        NvtxRange rfront("sharp_run_parts: front of a part");
        SHARP_TRY(part_front(G.runs[j], s, e, nullptr, rm, P.reind, Qp));
        lap(t_front, tl);
        SHARP_CUDA(cudaEventRecord(s->ev_ready, s->stream));
    }
    for (int j = 0; j < np; j++) SHARP_CUDA(cudaStreamWaitEvent(G.blocks->stream, G.subs[j]->ev_ready, 0));
    std::vector<PartRun *> ptrs(np);
    for (int j = 0; j < np; j++) ptrs[j] = &G.runs[j];
    { NvtxRange rb("sharp_run_parts: blocks of a group"); SHARP_TRY(run_blocks(G.blocks, ptrs.data(), np)); }
    SHARP_CUDA(cudaEventRecord(G.blocks->ev_blocks, G.blocks->stream));
    lap(t_blocks, tl);
    for (int j = 0; j < np; j++) {
        sharp_ctx *s = G.subs[j];
        PartRun &R = G.runs[j];
        SHARP_CUDA(cudaStreamWaitEvent(s->stream, G.blocks->ev_blocks, 0));
        NvtxRange rback("sharp_run_parts: back of a part");
        SHARP_TRY(part_back(R));
        if ((size_t)R.na * 4 > s->h_labels_cap) {
            if (s->h_labels) cudaFreeHost(s->h_labels);
            s->h_labels = nullptr;
            s->h_labels_cap = 0;
            size_t want = ((size_t)R.na * 4 * 9 / 8 + 4095) & ~(size_t)4095;
            SHARP_CUDA(cudaMallocHost((void **)&s->h_labels, want));
            s->h_labels_cap = want;
        }
        SHARP_TRY(d2h(s, s->h_labels, R.labels_dev, (size_t)R.na * 4));
        SHARP_CUDA(cudaEventRecord(s->ev_done, s->stream));
    }
    lap(t_back, tl);
    if (trace && t_up + t_front + t_blocks + t_back > 10.0)
        fprintf(stderr, "[sharp trace group_issue] slow issue: upload %.2f ms, front %.2f ms, blocks %.2f ms, back %.2f ms\n", t_up, t_front,
                t_blocks, t_back);
    return 0;
}

// Upload look-ahead: the parts of the group that will run on these sub-contexts NEXT are copied into the sub-contexts'
// expression buffers as soon as the front stage of the part that occupies them now is done (ev_ready).  Only when
// the buffers are already large enough (growing one would free memory that is in use); otherwise group_issue uploads
// the part itself, as it does for the first groups.
static int group_prefetch(const std::vector<int> &idx, const std::vector<sharp_ctx *> &subs, sharp_part *parts, int m,
                          cudaStream_t up) {
    for (size_t j = 0; j < idx.size(); j++) {
        const sharp_part &P = parts[idx[j]];
        sharp_ctx *s = subs[j];
        s->pf_part = -1;
        if (P.dev || m <= 0 || P.n < 0 || P.sharded) continue; /* a sharded part uploads only this rank's columns (group_issue) */
        if (P.dense) {
            const size_t bytes = (size_t)m * P.n * sizeof(double);
            if (s->ws[WS_EX_A].cap < bytes) continue;
            SHARP_CUDA(cudaStreamWaitEvent(up, s->ev_ready, 0));
            SHARP_CUDA(cudaMemcpyAsync(s->ws[WS_EX_A].ptr, P.dense, bytes, cudaMemcpyHostToDevice, up));
        } else {
            if (!P.colptr || !P.rowidx || !P.val) continue;
            const int64_t nnz = P.colptr[P.n];
            if (s->ws[WS_EX_A].cap < (size_t)(P.n + 1) * 8 || s->ws[WS_EX_B].cap < (size_t)nnz * 4 || s->ws[WS_EX_C].cap < (size_t)nnz * 8) continue;
            SHARP_CUDA(cudaStreamWaitEvent(up, s->ev_ready, 0));
            SHARP_CUDA(cudaMemcpyAsync(s->ws[WS_EX_A].ptr, P.colptr, (size_t)(P.n + 1) * 8, cudaMemcpyHostToDevice, up));
            SHARP_CUDA(cudaMemcpyAsync(s->ws[WS_EX_B].ptr, P.rowidx, (size_t)nnz * 4, cudaMemcpyHostToDevice, up));
            SHARP_CUDA(cudaMemcpyAsync(s->ws[WS_EX_C].ptr, P.val, (size_t)nnz * 8, cudaMemcpyHostToDevice, up));
        }
        SHARP_CUDA(cudaEventRecord(s->ev_up, up));
        s->pf_part = idx[j];
        s->pf_src = P.dense ? (const void *)P.dense : (const void *)P.val;
    }
    return 0;
}

static int group_complete(GroupRun &G, sharp_part *parts, int small_thre, int cen_cap) {
    NvtxRange rc("sharp_run_parts: complete a group (merge, relabel, centroids)");
    const int np = (int)G.idx.size();
    G.nclust.assign(np, 0);
    G.h_cen.assign(np, nullptr);
    G.h_cnt.assign(np, nullptr);
    std::vector<int> coff, corder;
    for (int j = 0; j < np; j++) {
        sharp_part &P = parts[G.idx[j]];
        sharp_ctx *s = G.subs[j];
        PartRun &R = G.runs[j];
        SHARP_CUDA(cudaEventSynchronize(s->ev_done));
        SHARP_TRY(part_status(R));
        memcpy(P.pred, s->h_labels, (size_t)R.na * 4);
        const int thre = (R.Q.n_cluster == 0 && R.na > 10000) ? small_thre : 0;
        const int nclust = host_merge_relabel(P.pred, R.na, thre, coff, corder);
        G.nclust[j] = nclust;
        P.nclust = nclust;
        if (!P.cen) continue;
        if (nclust > cen_cap) return set_error(SHARP_E_NOMEM, "run_parts: part %d has %d clusters but cen has room for %d", G.idx[j], nclust, cen_cap);
        // centroids of viE per final cluster (colMeans of R/sMetaC.R:58-63 for the global sMetaC)
        const size_t ib = bump_size({(size_t)R.na * 4, (size_t)(nclust + 1) * 4, 64, (size_t)nclust * 8});
        SHARP_TRY(s->ws[WS_TMP0].reserve(ib));
        Bump b(s->ws[WS_TMP0].ptr, ib);
        int *corder_d = b.take<int>(R.na), *coff_d = b.take<int>(nclust + 1), *nc_d = b.take<int>(1);
        int64_t *cnt_d = b.take<int64_t>(nclust);
        SHARP_TRY(s->ws[WS_CEN].reserve((size_t)nclust * R.p * 8));
        SHARP_TRY(s->reserve_pinned((size_t)R.na * 4 + (size_t)(nclust + 2) * 4));
        int *hp = reinterpret_cast<int *>(s->pinned);
        memcpy(hp, corder.data(), (size_t)R.na * 4);
        memcpy(hp + R.na, coff.data(), (size_t)(nclust + 1) * 4);
        hp[R.na + nclust + 1] = nclust;
        SHARP_TRY(h2d_staged(s, corder_d, hp, (size_t)R.na * 4));
        SHARP_TRY(h2d_staged(s, coff_d, hp + R.na, (size_t)(nclust + 1) * 4));
        SHARP_TRY(h2d_staged(s, nc_d, hp + R.na + nclust + 1, 4));
        SHARP_TRY(launch_sm_centroids(s, R.vieu, R.p, corder_d, coff_d, nc_d, nclust, s->ws[WS_CEN].as<double>(), cnt_d));
        SHARP_TRY(s->reserve_pinned((size_t)nclust * R.p * 8 + (size_t)nclust * 8));
        G.h_cen[j] = reinterpret_cast<double *>(s->pinned);
        G.h_cnt[j] = reinterpret_cast<int64_t *>(G.h_cen[j] + (size_t)nclust * R.p);
        SHARP_TRY(d2h(s, G.h_cen[j], s->ws[WS_CEN].ptr, (size_t)nclust * R.p * 8));
        SHARP_TRY(d2h(s, G.h_cnt[j], cnt_d, (size_t)nclust * 8));
    }
    for (int j = 0; j < np; j++) {
        sharp_part &P = parts[G.idx[j]];
        sharp_ctx *s = G.subs[j];
        SHARP_TRY(sync(s));
        if (P.cen) memcpy(P.cen, G.h_cen[j], (size_t)G.nclust[j] * G.runs[j].p * 8);
        if (P.counts) memcpy(P.counts, G.h_cnt[j], (size_t)G.nclust[j] * 8);
    }
    return 0;
}

}  // namespace sharp

using namespace sharp;

// =====================================================================================================
// C ABI
// =====================================================================================================
extern "C" {

int sharp_abi_version(void) { return SHARP_B200_ABI_VERSION; }
const char *sharp_last_error(void) { return g_err; }

int sharp_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int sharp_device_info(int device, char *name, int name_len, int *sm_count, int *cc_major, int *cc_minor, size_t *total_mem) {
    cudaDeviceProp prop;
    SHARP_CUDA(cudaGetDeviceProperties(&prop, device));
    if (name && name_len > 0) {
        strncpy(name, prop.name, name_len - 1);
        name[name_len - 1] = 0;
    }
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (total_mem) *total_mem = prop.totalGlobalMem;
    return 0;
}

int sharp_ctx_create(int device, sharp_ctx **out) {
    if (!out) return set_error(SHARP_E_ARG, "null output pointer");
    /* sharp_run_parts keeps ~10 streams busy: more hardware queues than the default 8, or streams alias and copies /
       kernels of different parts serialise on false dependencies (only effective before CUDA initialises) */
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return set_error(SHARP_E_CUDA, "no CUDA device available (%s): libsharpb200 has no CPU fallback",
                         e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    if (device < 0 || device >= ndev) return set_error(SHARP_E_ARG, "device %d out of range (0..%d)", device, ndev - 1);
    SHARP_CUDA(cudaSetDevice(device));
    sharp_ctx *c = new sharp_ctx();
    c->device = device;
    cudaDeviceProp prop;
    SHARP_CUDA(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    SHARP_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    SHARP_CUDA(cudaEventCreate(&c->ev0));
    SHARP_CUDA(cudaEventCreate(&c->ev1));
    SHARP_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    *out = c;
    g_live_ctx++;
    return 0;
}

void sharp_ctx_destroy(sharp_ctx *c) {
    if (!c) return;
    g_live_ctx--;
    cudaSetDevice(c->device);
    if (c->comm) { cudaStreamSynchronize(c->stream); comm_destroy(c); }
    destroy_ctx_resources(c);
    delete c;
}

void *sharp_ctx_stream(sharp_ctx *c) { return c ? (void *)c->stream : nullptr; }
int sharp_ctx_sync(sharp_ctx *c) {
    SHARP_TRY(use(c));
    return sync(c);
}
int sharp_timer_start(sharp_ctx *c) {
    SHARP_TRY(use(c));
    SHARP_CUDA(cudaEventRecord(c->ev0, c->stream));
    return 0;
}
int sharp_timer_stop_ms(sharp_ctx *c, double *ms) {
    SHARP_TRY(use(c));
    SHARP_CUDA(cudaEventRecord(c->ev1, c->stream));
    SHARP_CUDA(cudaEventSynchronize(c->ev1));
    float f = 0;
    SHARP_CUDA(cudaEventElapsedTime(&f, c->ev0, c->ev1));
    if (ms) *ms = f;
    return 0;
}
int64_t sharp_ctx_launch_count(sharp_ctx *c) {
    if (!c) return 0;
    int64_t n = c->launches;
    for (sharp_ctx *s : c->subs) n += s->launches;
    return n;
}
int sharp_ctx_set_block_budget(sharp_ctx *c, int gigabytes) {
    if (!c || gigabytes < 1) return set_error(SHARP_E_ARG, "set_block_budget: bad arguments");
    c->block_budget_gb = gigabytes;
    return 0;
}

static const char *const g_kernel_names[KID_COUNT] = {
    "rp_project", "colsum", "unit_rows", "corrdist", "hclust", "hclust_small", "sweep_nested", "sweep_exact",
    "wm_weights", "wm_similarity", "wmetac_misc", "sm_centroids", "smetac_misc", "ene_scatter", "misc", "h2d_expr", "nccl_allgather"};

int sharp_ctx_set_serial(sharp_ctx *c, int on) {
    if (!c) return set_error(SHARP_E_ARG, "null context");
    c->serial = on != 0;
    return 0;
}

int sharp_ctx_set_rp_variant(sharp_ctx *c, int variant) {
    if (!c) return set_error(SHARP_E_ARG, "null context");
    if (variant < 0 || variant > 3) return set_error(SHARP_E_ARG, "set_rp_variant: variant %d (0..3)", variant);
    c->rp_variant = variant;
    return 0;
}

int sharp_prof_enable(sharp_ctx *c, int on) {
    SHARP_TRY(use(c));
    prof_collect(c);
    c->prof_on = on != 0;
    for (sharp_ctx *s : c->subs) { prof_collect(s); s->prof_on = on != 0; }
    return 0;
}
int sharp_prof_reset(sharp_ctx *c) {
    SHARP_TRY(use(c));
    prof_collect(c);
    for (int i = 0; i < KID_COUNT; i++) { c->prof_ms[i] = 0; c->prof_n[i] = 0; }
    for (sharp_ctx *s : c->subs) {
        prof_collect(s);
        for (int i = 0; i < KID_COUNT; i++) { s->prof_ms[i] = 0; s->prof_n[i] = 0; }
    }
    return 0;
}
int sharp_prof_kernels(void) { return KID_COUNT; }
const char *sharp_prof_name(int kid) { return (kid >= 0 && kid < KID_COUNT) ? g_kernel_names[kid] : ""; }
int sharp_prof_get(sharp_ctx *c, int kid, double *ms, int64_t *launches) {
    SHARP_TRY(use(c));
    if (kid < 0 || kid >= KID_COUNT) return set_error(SHARP_E_ARG, "prof_get: bad kernel id %d", kid);
    prof_collect(c);
    double t = c->prof_ms[kid];
    int64_t n = c->prof_n[kid];
    for (sharp_ctx *s : c->subs) { /* the sub-contexts of group runs */
        prof_collect(s);
        t += s->prof_ms[kid];
        n += s->prof_n[kid];
    }
    if (ms) *ms = t;
    if (launches) *launches = n;
    return 0;
}

// ---- ranM upload: dgCMatrix slots of K matrices -> gene-major ternary entries -------------------------
// Device buffers of the projection matrices come from a small per-process cache: SHARP_unlimited uploads a fresh set of
// matrices on every call and frees it at the end, and cudaMalloc / cudaFree are device-wide synchronisations whose cost
// is unbounded on a busy 180 GB device (hundreds of ms were measured between two steps).  A released buffer is kept
// (up to 64 of them, a few MB in all) and handed to the next request of at most its size and at least half of it.
// fn(t) for t = 0..nt-1 on nt host threads (the caller runs t = 0)
extern "C++" {
template <class F>
static void host_parallel(int nt, F fn) {
    std::vector<std::thread> th;
    for (int t = 1; t < nt; t++) th.emplace_back([&fn, t]() { fn(t); });
    fn(0);
    for (auto &x : th) x.join();
}
}

namespace {
struct RmCacheEnt { int device; size_t cap; void *ptr; };
std::mutex g_rm_mu;
std::vector<RmCacheEnt> g_rm_cache;
std::vector<RmCacheEnt> g_rm_live;
cudaError_t rm_alloc(int device, void **out, size_t bytes) {
    bytes = std::max<size_t>(bytes, 256);
    std::lock_guard<std::mutex> lk(g_rm_mu);
    for (size_t i = 0; i < g_rm_cache.size(); i++) {
        RmCacheEnt e = g_rm_cache[i];
        if (e.device == device && e.cap >= bytes && e.cap <= 2 * bytes) {
            g_rm_cache.erase(g_rm_cache.begin() + i);
            g_rm_live.push_back(e);
            *out = e.ptr;
            return cudaSuccess;
        }
    }
    const size_t cap = (bytes + bytes / 8 + 4095) & ~(size_t)4095;
    cudaError_t err = cudaMalloc(out, cap);
    if (err == cudaSuccess) g_rm_live.push_back({device, cap, *out});
    return err;
}
void rm_release(void *ptr) {
    if (!ptr) return;
    std::lock_guard<std::mutex> lk(g_rm_mu);
    for (size_t i = 0; i < g_rm_live.size(); i++)
        if (g_rm_live[i].ptr == ptr) {
            RmCacheEnt e = g_rm_live[i];
            g_rm_live.erase(g_rm_live.begin() + i);
            if (g_rm_cache.size() < 64) g_rm_cache.push_back(e);
            else cudaFree(ptr);
            return;
        }
    cudaFree(ptr);
}
}  // namespace

// The K ranM matrices (dgCMatrix slots, concatenated) in the layouts of the projection kernels: gene-major CSR (fp64
// variant, overflow genes of the records), the padded vectors of the r1 fixed-point kernel (dense input) and the
// fixed-size records of rp_project_v3.cu.  Built by a few host threads (members, then gene ranges) into ONE pinned
// blob and copied with ONE cudaMemcpyAsync on a stream of its own, so the copy does not queue behind a look-ahead
// upload of expression data that is already in flight.
int sharp_rm_upload(sharp_ctx *c, int m, int p, int K, const int32_t *colptr, const int32_t *rowidx, const double *x,
                    const int64_t *nnz_off, sharp_rm_dev **out) {
    SHARP_TRY(use(c));
    Trace tr(c);
    if (!out || m <= 0 || p <= 0 || K <= 0 || !colptr || !nnz_off) return set_error(SHARP_E_ARG, "rm_upload: bad arguments");
    const int64_t nnz = nnz_off[K];
    if (nnz > 0 && (!rowidx || !x)) return set_error(SHARP_E_ARG, "rm_upload: missing slots");
    if ((int64_t)K * p > 0x7fffffffLL) return set_error(SHARP_E_LIMIT, "K*p too large");
    for (int k = 0; k < K; k++)
        if (colptr[(size_t)k * (p + 1) + p] != nnz_off[k + 1] - nnz_off[k])
            return set_error(SHARP_E_ARG, "rm_upload: colptr/nnz_off mismatch for matrix %d", k);
    const bool e16 = (int64_t)K * p <= 32768;
    const int NT = std::max(1, std::min(K, 8));
    // ---- per member: entries per gene, magnitude check, widest column ----
    std::vector<uint32_t> cntk((size_t)K * m, 0);
    std::vector<double> magk((size_t)K, 0.0), mag2k((size_t)K, 0.0);
    std::vector<int> badk((size_t)K, 0), colk((size_t)K, 0);
    host_parallel(NT, [&](int t) {
        for (int k = t; k < K; k += NT) {
            uint32_t *cnt = cntk.data() + (size_t)k * m;
            double mag = 0.0, other = 0.0;
            for (int64_t q = nnz_off[k]; q < nnz_off[k + 1]; q++) {
                if (rowidx[q] < 0 || rowidx[q] >= m) { badk[k] = 1; break; }
                const double a = std::fabs(x[q]);
                if (a == 0.0) continue;
                cnt[rowidx[q]]++;
                if (mag == 0.0) mag = a;
                else if (a != mag) other = a;
            }
            magk[k] = mag; mag2k[k] = other;
            const int32_t *cp = colptr + (size_t)k * (p + 1);
            int w = 0;
            for (int j = 0; j < p; j++) w = std::max(w, cp[j + 1] - cp[j]);
            colk[k] = w;
        }
    });
    double mag = 0.0;
    int max_col_nnz = 0;
    for (int k = 0; k < K; k++) {
        if (badk[k]) return set_error(SHARP_E_ARG, "rm_upload: row index out of range");
        double other = mag2k[k];
        if (magk[k] != 0.0) {
            if (mag == 0.0) mag = magk[k];
            else if (magk[k] != mag) other = magk[k];
        }
        if (other != 0.0)
            return set_error(SHARP_E_LIMIT, "rm_upload: the projection matrices are not ternary (|x| = %g and %g); only ranM()-style matrices are supported", mag, other);
        max_col_nnz = std::max(max_col_nnz, colk[k]);
    }
    // rowptr, and cntk[k][i] becomes the position of member k's first entry of gene i (members in order inside a gene)
    std::vector<uint32_t> rowptr((size_t)m + 1, 0);
    {
        uint32_t run = 0;
        for (int i = 0; i < m; i++) {
            for (int k = 0; k < K; k++) {
                uint32_t &cv = cntk[(size_t)k * m + i];
                const uint32_t t = cv;
                cv = run;
                run += t;
            }
            rowptr[i + 1] = run;
        }
    }
    const int64_t nz = rowptr[m];
    std::vector<uint32_t> ent((size_t)nz + 8, 0);
    host_parallel(NT, [&](int t) {
        for (int k = t; k < K; k += NT) {
            uint32_t *fill = cntk.data() + (size_t)k * m;
            const int32_t *cp = colptr + (size_t)k * (p + 1);
            for (int j = 0; j < p; j++)
                for (int32_t q = cp[j]; q < cp[j + 1]; q++) {
                    const int64_t g = nnz_off[k] + q;
                    if (x[g] == 0.0) continue;
                    const uint32_t col = (uint32_t)(k * p + j);
                    ent[fill[rowidx[g]]++] = e16 ? (col | (x[g] < 0 ? 0x8000u : 0u)) : (col | (x[g] < 0 ? 0x80000000u : 0u));
                }
        }
    });
    tr.mark("rm_csr");
    sharp_rm_dev *r = new sharp_rm_dev();
    r->device = c->device;
    r->m = m; r->p = p; r->K = K; r->mag = mag; r->nnz = nz;
    r->max_col_nnz = max_col_nnz;
    r->tile_genes = RP_TILE_GENES_HOST;
    r->ntiles = (m + r->tile_genes - 1) / r->tile_genes;
    for (int t = 0; t < r->ntiles; t++) {
        int g0 = t * r->tile_genes, g1 = std::min(m, g0 + r->tile_genes);
        r->max_tile_entries = std::max<int>(r->max_tile_entries, (int)(rowptr[g1] - rowptr[g0]));
    }
    // ---- layout of the blob ----
    const bool want_padded = (int64_t)K * p <= 32700;
    std::vector<uint32_t> vecptr;
    int rv = 0;
    if (want_padded) {
        const int kpr = (K * p + 31) & ~31;
        r->kpd = kpr + 32;
        vecptr.assign((size_t)m + 1, 0);
        double mean = 0, sq = 0;
        for (int i = 0; i < m; i++) {
            const uint32_t cn = rowptr[i + 1] - rowptr[i];
            vecptr[i + 1] = vecptr[i] + (cn + 7) / 8;
            mean += cn;
            sq += (double)cn * cn;
        }
        mean /= m;
        const double sd = std::sqrt(std::max(0.0, sq / m - mean * mean));
        r->vec_per_gene = std::min(6, std::max(1, (int)std::ceil((mean + 2.5 * sd) / 8.0)));
        if (r->kpd <= 8191 && max_col_nnz <= 255 && e16) {
            /* records of the record-gather kernel: the smallest record that fewer than 0.2 % of the genes overflow */
            for (rv = 2; rv <= 16; rv *= 2) {
                int over = 0;
                for (int i = 0; i < m; i++) over += (int)(rowptr[i + 1] - rowptr[i]) > 7 * rv;
                if (over <= m / 500) break;
            }
            if (rv > 16) rv = 0;
        }
    }
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t o_rowptr = 0;
    const size_t o_ent = al(o_rowptr + (size_t)(m + 1) * 4);
    const size_t ent_bytes = e16 ? ((size_t)nz + 16) * 2 : ((size_t)nz + 8) * 4;
    const size_t o_vecptr = al(o_ent + ent_bytes);
    const size_t o_entvec = al(o_vecptr + (want_padded ? (size_t)(m + 1) * 4 : 0));
    const size_t entvec_bytes = want_padded ? ((size_t)vecptr[m] * 8 + 8) * 2 : 0;
    const size_t o_rec = al(o_entvec + entvec_bytes);
    const size_t rec_bytes = (size_t)m * rv * 16;
    const size_t total = al(o_rec + rec_bytes);
    if (c->rm_stage_cap < total) {
        if (c->rm_stage) cudaFreeHost(c->rm_stage);
        c->rm_stage = nullptr;
        c->rm_stage_cap = 0;
        const size_t cap = total + total / 4;
        cudaError_t eh = cudaHostAlloc((void **)&c->rm_stage, cap, cudaHostAllocDefault);
        if (eh != cudaSuccess) { delete r; return set_error(SHARP_E_CUDA, "rm_upload: pinned staging of %zu bytes: %s", cap, cudaGetErrorString(eh)); }
        c->rm_stage_cap = cap;
    }
    unsigned char *H = c->rm_stage;
    memcpy(H + o_rowptr, rowptr.data(), (size_t)(m + 1) * 4);
    if (want_padded) memcpy(H + o_vecptr, vecptr.data(), (size_t)(m + 1) * 4);
    const int kpr = r->kpd - 32;
    const uint32_t kpd4 = (uint32_t)r->kpd * 4u;
    const int GT = 4; /* gene ranges */
    host_parallel(GT, [&](int t) {
        const int g0 = (int)((int64_t)m * t / GT), g1 = (int)((int64_t)m * (t + 1) / GT);
        if (e16) {
            uint16_t *E = reinterpret_cast<uint16_t *>(H + o_ent);
            for (uint32_t q = rowptr[g0]; q < rowptr[g1]; q++) E[q] = (uint16_t)ent[q];
            if (t == GT - 1) for (int64_t q = nz; q < nz + 16; q++) E[q] = 0;
        } else {
            uint32_t *E = reinterpret_cast<uint32_t *>(H + o_ent);
            for (uint32_t q = rowptr[g0]; q < rowptr[g1]; q++) E[q] = ent[q];
            if (t == GT - 1) for (int64_t q = nz; q < nz + 8; q++) E[q] = 0;
        }
        if (want_padded) {
            uint16_t *Pd = reinterpret_cast<uint16_t *>(H + o_entvec);
            for (int i = g0; i < g1; i++) {
                const uint32_t cn = rowptr[i + 1] - rowptr[i];
                uint16_t *dst = Pd + (size_t)vecptr[i] * 8;
                for (uint32_t q = 0; q < cn; q++) dst[q] = (uint16_t)ent[rowptr[i] + q];
                for (uint32_t q = cn; q < ((cn + 7) & ~7u); q++) dst[q] = (uint16_t)(kpr + (i & 31));
            }
            if (t == GT - 1) for (int q = 0; q < 8; q++) Pd[(size_t)vecptr[m] * 8 + q] = (uint16_t)kpr;
        }
        if (rv) {
            uint16_t *R0 = reinterpret_cast<uint16_t *>(H + o_rec);
            for (int i = g0; i < g1; i++) {
                const uint32_t cn = rowptr[i + 1] - rowptr[i];
                uint16_t *R = R0 + (size_t)i * rv * 8;
                /* unused slots point at one of the 32 dummy words behind the + counts (never read back), spread over the
                   banks: a kernel may scatter all seven slots of a vector without looking at the count */
                for (int v = 0; v < rv; v++) {
                    R[v * 8] = 0;
                    for (int sl = 1; sl < 8; sl++) R[v * 8 + sl] = (uint16_t)((kpr + ((i * rv + v + 5 * sl) & 31)) << 2);
                }
                if ((int)cn > 7 * rv) { R[0] = 0xffffu; continue; }
                for (uint32_t q = 0; q < cn; q++) {
                    const uint32_t en = ent[rowptr[i] + q];
                    const int v = (int)(q % rv), slot = 1 + (int)(q / rv);
                    R[v * 8 + slot] = (uint16_t)(((en & 0x7fffu) << 2) + ((en & 0x8000u) ? kpd4 : 0u)); /* offset into [+ | -] */
                    R[v * 8]++;
                }
            }
        }
    });
    tr.mark("rm_layouts");
    unsigned char *dev = nullptr;
    cudaError_t e1 = rm_alloc(c->device, (void **)&dev, total);
    if (e1 == cudaSuccess && !c->rm_stream) e1 = cudaStreamCreateWithFlags(&c->rm_stream, cudaStreamNonBlocking);
    /* copied by a kernel that reads the page-locked blob, not by the copy engine: the engine serves copies in submission
       order, and a look-ahead upload of expression data (a GB per part) may already be queued */
    if (e1 == cudaSuccess && h2d_by_kernel(c->rm_stream, dev, H, total) != 0) e1 = cudaErrorUnknown;
    if (e1 == cudaSuccess) e1 = cudaStreamSynchronize(c->rm_stream);
    tr.mark("rm_copy");
    if (e1 != cudaSuccess) {
        if (dev) rm_release(dev);
        delete r;
        return set_error(SHARP_E_CUDA, "rm_upload: %s", cudaGetErrorString(e1));
    }
    r->blob = dev;
    r->rowptr = reinterpret_cast<uint32_t *>(dev + o_rowptr);
    if (e16) r->ent16 = reinterpret_cast<uint16_t *>(dev + o_ent);
    else r->ent32 = reinterpret_cast<uint32_t *>(dev + o_ent);
    if (want_padded) {
        r->vecptr = reinterpret_cast<uint32_t *>(dev + o_vecptr);
        r->entvec = reinterpret_cast<uint4 *>(dev + o_entvec);
    }
    if (rv) {
        r->rec = reinterpret_cast<uint4 *>(dev + o_rec);
        r->rv = rv;
    }
    *out = r;
    return 0;
}

void sharp_rm_free(sharp_rm_dev *r) {
    if (!r) return;
    cudaSetDevice(r->device);
    rm_release(r->blob); /* every layout lives in the one blob */
    delete r;
}

int sharp_expr_upload(sharp_ctx *c, int m, int64_t n, const double *dense, const int64_t *colptr, const int32_t *rowidx,
                      const double *val, sharp_expr_dev **out) {
    SHARP_TRY(use(c));
    if (!out) return set_error(SHARP_E_ARG, "null output pointer");
    sharp_expr_dev *e = new sharp_expr_dev();
    int rc = upload_expr(c, m, n, dense, colptr, rowidx, val, e);
    if (rc) {
        free_expr(e);
        delete e;
        return rc;
    }
    *out = e;
    return 0;
}

void sharp_expr_free(sharp_expr_dev *e) {
    if (!e) return;
    cudaSetDevice(e->device);
    free_expr(e);
    delete e;
}

int sharp_rp_project(sharp_ctx *c, int m, int64_t n, const double *dense, const int64_t *colptr, const int32_t *rowidx,
                     const double *val, const int64_t *cells, int64_t ncell, int normalize, const double *colsum,
                     double norm_mul, int logkind, int round_digits, const sharp_rm_dev *rm, double *out) {
    SHARP_TRY(use(c));
    if (!rm || !out) return set_error(SHARP_E_ARG, "rp_project: null argument");
    sharp_expr_dev e;
    int rc = upload_expr(c, m, n, dense, colptr, rowidx, val, &e, true);
    auto body = [&]() -> int {
        if (rc) return rc;
        int64_t *cells_dev = nullptr;
        if (cells) {
            for (int64_t i = 0; i < ncell; i++)
                if (cells[i] < 0 || cells[i] >= n) return set_error(SHARP_E_ARG, "rp_project: cell index out of range");
            SHARP_TRY(c->ws[WS_SRC].reserve((size_t)ncell * 8));
            cells_dev = c->ws[WS_SRC].as<int64_t>();
            SHARP_TRY(h2d(c, cells_dev, cells, (size_t)ncell * 8));
        } else ncell = n;
        double *cs = nullptr;
        if (normalize) {
            SHARP_TRY(c->ws[WS_COLSUM].reserve((size_t)n * 8));
            cs = c->ws[WS_COLSUM].as<double>();
            if (normalize == 1) {
                if (!colsum) return set_error(SHARP_E_ARG, "normalize = 1 needs colsum");
                SHARP_TRY(h2d(c, cs, colsum, (size_t)n * 8));
            }
        }
        size_t ob = (size_t)rm->K * ncell * rm->p * 8;
        SHARP_TRY(c->ws[WS_PROJ].reserve(ob));
        SHARP_TRY(launch_rp_project(c, e, cells_dev, ncell, cs, normalize != 2, normalize, norm_mul, logkind, round_digits, *rm, c->ws[WS_PROJ].as<double>()));
        SHARP_TRY(d2h(c, out, c->ws[WS_PROJ].ptr, ob));
        return sync(c);
    };
    rc = body();
    cudaStreamSynchronize(c->stream);
    free_expr(&e);
    return rc;
}

int sharp_corrdist(sharp_ctx *c, int n, int p, const double *mat, double *dist) {
    SHARP_TRY(use(c));
    if (n < 1 || p < 1 || !mat || !dist) return set_error(SHARP_E_ARG, "corrdist: bad arguments");
    const int ld = ld_of(n), ldu = ldu_of(p);
    SHARP_TRY(c->ws[WS_TMP0].reserve((size_t)n * p * 8));
    SHARP_TRY(c->ws[WS_U].reserve((size_t)n * ldu * 8));
    SHARP_TRY(c->ws[WS_D].reserve((size_t)n * ld * 8));
    SHARP_TRY(c->ws[WS_TMP1].reserve((size_t)n * n * 8));
    SHARP_TRY(h2d(c, c->ws[WS_TMP0].ptr, mat, (size_t)n * p * 8));
    SHARP_TRY(launch_unit_rows(c, c->ws[WS_TMP0].as<double>(), n, p, ldu, c->ws[WS_U].as<double>()));
    SHARP_TRY(c->reserve_pinned(4096));
    SHARP_TRY(c->ws[WS_DESC].reserve(4096));
    Bump hb(c->pinned, 4096), db(c->ws[WS_DESC].ptr, 4096);
    GemmProb *gp = hb.take<GemmProb>(1);
    int *tp = hb.take<int>(2);
    gp->U = c->ws[WS_U].as<double>(); gp->n = n; gp->ld = ld; gp->D = c->ws[WS_D].as<double>(); gp->Dw = nullptr;
    tp[0] = 0; tp[1] = corrdist_tiles(n);
    GemmProb *gpd = db.take<GemmProb>(1);
    int *tpd = db.take<int>(2);
    SHARP_TRY(h2d_staged(c, c->ws[WS_DESC].ptr, c->pinned, hb.off));
    SHARP_TRY(launch_corrdist_batched(c, gpd, tpd, 1, tp[1], ldu));
    prof_begin(c, KID_MISC);
    pad_copy_kernel<<<grid1d((size_t)n * n, 256), 256, 0, c->stream>>>(n, n, c->ws[WS_D].as<double>(), ld, c->ws[WS_TMP1].as<double>(), n);
    prof_end(c);
    SHARP_TRY(d2h(c, dist, c->ws[WS_TMP1].ptr, (size_t)n * n * 8));
    return sync(c);
}

int sharp_hclust(sharp_ctx *c, int n, const double *dist, int method, int32_t *ia, int32_t *ib, double *height) {
    SHARP_TRY(use(c));
    if (n < 2) return set_error(SHARP_E_RSTOP, "hclust: must have n >= 2 objects to cluster");
    if (!dist || !ia || !ib || !height) return set_error(SHARP_E_ARG, "hclust: null argument");
    const int ld = ld_of(n);
    SHARP_TRY(c->ws[WS_TMP1].reserve((size_t)n * n * 8));
    SHARP_TRY(c->ws[WS_DW].reserve((size_t)n * ld * 8));
    SHARP_TRY(h2d(c, c->ws[WS_TMP1].ptr, dist, (size_t)n * n * 8));
    prof_begin(c, KID_MISC);
    pad_copy_kernel<<<grid1d((size_t)n * n, 256), 256, 0, c->stream>>>(n, n, c->ws[WS_TMP1].as<double>(), n, c->ws[WS_DW].as<double>(), ld);
    prof_end(c);
    size_t ibytes = bump_size({(size_t)n * 4, (size_t)n * 4}), dbytes = bump_size({(size_t)n * 8});
    SHARP_TRY(c->ws[WS_HC_INT].reserve(ibytes));
    SHARP_TRY(c->ws[WS_HC_DBL].reserve(dbytes));
    Bump bi(c->ws[WS_HC_INT].ptr, ibytes), bd(c->ws[WS_HC_DBL].ptr, dbytes);
    int *iad = bi.take<int>(n), *ibd = bi.take<int>(n);
    double *cr = bd.take<double>(n);
    SHARP_TRY(c->reserve_pinned(4096));
    SHARP_TRY(c->ws[WS_DESC].reserve(4096));
    HcProb *hp = reinterpret_cast<HcProb *>(c->pinned);
    memset(hp, 0, sizeof(HcProb));
    hp->n = n; hp->ld = ld; hp->D = nullptr; hp->Dw = c->ws[WS_DW].as<double>(); hp->ia = iad; hp->ib = ibd; hp->crit = cr;
    SHARP_TRY(h2d(c, c->ws[WS_DESC].ptr, hp, sizeof(HcProb)));
    SHARP_TRY(launch_hclust(c, c->ws[WS_DESC].as<HcProb>(), 1, n, method));
    int st = 0;
    SHARP_TRY(d2h(c, ia, iad, (size_t)(n - 1) * 4));
    SHARP_TRY(d2h(c, ib, ibd, (size_t)(n - 1) * 4));
    SHARP_TRY(d2h(c, height, cr, (size_t)(n - 1) * 8));
    SHARP_TRY(d2h(c, &st, &c->ws[WS_DESC].as<HcProb>()->status, 4));
    SHARP_TRY(sync(c));
    if (st != 0) return rstop(st, "hclust");
    return 0;
}

static int fetch_opt(sharp_ctx *c, int nrow, const OptResult &R, int32_t *f, int32_t *v, int *nlev, double *msil,
                     double *chind, double *height, int *optn, double *maxsil, int *oind, const char *where) {
    int meta[8];
    SHARP_TRY(d2h(c, meta, R.meta, sizeof meta));
    SHARP_TRY(sync(c));
    if (meta[3] != 0) return rstop(meta[3], where);
    const int L = meta[0];
    if (f) SHARP_TRY(d2h(c, f, R.f, (size_t)nrow * 4));
    if (v && R.v) SHARP_TRY(d2h(c, v, R.v, (size_t)nrow * L * 4));
    if (msil) SHARP_TRY(d2h(c, msil, R.msil, (size_t)L * 8));
    if (chind) SHARP_TRY(d2h(c, chind, R.chind, (size_t)L * 8));
    if (height) SHARP_TRY(d2h(c, height, R.crit, (size_t)(nrow - 1) * 8));
    if (maxsil) SHARP_TRY(d2h(c, maxsil, R.maxsil, 8));
    SHARP_TRY(sync(c));
    if (nlev) *nlev = L;
    if (optn) *optn = meta[1];
    if (oind) *oind = meta[2];
    return 0;
}

int sharp_opt_hclust(sharp_ctx *c, int nrow, int ncol, const double *mat, int symmetric, int exact,
                     const sharp_hc_params *prm, int32_t *f, int32_t *v, int *nlev, double *msil, double *chind,
                     double *height, int *optn, double *maxsil, int *oind) {
    SHARP_TRY(use(c));
    if (!mat || !prm || nrow < 1 || ncol < 1) return set_error(SHARP_E_ARG, "opt_hclust: bad arguments");
    SHARP_TRY(c->ws[WS_TMP0].reserve((size_t)nrow * ncol * 8));
    SHARP_TRY(h2d(c, c->ws[WS_TMP0].ptr, mat, (size_t)nrow * ncol * 8));
    OptResult R;
    SHARP_TRY(opt_hclust_dev(c, nrow, ncol, c->ws[WS_TMP0].as<double>(), symmetric ? 1 : 0, exact, *prm, v != nullptr, &R));
    return fetch_opt(c, nrow, R, f, v, nlev, msil, chind, height, optn, maxsil, oind, "get_opt_hclust");
}

int sharp_getrowcolor(sharp_ctx *c, int n, int p, const double *emat, const sharp_hc_params *prm, int32_t *color, double *maxsil) {
    SHARP_TRY(use(c));
    if (!emat || !prm || !color) return set_error(SHARP_E_ARG, "getrowcolor: bad arguments");
    SHARP_TRY(c->ws[WS_TMP0].reserve((size_t)n * p * 8));
    SHARP_TRY(h2d(c, c->ws[WS_TMP0].ptr, emat, (size_t)n * p * 8));
    OptResult R;
    SHARP_TRY(opt_hclust_dev(c, n, p, c->ws[WS_TMP0].as<double>(), 0, 0, *prm, false, &R));
    prof_begin(c, KID_MISC);
    colour_wrap_kernel<<<grid1d(n, 256), 256, 0, c->stream>>>(R.f, n);
    prof_end(c);
    return fetch_opt(c, n, R, color, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, maxsil, nullptr, "getrowColor");
}

int sharp_wmetac(sharp_ctx *c, int N, int C, const int32_t *labels, const sharp_hc_params *prm, int32_t *finalc,
                 int *ncluster, double *x0, int max_x0_cols, double *w1) {
    SHARP_TRY(use(c));
    if (N < 1 || C < 1 || !labels || !prm || !finalc) return set_error(SHARP_E_ARG, "wmetac: bad arguments");
    // the kernels want codes 1..L in first-appearance order per column: remap on the host (R glue passes
    // match(x, unique(x)) already; this makes any integer coding valid)
    std::vector<int32_t> lab((size_t)N * C);
    int max_label = 1;
    for (int k = 0; k < C; k++) {
        std::vector<std::pair<int32_t, int>> seen;
        for (int i = 0; i < N; i++) {
            int32_t l = labels[(size_t)k * N + i];
            int code = 0;
            for (auto &s : seen)
                if (s.first == l) { code = s.second; break; }
            if (!code) { code = (int)seen.size() + 1; seen.push_back({l, code}); }
            lab[(size_t)k * N + i] = code;
        }
        max_label = std::max<int>(max_label, (int)seen.size());
    }
    SHARP_TRY(c->ws[WS_ENRP].reserve((size_t)N * C * 4));
    SHARP_TRY(h2d(c, c->ws[WS_ENRP].ptr, lab.data(), (size_t)N * C * 4));
    int64_t start[2] = {0, N};
    SHARP_TRY(c->ws[WS_START].reserve(16));
    SHARP_TRY(h2d(c, c->ws[WS_START].ptr, start, 16));
    SHARP_TRY(sync(c));
    WmBuffers W;
    SHARP_TRY(wmetac_dev(c, c->ws[WS_ENRP].as<int32_t>(), N, C, 1, start, c->ws[WS_START].as<int64_t>(), max_label, *prm, &W));
    int st = 0, nu = 0;
    SHARP_TRY(d2h(c, &st, W.A.status, 4));
    SHARP_TRY(d2h(c, &nu, W.A.ucount, 4));
    SHARP_TRY(sync(c));
    if (st != 0) return rstop(st, "wMetaC");
    SHARP_TRY(d2h(c, finalc, W.A.finalc, (size_t)N * 4));
    if (w1) SHARP_TRY(d2h(c, w1, W.A.w1, (size_t)N * 8));
    if (ncluster) *ncluster = nu;
    if (x0) {
        if (nu > max_x0_cols) return set_error(SHARP_E_NOMEM, "wmetac: x0 needs %d columns", nu);
        SHARP_TRY(c->ws[WS_X0].reserve((size_t)N * nu * 8));
        SHARP_CUDA(cudaMemsetAsync(c->ws[WS_X0].ptr, 0, (size_t)N * nu * 8, c->stream));
        SHARP_TRY(launch_wmetac_x0(c, W.A, W.outs_dev, 1, N, nullptr, nullptr, nullptr, c->ws[WS_X0].as<double>(), nu));
        SHARP_TRY(d2h(c, x0, c->ws[WS_X0].ptr, (size_t)N * nu * 8));
    }
    return sync(c);
}

// host helper: first-appearance codes + member lists
static void codes_and_lists(int64_t n, const int32_t *labels, std::vector<int> &code, std::vector<int> &corder,
                            std::vector<int> &coff, int *nc) {
    code.resize(n);
    std::vector<std::pair<int32_t, int>> sorted;
    int next = 0;
    for (int64_t i = 0; i < n; i++) {
        int32_t l = labels[i];
        auto it = std::lower_bound(sorted.begin(), sorted.end(), std::make_pair(l, -1));
        if (it == sorted.end() || it->first != l) it = sorted.insert(it, {l, next++});
        code[i] = it->second;
    }
    *nc = next;
    coff.assign(next + 1, 0);
    for (int64_t i = 0; i < n; i++) coff[code[i] + 1]++;
    for (int q = 0; q < next; q++) coff[q + 1] += coff[q];
    corder.resize(n);
    std::vector<int> fill(coff.begin(), coff.end() - 1);
    for (int64_t i = 0; i < n; i++) corder[fill[code[i]]++] = (int)i;
}

int sharp_smetac(sharp_ctx *c, int64_t ncells, int p, const int32_t *labels, const double *se1,
                 const sharp_hc_params *prm, int32_t *finalcolor, int32_t *tf, int *nc) {
    SHARP_TRY(use(c));
    if (ncells < 1 || p < 2 || !labels || !se1 || !prm || !finalcolor) return set_error(SHARP_E_ARG, "smetac: bad arguments");
    if (ncells > 2000000000LL) return set_error(SHARP_E_LIMIT, "smetac: too many cells");
    std::vector<int> code, corder, coff;
    int nC = 0;
    codes_and_lists(ncells, labels, code, corder, coff, &nC);
    if (nC < 2) return rstop(SM_E_FEWCLUSTERS, "sMetaC");
    size_t ib = bump_size({(size_t)ncells * 4, (size_t)ncells * 4, (size_t)(nC + 1) * 4, 64, 64, (size_t)ncells * 4});
    SHARP_TRY(c->ws[WS_TMP0].reserve(ib));
    Bump b(c->ws[WS_TMP0].ptr, ib);
    int *code_d = b.take<int>(ncells), *corder_d = b.take<int>(ncells), *coff_d = b.take<int>(nC + 1);
    int *nc_d = b.take<int>(1), *st_d = b.take<int>(1);
    int *out_d = b.take<int>(ncells);
    SHARP_TRY(c->ws[WS_E1].reserve((size_t)ncells * p * 8));
    SHARP_TRY(h2d(c, c->ws[WS_E1].ptr, se1, (size_t)ncells * p * 8));
    SHARP_TRY(h2d(c, code_d, code.data(), (size_t)ncells * 4));
    SHARP_TRY(h2d(c, corder_d, corder.data(), (size_t)ncells * 4));
    SHARP_TRY(h2d(c, coff_d, coff.data(), (size_t)(nC + 1) * 4));
    int zero = 0;
    SHARP_TRY(h2d(c, nc_d, &nC, 4));
    SHARP_TRY(h2d(c, st_d, &zero, 4));
    SHARP_TRY(sync(c));
    SmBuffers B;
    SHARP_TRY(smetac_dev(c, nC, p, nc_d, st_d, c->ws[WS_E1].as<double>(), corder_d, coff_d, nullptr, ncells, *prm, &B));
    int st = 0;
    SHARP_TRY(d2h(c, &st, B.status, 4));
    SHARP_TRY(sync(c));
    if (st != 0) return rstop(st, "sMetaC");
    SHARP_TRY(launch_sm_relabel(c, ncells, code_d, B.tf, 0, nullptr, out_d));
    SHARP_TRY(d2h(c, finalcolor, out_d, (size_t)ncells * 4));
    if (tf) SHARP_TRY(d2h(c, tf, B.tf, (size_t)nC * 4));
    if (nc) *nc = nC;
    return sync(c);
}

int sharp_smetac_centroids(sharp_ctx *c, int nC, int p, const double *cen, int64_t ncells_total,
                           const sharp_hc_params *prm, int32_t *tf) {
    SHARP_TRY(use(c));
    if (nC < 1 || p < 2 || !cen || !prm || !tf) return set_error(SHARP_E_ARG, "smetac_centroids: bad arguments");
    if (nC < 2) return rstop(SM_E_FEWCLUSTERS, "sMetaC");
    SHARP_TRY(c->ws[WS_CEN].reserve((size_t)nC * p * 8 + 64));
    SHARP_TRY(c->ws[WS_TMP0].reserve(256));
    int *nc_d = c->ws[WS_TMP0].as<int>(), *st_d = nc_d + 16;
    int zero = 0;
    SHARP_TRY(h2d(c, c->ws[WS_CEN].ptr, cen, (size_t)nC * p * 8));
    SHARP_TRY(h2d(c, nc_d, &nC, 4));
    SHARP_TRY(h2d(c, st_d, &zero, 4));
    SHARP_TRY(sync(c));
    SmBuffers B;
    SHARP_TRY(smetac_dev(c, nC, p, nc_d, st_d, nullptr, nullptr, nullptr, c->ws[WS_CEN].as<double>(), ncells_total, *prm, &B));
    int st = 0;
    SHARP_TRY(d2h(c, &st, B.status, 4));
    SHARP_TRY(sync(c));
    if (st != 0) return rstop(st, "sMetaC");
    SHARP_TRY(d2h(c, tf, B.tf, (size_t)nC * 4));
    return sync(c);
}

int sharp_run_dev(sharp_ctx *c, const sharp_expr_dev *e, const double *colsum, const sharp_rm_dev *rm,
                  const int64_t *reind, const sharp_run_params *prm, int32_t *labels, double *vie, double *x0,
                  int *x0_cols, int max_x0_cols) {
    SHARP_TRY(use(c));
    if (!e || !rm || !prm || !labels) return set_error(SHARP_E_ARG, "run: null argument");
    if (e->m != rm->m) return set_error(SHARP_E_ARG, "run: expression has %d genes but ranM has %d rows", e->m, rm->m);
    return run_core(c, *e, colsum, *rm, reind, *prm, labels, vie, x0, x0_cols, max_x0_cols);
}

int sharp_run(sharp_ctx *c, int m, int64_t n, const double *dense, const int64_t *colptr, const int32_t *rowidx,
              const double *val, const double *colsum, const sharp_rm_dev *rm, const int64_t *reind,
              const sharp_run_params *prm, int32_t *labels, double *vie, double *x0, int *x0_cols, int max_x0_cols) {
    SHARP_TRY(use(c));
    if (!rm || !prm || !labels) return set_error(SHARP_E_ARG, "run: null argument");
    sharp_expr_dev e;
    int rc;
    const bool shuffled = prm->large && reind && n < 100000;
    if (prm->large && prm->shard && c->comm && c->comm_world > 1 && !shuffled && n >= 2) {
        /* a sharded run on un-shuffled host data: this rank's blocks are a contiguous range of columns -- upload only those */
        std::vector<int64_t> start;
        make_blocks(n, 1, prm->partition_ncells, start);
        const int T = (int)start.size() - 1;
        if (T < c->comm_world) return set_error(SHARP_E_ARG, "sharded run: %d blocks cannot be dealt over %d ranks", T, c->comm_world);
        int t0, t1;
        shard_range(T, c->comm_rank, c->comm_world, prm->shard_rotate, &t0, &t1);
        const int64_t c0 = start[t0], nc = start[t1] - start[t0];
        if (dense) rc = upload_expr(c, m, nc, dense + (size_t)c0 * m, nullptr, nullptr, nullptr, &e, true);
        else {
            if (!colptr) return set_error(SHARP_E_ARG, "expression matrix: neither dense nor complete CSC slots given");
            rc = c->reserve_pinned((size_t)(nc + 1) * 8);
            if (!rc) {
                int64_t *cp = reinterpret_cast<int64_t *>(c->pinned);
                const int64_t base = colptr[c0];
                for (int64_t i = 0; i <= nc; i++) cp[i] = colptr[c0 + i] - base;
                rc = upload_expr(c, m, nc, nullptr, cp, rowidx + base, val + base, &e, true);
            }
        }
        e.col0 = c0;
        e.n_total = n;
    } else rc = upload_expr(c, m, n, dense, colptr, rowidx, val, &e, true);
    if (!rc) rc = sharp_run_dev(c, &e, colsum, rm, reind, prm, labels, vie, x0, x0_cols, max_x0_cols);
    cudaStreamSynchronize(c->stream);
    free_expr(&e);
    return rc;
}

}  // extern "C"

// group boundaries of a run over `nparts` parts (shared by sharp_run_parts and sharp_parts_prefetch)
static std::vector<int> plan_groups(int nparts, bool host_first, int &group, int &lanes) {
    /* parts in host memory, many of them: three groups of two in flight hide the uploads best (B200, r2, 26 parts of
       50 000 cells: 709 ms per run against 750 for two groups of four; device-resident parts: 668 vs 665) */
    const bool many_host = host_first && nparts > 8;
    if (lanes <= 0) lanes = (nparts > 8 && group <= 0) ? 3 : 2; /* device-resident, 26 parts: 655 ms with 3 lanes of 4 vs 664 with 2 */
    /* default group size: 4 parts share the block-clustering launches when there are many parts; with few parts (a rank
       of a multi-GPU job) smaller groups keep both lanes busy -- a group running alone leaves the device half idle
       during its latency-bound stages */
    /* measured on B200 (r2, 50 000-cell parts): 4 resident parts 107 ms as 2+2 vs 111 ms as 1+1+1+1; 4 parts in host
       memory 141 ms as 1+2+1 vs 123 ms as 1+1+1+1 (a part's upload hides behind ONE part's clustering) */
    if (group <= 0) group = nparts <= 8 ? (host_first ? 1 : 2) : (many_host ? 2 : 4);
    group = std::min(group, nparts);
    // the parts are split as evenly as the group size allows, the smaller groups first and last (the first group's start
    // and the last group's tail are the stretches of the run nothing else overlaps with).  With host data the first
    // group is half a group: its upload is the only one that is not hidden behind another group.
    std::vector<int> gstart{0};
    const bool host_data = host_first && group >= 2 && nparts > group;
    if (host_data) gstart.push_back(group / 2);
    const int rest = nparts - gstart.back();
    const int ng = (rest + group - 1) / group;
    const int base = rest / ng, extra = rest % ng;
    const int first_big = host_data ? 0 : (ng - extra) / 2; /* groups [first_big, first_big + extra) get base + 1 parts */
    for (int g = 0; g < ng; g++) gstart.push_back(gstart.back() + base + ((g >= first_big && g < first_big + extra) ? 1 : 0));
    return gstart;
}

// groups of a parts list whose block-sharded parts (sharp_part.sharded, all at the END of the list) carry only 1 / world of
// a part's work each: the whole parts are grouped as usual and the sharded ones ride with the last group, so that their
// few (member, block) problems share that group's launches instead of paying a wave of their own
static int plan_groups_parts(const sharp_part *parts, int nparts, bool host_first, int &group, int &lanes, std::vector<int> &gstart,
                             int *gmax) {
    int ns = 0;
    for (int i = 0; i < nparts; i++) {
        if (parts[i].sharded) ns++;
        else if (ns) return set_error(SHARP_E_ARG, "run_parts: block-sharded parts must come after the rank's own parts");
    }
    const int nwhole = nparts - ns;
    gstart = plan_groups(nwhole > 0 ? nwhole : nparts, host_first, group, lanes);
    if (nwhole > 0) gstart.back() = nparts;
    *gmax = 1;
    for (size_t g = 0; g + 1 < gstart.size(); g++) *gmax = std::max(*gmax, gstart[g + 1] - gstart[g]);
    return 0;
}

static int ensure_subs(sharp_ctx *c, size_t need) {
    while (c->subs.size() < need) {
        sharp_ctx *s = nullptr;
        SHARP_TRY(make_child(c, &s));
        c->subs.push_back(s);
    }
    return 0;
}

extern "C" {

int sharp_run_parts(sharp_ctx *c, int m, int nparts, sharp_part *parts, const sharp_rm_dev *rm,
                    const sharp_run_params *prm, int small_thre, int cen_cap, int group, int lanes) {
    SHARP_TRY(use(c));
    if (!parts || !rm || !prm || nparts < 1) return set_error(SHARP_E_ARG, "run_parts: bad arguments");
    if (!prm->large) return set_error(SHARP_E_ARG, "run_parts: only the SHARP_large path (every part >= base.ncells cells)");
    for (int i = 0; i < nparts; i++) {
        if (!parts[i].pred) return set_error(SHARP_E_ARG, "run_parts: part %d has no output buffer", i);
        if (!parts[i].dev && !parts[i].dense && !parts[i].colptr) return set_error(SHARP_E_ARG, "run_parts: part %d has no data", i);
        if (parts[i].dev && parts[i].dev->n != parts[i].n) return set_error(SHARP_E_ARG, "run_parts: part %d: n does not match the device matrix", i);
    }
    std::vector<int> gstart;
    int gmax = 1;
    SHARP_TRY(plan_groups_parts(parts, nparts, !parts[0].dev, group, lanes, gstart, &gmax));
    const int ngroups = (int)gstart.size() - 1;
    lanes = std::min(lanes, ngroups);
    const int stride = gmax + 1;   /* sub-contexts per lane: one per part of the largest group + the group's block context */
    const size_t need = (size_t)lanes * stride;
    SHARP_TRY(ensure_subs(c, need));
    for (sharp_ctx *s : c->subs) {
        if (c->serial) s->pf_part = -1;
        s->prof_on = c->prof_on;
        s->rp_variant = c->rp_variant;
        s->comm = c->comm; s->comm_rank = c->comm_rank; s->comm_world = c->comm_world;
        s->block_budget_gb = std::max(1, c->block_budget_gb / lanes);
    }
    // serial mode (profiling): all sub-contexts enqueue on the context's own stream, so no two kernels overlap and
    // the per-kernel event brackets measure each launch alone; same launches and grids as the concurrent run
    struct StreamSwap {
        sharp_ctx *c;
        std::vector<cudaStream_t> saved;
        ~StreamSwap() {
            for (size_t i = 0; i < saved.size(); i++) c->subs[i]->stream = saved[i];
        }
    } swap{c, {}};
    if (c->serial)
        for (size_t i = 0; i < need; i++) {
            swap.saved.push_back(c->subs[i]->stream);
            c->subs[i]->stream = c->stream;
        }
    // children start after whatever is queued on the context's own stream (and the timer's start event)
    SHARP_CUDA(cudaEventRecord(c->ev_fork, c->stream));
    for (size_t i = 0; i < need; i++) SHARP_CUDA(cudaStreamWaitEvent(c->subs[i]->stream, c->ev_fork, 0));
    cudaStream_t up = nullptr;
    if (!c->serial) {
        if (!c->up_stream) SHARP_CUDA(cudaStreamCreateWithFlags(&c->up_stream, cudaStreamNonBlocking));
        up = c->up_stream;
        SHARP_CUDA(cudaStreamWaitEvent(up, c->ev_fork, 0));
    }
    std::vector<GroupRun> G(ngroups);
    int rc = 0;
    static const bool trace = getenv("SHARP_B200_TRACE") != nullptr;
    const auto t_start = std::chrono::steady_clock::now();
    for (int gi = 0; gi < ngroups && !rc; gi++) {
        GroupRun &g = G[gi];
        const int lane = gi % lanes;
        for (int i = gstart[gi]; i < gstart[gi + 1]; i++) g.idx.push_back(i);
        g.blocks = c->subs[(size_t)lane * stride + gmax];
        g.up = up;
        for (size_t j = 0; j < g.idx.size(); j++) g.subs.push_back(c->subs[(size_t)lane * stride + j]);
        const auto t_a = std::chrono::steady_clock::now();
        rc = group_issue(g, parts, m, *rm, *prm);
        if (!rc && !c->serial && lanes >= 2 && gi >= 1 && gi - 1 + lanes < ngroups) {
            /* look-ahead for the group that gets the PREVIOUS group's lane next (its uploads queue up behind this group's,
               so the copy engine sees the parts in order) */
            const int ng = gi - 1 + lanes, nlane = (gi - 1) % lanes;
            std::vector<int> nidx;
            std::vector<sharp_ctx *> lsubs;
            for (int i = gstart[ng]; i < gstart[ng + 1]; i++) {
                nidx.push_back(i);
                lsubs.push_back(c->subs[(size_t)nlane * stride + (i - gstart[ng])]);
            }
            rc = group_prefetch(nidx, lsubs, parts, m, up);
        }
        const auto t_b = std::chrono::steady_clock::now();
        /* the group that used this lane's contexts `lanes` groups ago has been completed below before its contexts are
           reused here; complete the oldest outstanding group while the newer ones keep the device busy */
        if (!rc && gi + 1 >= lanes) rc = group_complete(G[gi + 1 - lanes], parts, small_thre, cen_cap);
        if (trace) {
            const auto t_c = std::chrono::steady_clock::now();
            fprintf(stderr, "[sharp trace run_parts] group %d lane %d: issue %.2f ms, complete(prev) %.2f ms, t=%.2f ms\n", gi, lane,
                    std::chrono::duration<double, std::milli>(t_b - t_a).count(),
                    std::chrono::duration<double, std::milli>(t_c - t_b).count(),
                    std::chrono::duration<double, std::milli>(t_c - t_start).count());
        }
    }
    for (int gi = std::max(0, ngroups - lanes + 1); gi < ngroups && !rc; gi++) rc = group_complete(G[gi], parts, small_thre, cen_cap);
    // join: the context's stream (and its timer) sees the end of all the work
    for (size_t i = 0; i < need; i++) {
        cudaEventRecord(c->subs[i]->ev_ready, c->subs[i]->stream);
        cudaStreamWaitEvent(c->stream, c->subs[i]->ev_ready, 0);
    }
    if (rc) {
        std::string msg = g_err; /* keep the first error across the drain */
        if (up) cudaStreamSynchronize(up);
        for (size_t i = 0; i < need; i++) cudaStreamSynchronize(c->subs[i]->stream);
        cudaGetLastError();
        snprintf(g_err, sizeof g_err, "%s", msg.c_str());
        return rc;
    }
    return sync(c);
}

int sharp_plan_groups(int nparts, int host_data, int group, int lanes, int *gstart, int cap, int *ngroups, int *group_used,
                      int *lanes_used) {
    if (nparts < 1 || !gstart || !ngroups) return set_error(SHARP_E_ARG, "plan_groups: bad arguments");
    std::vector<int> g = plan_groups(nparts, host_data != 0, group, lanes);
    if ((int)g.size() > cap) return set_error(SHARP_E_NOMEM, "plan_groups: %d boundaries but room for %d", (int)g.size(), cap);
    for (size_t i = 0; i < g.size(); i++) gstart[i] = g[i];
    *ngroups = (int)g.size() - 1;
    if (group_used) *group_used = group;
    if (lanes_used) *lanes_used = std::min(lanes, *ngroups);
    return 0;
}

int sharp_parts_prefetch(sharp_ctx *c, int m, int nparts, sharp_part *parts, int group, int lanes) {
    SHARP_TRY(use(c));
    if (!parts || nparts < 1) return set_error(SHARP_E_ARG, "parts_prefetch: bad arguments");
    if (c->serial || parts[0].dev) return 0;
    std::vector<int> gstart;
    int gmax = 1;
    SHARP_TRY(plan_groups_parts(parts, nparts, true, group, lanes, gstart, &gmax));
    const int ngroups = (int)gstart.size() - 1;
    lanes = std::min(lanes, ngroups);
    SHARP_TRY(ensure_subs(c, (size_t)lanes * (gmax + 1)));
    if (!c->up_stream) SHARP_CUDA(cudaStreamCreateWithFlags(&c->up_stream, cudaStreamNonBlocking));
    /* behind whatever is queued on the context's stream (the caller's timer start, the previous call's end) */
    SHARP_CUDA(cudaEventRecord(c->ev_fork, c->stream));
    SHARP_CUDA(cudaStreamWaitEvent(c->up_stream, c->ev_fork, 0));
    std::vector<int> idx;
    std::vector<sharp_ctx *> subs;
    /* only the FIRST part of group 0: this call is made before the projection matrices exist (their draw takes about as
       long as one part's copy), and the copy of those matrices must not queue behind a whole group of parts -- the
       first kernels need both.  The other parts of the group are copied by the run itself, behind the matrices. */
    for (int i = gstart[0]; i < gstart[1] && i < gstart[0] + 1; i++) {
        idx.push_back(i);
        subs.push_back(c->subs[(size_t)(i - gstart[0])]); /* lane 0 */
    }
    return group_prefetch(idx, subs, parts, m, c->up_stream);
}

int sharp_last_member(sharp_ctx *c, int k, int64_t n, int32_t *rowcolor, double *inde) {
    SHARP_TRY(use(c));
    if (n != c->last_n || !c->ws[WS_ENRP].ptr || !c->ws[WS_PROJ].ptr)
        return set_error(SHARP_E_ARG, "last_member: no matching run on this context");
    if (k < 0 || k >= c->last_K) return set_error(SHARP_E_ARG, "last_member: member %d out of range (0..%d)", k, c->last_K - 1);
    if (rowcolor) SHARP_TRY(d2h(c, rowcolor, c->ws[WS_ENRP].as<int32_t>() + (size_t)k * n, (size_t)n * 4));
    if (inde) SHARP_TRY(d2h(c, inde, c->ws[WS_PROJ].as<double>() + (size_t)k * n * c->last_p, (size_t)n * c->last_p * 8));
    return sync(c);
}

int sharp_last_vie(sharp_ctx *c, int64_t n, int p, double *vie) {
    SHARP_TRY(use(c));
    if (!vie) return set_error(SHARP_E_ARG, "last_vie: null output");
    if (n != c->last_n || p != c->last_p || !c->ws[WS_VIEU].ptr)
        return set_error(SHARP_E_ARG, "last_vie: no matching run on this context");
    SHARP_TRY(d2h(c, vie, c->ws[WS_VIEU].ptr, (size_t)n * p * 8));
    return sync(c);
}

int sharp_centroids(sharp_ctx *c, int64_t n, const int32_t *labels, int nclust, double *cen, int64_t *counts) {
    SHARP_TRY(use(c));
    if (!labels || !cen || nclust < 1) return set_error(SHARP_E_ARG, "centroids: bad arguments");
    if (n != c->last_n || !c->ws[WS_VIEU].ptr) return set_error(SHARP_E_ARG, "centroids: no matching run on this context");
    const int p = c->last_p;
    std::vector<int> coff(nclust + 1, 0), corder(n);
    for (int64_t i = 0; i < n; i++) {
        if (labels[i] < 1 || labels[i] > nclust) return set_error(SHARP_E_ARG, "centroids: label out of range");
        coff[labels[i]]++;
    }
    for (int q = 0; q < nclust; q++) coff[q + 1] += coff[q];
    {
        std::vector<int> fill(coff.begin(), coff.end() - 1);
        for (int64_t i = 0; i < n; i++) corder[fill[labels[i] - 1]++] = (int)i;
    }
    size_t ib = bump_size({(size_t)n * 4, (size_t)(nclust + 1) * 4, 64, (size_t)nclust * 8});
    SHARP_TRY(c->ws[WS_TMP0].reserve(ib));
    Bump b(c->ws[WS_TMP0].ptr, ib);
    int *corder_d = b.take<int>(n), *coff_d = b.take<int>(nclust + 1), *nc_d = b.take<int>(1);
    int64_t *cnt_d = b.take<int64_t>(nclust);
    SHARP_TRY(c->ws[WS_CEN].reserve((size_t)nclust * p * 8));
    SHARP_TRY(h2d(c, corder_d, corder.data(), (size_t)n * 4));
    SHARP_TRY(h2d(c, coff_d, coff.data(), (size_t)(nclust + 1) * 4));
    SHARP_TRY(h2d(c, nc_d, &nclust, 4));
    SHARP_TRY(sync(c));
    SHARP_TRY(launch_sm_centroids(c, c->ws[WS_VIEU].as<double>(), p, corder_d, coff_d, nc_d, nclust, c->ws[WS_CEN].as<double>(), cnt_d));
    SHARP_TRY(d2h(c, cen, c->ws[WS_CEN].ptr, (size_t)nclust * p * 8));
    if (counts) SHARP_TRY(d2h(c, counts, cnt_d, (size_t)nclust * 8));
    return sync(c);
}

}  /* extern "C" */
