// internal.cuh -- shared declarations of libsharpb200 (not part of the public ABI).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/sharp_b200.h"

namespace sharp {

// ---- errors ------------------------------------------------------------------------------------
int set_error(int code, const char *fmt, ...);
#define SHARP_CUDA(expr)                                                                                 \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            return ::sharp::set_error(SHARP_E_CUDA, "%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, \
                                      cudaGetErrorString(_e));                                           \
    } while (0)
#define SHARP_TRY(expr)           \
    do {                          \
        int _rc = (expr);         \
        if (_rc != 0) return _rc; \
    } while (0)

// Dynamic shared memory opt-in of every kernel: always the sm_100 maximum (227 KB), never the per-launch size -- host
// threads driving different contexts set it concurrently, and a smaller value set by one thread between another
// thread's attribute call and its launch would make that launch fail.
constexpr int SHARP_SMEM_OPTIN = 222 * 1024;  /* 227 KB minus room for the kernels' small static arrays */

// The opt-in is set ONCE per kernel and device, not per launch: cudaFuncSetAttribute is not a cheap host-side setter --
// under SHARP_B200_TRACE it was seen blocking the enqueueing thread for hundreds of ms while other streams were
// running the same kernel, which stalled the whole group pipeline.  `flags` is a function-local static array.
#define SHARP_SMEM_OPTIN_ONCE(kernel, device)                                                                     \
    do {                                                                                                          \
        static std::atomic<unsigned char> _done[64];                                                              \
        const int _d = (device) & 63;                                                                             \
        if (!_done[_d].load(std::memory_order_acquire)) {                                                         \
            SHARP_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SHARP_SMEM_OPTIN)); \
            _done[_d].store(1, std::memory_order_release);                                                        \
        }                                                                                                         \
    } while (0)

// ---- device buffer that grows and never shrinks (workspace slot) -------------------------------------
struct DevBuf {
    void *ptr = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes);
    void release();
    template <class T> T *as() const { return reinterpret_cast<T *>(ptr); }
};

// A problem for the hierarchical clustering kernels (lives in DEVICE memory; n may be written by a kernel).
struct HcProb {
    int n;            // number of objects (0 = skip)
    int ld;           // leading dimension of the distance matrices
    double *D;        // full symmetric n x ld, read-only after it is built (silhouettes)
    double *Dw;       // working copy, destroyed by the agglomeration
    int *ia;          // [n-1] 1-based representative kept      (hclust.f IA)
    int *ib;          // [n-1] 1-based representative absorbed  (hclust.f IB)
    double *crit;     // [n-1] heights
    const double *Y;  // data for the CH index: feature problems: unit rows (n x p); similarity problems: S
    int p;            // columns of Y
    int ldy;          // leading dimension of Y
    int status;       // 0 ok, else a SWEEP_E_* / WM_E_* code
    int fallback;     // set by hclust_rnn_kernel: exact ties, redo with the exact kernel
    double *E;        // second working buffer of hclust_rnn_kernel (ecap doubles), else null
    long long ecap;
};

// results of the cluster-number sweep for one problem (device pointers)
struct SweepOut {
    int *f;          // [n] chosen labels (cutree ids, first-appearance numbering, 1-based)
    int *v;          // optional [n x nlev] row-major
    double *msil;    // [maxlev]
    double *chind;   // [maxlev]
    int *meta;       // [8]: 0 nlev, 1 optn, 2 oind, 3 status, 4 kmin, 5 kmax
    double *maxsil;  // [1]
};

struct HcParamsDev {
    int hmethod, n_cluster, min_n, max_n;
    double sil_thre, height_ntimes;
};

inline HcParamsDev to_dev(const sharp_hc_params &p) {
    HcParamsDev d;
    d.hmethod = p.hmethod ? p.hmethod : SHARP_WARD_D;
    d.n_cluster = p.n_cluster;
    d.min_n = p.min_n > 0 ? p.min_n : 2;
    d.max_n = p.max_n > 0 ? p.max_n : 40;
    d.sil_thre = p.sil_thre;
    d.height_ntimes = p.height_ntimes;
    return d;
}

enum { SWEEP_E_HCLUST_NA = 12, SWEEP_E_KRANGE = 15, SWEEP_E_NAN = 17, SWEEP_E_CHNAN = 18, SWEEP_E_UNSORTED = 19,
       WM_E_ONECLUSTER = 20, SWEEP_E_FEWPOINTS = 21, SWEEP_E_OIND = 22, WM_E_FEWCLUSTERS = 23, SM_E_FEWCLUSTERS = 24,
       WM_E_TOOMANY = 30 };

const char *status_message(int st);

}  // namespace sharp

namespace sharp {
// kernel classes of the per-kernel device-time profile (sharp_prof_*; bench.py's roofline object)
enum KernelId { KID_RP_PROJECT = 0, KID_COLSUM, KID_UNIT_ROWS, KID_CORRDIST, KID_HCLUST, KID_HCLUST_SMALL,
                KID_SWEEP_NESTED, KID_SWEEP_EXACT, KID_WM_WEIGHTS, KID_WM_SIMILARITY, KID_WMETAC, KID_SM_CENTROIDS,
                KID_SMETAC, KID_ENE, KID_MISC, KID_H2D, KID_COMM, KID_COUNT };
struct ProfPending { int kid; cudaEvent_t a, b; };
}  // namespace sharp

struct sharp_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int64_t launches = 0;
    sharp::DevBuf ws[48];  // grow-only workspaces (slots named in pipeline.cu)
    // Pinned staging for descriptor uploads / small downloads: a ring arena of two halves.  reserve_pinned(bytes) points
    // `pinned` at a FRESH chunk, so a chunk handed to cudaMemcpyAsync is never rewritten while the copy is in flight and
    // no stream synchronisation is needed between the stages of a run; re-entering a half waits for the event recorded
    // when it was left (every staged copy of this context is issued on its own stream).
    void *pinned = nullptr;
    size_t pinned_cap = 0;      // size of the chunk `pinned` points at
    unsigned char *arena = nullptr;
    size_t arena_cap = 0, arena_off = 0;
    cudaEvent_t ev_half[2] = {nullptr, nullptr};
    bool half_used[2] = {false, false};
    int half_cur = 0;
    // sub-contexts of a group run (sharp_run_parts): own workspaces and streams, created on first use
    std::vector<sharp_ctx *> subs;
    sharp_ctx *parent = nullptr;
    cudaEvent_t ev_ready = nullptr, ev_blocks = nullptr, ev_done = nullptr, ev_fork = nullptr;
    // upload look-ahead of a group run: the expression buffers of a sub-context are free again as soon as its part's
    // front stage (projection) is done, so the NEXT part this sub-context will get is copied in behind that event, on a
    // stream of its own, while the current part is still being clustered
    cudaStream_t up_stream = nullptr;
    cudaEvent_t ev_up = nullptr;
    int pf_part = -1;             // index of the prefetched part (-1: none)
    const void *pf_src = nullptr; // host buffer it was copied from (a prefetch is only used for the same buffer)
    int32_t *h_labels = nullptr;  // pinned label mirror of a group run
    size_t h_labels_cap = 0;
    cudaStream_t rm_stream = nullptr;   // copies of the projection matrices (sharp_rm_upload)
    unsigned char *rm_stage = nullptr;  // their pinned staging blob
    size_t rm_stage_cap = 0;
    unsigned char *h_gather = nullptr;  // pinned staging of the host-side column gather of a sharded, shuffled part
    size_t h_gather_cap = 0;
    int block_budget_gb = 48;     // cap of the distance-matrix workspace (D + Dw) of one context
    size_t ws_budget = 0, ws_budget_seen = 0;  // last budget derived from cudaMemGetInfo and the workspace state it was derived for
    int64_t last_n = 0;     // state of the last run (for sharp_centroids)
    int last_p = 0;
    int last_K = 0;
    bool serial = false;    // sharp_run_parts: every sub-context enqueues on THIS context's stream (isolated kernel timings)
    // multi-GPU: an NCCL communicator (ncclComm_t) attached by sharp_comm_init; sub-contexts share their parent's
    void *comm = nullptr;
    int comm_rank = 0, comm_world = 1;
    int rp_variant = 0;     // projection kernel: 0 record-gather fixed point (CSC; default), 1 fp64 read-modify-write, 2 the r1
                            // fixed-point kernel (also what dense input takes), 3 record-gather with TMA-staged cell segments
    int reserve_pinned(size_t bytes);
    // per-kernel profile: CUDA events around every launch on `stream` while prof_on (off by default)
    bool prof_on = false;
    std::vector<cudaEvent_t> prof_pool;
    std::vector<sharp::ProfPending> prof_pending;
    int prof_open = -1;
    double prof_ms[sharp::KID_COUNT] = {0};
    int64_t prof_n[sharp::KID_COUNT] = {0};
};

struct sharp_rm_dev {
    int device = 0;
    int m = 0, p = 0, K = 0;
    double mag = 0.0;            // common |x| of all entries
    int tile_genes = 0, ntiles = 0, max_tile_entries = 0;
    int64_t nnz = 0;
    uint32_t *rowptr = nullptr;  // [m+1] offsets into ent (gene-major)
    uint16_t *ent16 = nullptr;   // [nnz] (col | sign<<15) when K*p <= 32768
    uint32_t *ent32 = nullptr;   // [nnz] (col | sign<<31) otherwise
    // padded copy for the fixed-point kernel (K*p <= 32767 only): every gene's entries start on a 16-byte boundary
    // and are padded to a multiple of 8 with dummy columns kpd-32+(gene&31); vecptr[g] counts 8-entry vectors
    uint32_t *vecptr = nullptr;  // [m+1]
    uint4 *entvec = nullptr;     // [vecptr[m]]
    int kpd = 0;                 // K*p rounded up to 32, plus the 32 dummy columns the padding entries point at
    int max_col_nnz = 0;         // largest number of entries in one column of one member (bound on adds per output)
    int vec_per_gene = 0;        // vectors preloaded per gene by the kernel (covers ~99 % of the genes)
    // fixed-size records of the record-gather kernel (rp_project_v3.cu): rv vectors of {count, 7 entries} per gene, the
    // gene's entries dealt round-robin over them; entry = byte offset of the output in an accumulator array | sign
    uint4 *rec = nullptr;        // [m * rv]
    unsigned char *blob = nullptr;  // the one device allocation all of the above point into
    int rv = 0;                  // vectors per record (2, 4, 8 or 16); 0: no records (K*p too large)
};

struct sharp_expr_dev {
    int device = 0;
    int m = 0;
    int64_t n = 0;
    double *dense = nullptr;
    int64_t *colptr = nullptr;
    int32_t *rowidx = nullptr;
    double *val = nullptr;
    int64_t nnz = 0;
    bool owned = true;
    // a column slice of a larger matrix (sharded runs on un-shuffled host data upload only the rank's columns):
    // column 0 of this matrix is column col0 of the whole one, which has n_total columns (0: this IS the whole matrix)
    int64_t col0 = 0, n_total = 0;
    // compact = true: column i of this matrix is the i-th cell of the rank's share of a SHUFFLED sharded matrix (gathered on
    // the host in shuffled order): the projection reads it with the identity map
    bool compact = false;
};

namespace sharp {

constexpr int RP_TILE_GENES_HOST = 512;

// every launch site brackets its <<<>>> with these (prof_end also counts the launch)
void prof_begin(sharp_ctx *c, int kid);
void prof_end(sharp_ctx *c);
void prof_collect(sharp_ctx *c);

// comm.cu: allgather of per-rank device segments (bytes[r] from rank r, concatenated in rank order at recv) on stream st
int comm_allgatherv_dev(sharp_ctx *c, const void *send, void *recv, const int64_t *bytes, cudaStream_t st);
void comm_destroy(sharp_ctx *c);

// ---- kernel launchers (each returns after enqueueing on ctx->stream) ---------------------------------
// rp_project.cu
int launch_colsum(sharp_ctx *c, const sharp_expr_dev &e, double *colsum);
// colsum_dev: [n] per source column when normalize != 0.  colsum_ready = false (normalize = 2): the sums have not been
// computed yet -- the launcher does it (fused into the record-gather kernel's pre-pass where that applies)
int launch_rp_project(sharp_ctx *c, const sharp_expr_dev &e, const int64_t *cells_dev, int64_t ncell,
                      double *colsum_dev, bool colsum_ready, int normalize, double norm_mul, int logkind, int round_digits,
                      const sharp_rm_dev &rm, double *out /* [K][ncell][p] */);
constexpr int WS_COMM_SLOT = 46;      // staging of the host-buffer collectives (comm.cu)
constexpr int WS_CELLINFO_SLOT = 47;  // workspace slot of the per-cell records of the record-gather kernel (last of ws[48])
// corrdist.cu
int launch_unit_rows(sharp_ctx *c, const double *X, int64_t rows, int p, int ldu, double *U);
struct GemmProb {
    const double *U;  // first unit row of the problem (row stride ldu)
    int n;
    int ld;
    double *D;
    double *Dw;       // may be null
};
int corrdist_tiles(int n);
int launch_corrdist_batched(sharp_ctx *c, const GemmProb *probs_dev, const int *tile_prefix_dev, int nprob,
                            int total_tiles, int ldu);
// ward.cu
int launch_hclust(sharp_ctx *c, HcProb *probs_dev, int nprob, int max_n, int method, int fast = 0);
// fast = 1 will take the round-parallel kernel: the problems then need HcProb::E, and the distance kernel need not
// write the working copy Dw (round 1 reads the pristine D)
bool hclust_fast_ok(int max_n, int method);
// sweep.cu
int h2d_by_kernel(cudaStream_t st, void *dst, const void *src, size_t bytes); /* pipeline.cu: copy engine bypass for small page-locked sources */
size_t sweep_nested_scratch_bytes(int max_n, int max_p, HcParamsDev prm);
int launch_sweep_nested(sharp_ctx *c, HcProb *probs_dev, SweepOut *outs_dev, int nprob, int max_n, int max_p,
                        HcParamsDev prm, double *scratch, size_t scratch_per_prob);
size_t sweep_exact_scratch_bytes(int max_n, int max_p);
int launch_sweep_exact(sharp_ctx *c, HcProb *probs_dev, SweepOut *outs_dev, int nprob, int max_n, int max_p,
                       const HcParamsDev *prm_dev, int max_levels, int kcap, double *scratch, size_t scratch_per_prob);
constexpr int NESTED_MAXK_HOST = 64;

}  // namespace sharp
