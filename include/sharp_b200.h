/*
 * sharp_b200.h -- C ABI of libsharpb200.so: the B200-native (sm_100a) implementation of SHARP's
 * ensemble random-projection clustering hot path.
 *
 * The reference (shibiaowan/SHARP v1.1.0) is a pure-R package with NO native code and NO FFI
 * (no src/, no useDynLib in NAMESPACE): its operator API is the set of exported R closures.  Each
 * entry point below is therefore what an R `.Call` glue for this path binds (INTEGRATION.md shows the
 * glue), and cites the R closure / lines whose body it replaces.  Everything is plain C: raw pointers,
 * sizes, int status.  No torch types, no C++ types.
 *
 * Conventions
 *  - All functions return 0 on success, a negative SHARP_E_* code otherwise; sharp_last_error() gives
 *    the message (per calling thread).  SHARP_E_RSTOP marks conditions where the reference itself
 *    would stop() (the message quotes the R error).
 *  - Host pointers unless a parameter is documented as "device".  Inputs are never written.
 *  - fp64 everywhere (R's numeric); int32 labels, 1-based like R.
 *  - Matrices: the expression matrix is genes x cells exactly as R stores it (dense column-major, or
 *    dgCMatrix slots i / p / x).  Cell-by-feature matrices (projections, viE, x0) are ROW-major
 *    ncells x ncol, i.e. the memory image of R's t(matrix).
 *  - There is NO CPU fallback: every call needs a CUDA device and fails with SHARP_E_CUDA without one.
 */
#ifndef SHARP_B200_H
#define SHARP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SHARP_B200_ABI_VERSION 4   /* 4: sharp_run_params.skip_smetac (SHARP_fpart); 3: + sharp_parts_prefetch, sharp_plan_groups, sharp_ctx_set_serial */

enum {
    SHARP_OK = 0,
    SHARP_E_ARG = -1,    /* bad argument */
    SHARP_E_CUDA = -2,   /* CUDA runtime failure / no device */
    SHARP_E_NOMEM = -3,  /* output buffer too small or device memory exhausted */
    SHARP_E_RSTOP = -4,  /* the reference would stop() here (message quotes it) */
    SHARP_E_LIMIT = -5   /* a documented size limit of this implementation */
};

/* stats::hclust method codes (index into hclust's METHODS vector = hclust.f iOpt) */
enum { SHARP_WARD_D = 1, SHARP_SINGLE = 2, SHARP_COMPLETE = 3, SHARP_AVERAGE = 4, SHARP_MCQUITTY = 5,
       SHARP_MEDIAN = 6, SHARP_CENTROID = 7, SHARP_WARD_D2 = 8 };

/* Arguments shared by get_opt_hclust / getrowColor / wMetaC / sMetaC   (R/get_opt_hclust.R:33-62) */
typedef struct {
    int hmethod;          /* SHARP_WARD_D ... ; 0 = default "ward.D" */
    int n_cluster;        /* N.cluster: 0 = NULL (choose automatically) */
    int min_n;            /* minN.cluster (0 = default 2) */
    int max_n;            /* maxN.cluster (0 = default 40) */
    double sil_thre;      /* sil.thre */
    double height_ntimes; /* height.Ntimes */
} sharp_hc_params;

typedef struct sharp_ctx sharp_ctx;       /* one CUDA device + stream + workspace */
typedef struct sharp_rm_dev sharp_rm_dev; /* K ranM matrices prepared on the device */

/* ---- library / device -------------------------------------------------------------------------
 * Replaces: registerDoParallel(n.cores) / detectCores()  (R/SHARP.R:162-167): `n.cores` is accepted by the
 * R wrappers and ignored; the unit of parallelism is the GPU. */
int sharp_abi_version(void);
const char *sharp_last_error(void);
int sharp_device_count(void);
int sharp_device_info(int device, char *name, int name_len, int *sm_count, int *cc_major, int *cc_minor,
                      size_t *total_mem);
int sharp_ctx_create(int device, sharp_ctx **ctx);
void sharp_ctx_destroy(sharp_ctx *ctx);
/* CUDA stream of the context as a void* (cudaStream_t), for callers that time with CUDA events. */
void *sharp_ctx_stream(sharp_ctx *ctx);
int sharp_ctx_sync(sharp_ctx *ctx);
/* device-time helpers on the context's stream (CUDA events): start/stop a region, elapsed in ms. */
int sharp_timer_start(sharp_ctx *ctx);
int sharp_timer_stop_ms(sharp_ctx *ctx, double *ms);
/* number of kernels this library launched on the context since creation (bench.py's gpu_launches). */
int64_t sharp_ctx_launch_count(sharp_ctx *ctx);

/* Projection kernel variant (all compute the same exact 64-bit fixed-point sums except 1; the switch exists so that tests
 * and ncu can compare them):
 *   0 (default) record-gather kernel for CSC input: one fixed-size record per gene fetched by one coalesced request, signed
 *     16-bit count fields for the non-zeros equal to 1..4, per-cell scale from a streaming pre-pass (rp_project_v3.cu);
 *     dense input and matrices it does not cover take variant 2;
 *   1 the fp64 read-modify-write kernel that adds every output's terms in ascending gene order (bit-identical to a
 *     sequential sparse product);
 *   2 the round-1 fixed-point kernel (pointer array + padded entry lists, 8-bit count fields);
 *   3 variant 0 with the cell's (rowidx, val) segments staged in shared memory one cell ahead by cp.async.bulk (TMA).
 * All meet the 1e-5 contract by seven orders of magnitude or more. */
int sharp_ctx_set_rp_variant(sharp_ctx *ctx, int variant);

/* per-kernel device-time profile (off by default): while enabled, every launch of this library on the context's
 * stream is bracketed by CUDA events and accumulated per kernel class.  bench.py's roofline object is computed from
 * these (live, over the timed region). */
int sharp_prof_enable(sharp_ctx *ctx, int on);
int sharp_prof_reset(sharp_ctx *ctx);
int sharp_prof_kernels(void);              /* number of kernel classes */
const char *sharp_prof_name(int kid);      /* name of class kid */
int sharp_prof_get(sharp_ctx *ctx, int kid, double *ms, int64_t *launches);

/* ---- ranM matrices ----------------------------------------------------------------------------
 * Replaces nothing: ranM()/ranM2() stay in R (R/ranM.R:11-33, R/ranM2.R:11-35) and their dgCMatrix slots
 * are handed in.  K matrices, each m x p: colptr K x (p+1) (slot `p`), rowidx (slot `i`, ascending inside a
 * column) and x (slot `x`) concatenated, nnz_off[K+1] offsets of each matrix in rowidx/x.
 * All non-zero entries must have the same magnitude (they do: +-sqrt(sqrt(m))); otherwise SHARP_E_LIMIT. */
int sharp_rm_upload(sharp_ctx *ctx, int m, int p, int K, const int32_t *colptr, const int32_t *rowidx,
                    const double *x, const int64_t *nnz_off, sharp_rm_dev **out);
void sharp_rm_free(sharp_rm_dev *rm);

/* R-compatible generators for hosts without R (sharp_b200/csrc/rrng.cpp; the reference's own calls are cited there).
 * sharp_r_ranm = ranM2(m, p, seed) for an integer seed (R/ranM2.R:11-35 == R/ranM.R:11-33 with m = nrow(scdata)) as
 * dgCMatrix slots; cap = room in rowidx / x, *nnz always set, SHARP_E_NOMEM if cap is too small.
 * sharp_r_sample_perm = set.seed(seed); sample(n) (R/SHARP.R:495-498), 1-based.  Both are pure host code. */
int sharp_r_ranm(int m, int p, int64_t seed, int32_t *colptr, int32_t *rowidx, double *x, int64_t cap, int64_t *nnz);
int sharp_r_sample_perm(int64_t seed, int64_t n, int64_t *out);

/* ---- a2/a3/a4: random projection with the normalisation fused into the load ---------------------
 * Replaces: RPmat()'s `1/sqrt(p) * t(x) %*% scdata` (R/RPmat.R:32), the inline projection of SHARP_large /
 * SHARP_fpart / testlog (R/SHARP.R:567-585, 899-905; R/SHARP_unlimited2.R:388-410), log2(x+1)
 * (R/SHARP.R:344,570) and the CPM normalisation t(t(x)/colSums(x))*1e6 (R/SHARP.R:113).
 *
 * E (m genes x n cells): dense column-major (`dense`) or CSC (`colptr` int64[n+1], `rowidx`, `val`).
 * cells: optional 0-based source column of every output row (E[, tind]); NULL = all n columns in order.
 * normalize: 0 = none; 1 = value / colsum[c] * norm_mul with colsum given (per SOURCE column, length n);
 *            2 = same with the column sums computed on the device.
 * logkind: 0 none, 2 = log2(x+1), 10 = log10(x+1).  round_digits: <0 none, else round(., digits).
 * out: K matrices, each row-major ncell x p, concatenated (out[k][cell][j]). */
int sharp_rp_project(sharp_ctx *ctx, int m, int64_t n, const double *dense, const int64_t *colptr,
                     const int32_t *rowidx, const double *val, const int64_t *cells, int64_t ncell, int normalize,
                     const double *colsum, double norm_mul, int logkind, int round_digits, const sharp_rm_dev *rm,
                     double *out);

/* ---- a6: distance   as.dist(1 - cor(t(scale-rows(mat))))   (R/get_opt_hclust.R:71-72) -------------
 * mat row-major n x p -> dist full symmetric n x n (diag 0).  Stage-level entry used by the parity tests. */
int sharp_corrdist(sharp_ctx *ctx, int n, int p, const double *mat, double *dist);

/* ---- a7: stats::hclust(d, method)   (R/get_opt_hclust.R:77) ---------------------------------------
 * dist: full symmetric n x n (what as.matrix(d) holds).  ia/ib[n-1]: 1-based representatives (min index of
 * each cluster) merged at every step, in hclust.f order; height[n-1]. */
int sharp_hclust(sharp_ctx *ctx, int n, const double *dist, int method, int32_t *ia, int32_t *ib, double *height);

/* ---- a6-a10: get_opt_hclust   (R/get_opt_hclust.R:33-244) -----------------------------------------
 * mat row-major nrow x ncol.  symmetric: 1 = similarity matrix (d = 1 - mat), 0 = feature matrix.
 * (isSymmetric(mat) is evaluated by the caller: it is a property of the R object.)
 * exact: 1 = every level swept with the reference's summation order (bit-compatible, O(levels n^2));
 *        0 = the nested one-pass sweep used for the 2000-cell blocks (same ties, last-bit differences).
 * Outputs (any may be NULL): f[nrow]; v row-major nrow x *nlev (room for nrow x (max_n-min_n+1));
 * msil/chind[*nlev]; height[nrow-1]; *optn = optN.cluster; *maxsil; *oind (1-based column of v). */
int sharp_opt_hclust(sharp_ctx *ctx, int nrow, int ncol, const double *mat, int symmetric, int exact,
                     const sharp_hc_params *prm, int32_t *f, int32_t *v, int *nlev, double *msil, double *chind,
                     double *height, int *optn, double *maxsil, int *oind);

/* ---- a5: getrowColor   (R/getrowColor.R:17-121) ---------------------------------------------------
 * emat row-major n x p.  color[n]: index (1..40) into the reference's colour vector (cluster ids in order of
 * first appearance, wrapping modulo 40 exactly like R/getrowColor.R:59-68). */
int sharp_getrowcolor(sharp_ctx *ctx, int n, int p, const double *emat, const sharp_hc_params *prm, int32_t *color,
                      double *maxsil);

/* ---- a11-a14: wMetaC   (R/wMetaC.R:15-226 incl. getA :242-283, getss/getnewk :299-320) -------------
 * labels: column-major N x C integer codes (the R glue passes match(nC[,c], unique(nC[,c]))).
 * finalc[N]: the meta-cluster id R returns as a string in finalC; *ncluster = length(unique(finalC));
 * x0 (may be NULL): row-major N x *ncluster, columns in unique(finalC) order, room for N x max_x0_cols;
 * w1 (may be NULL): the adjusted point weights (R/wMetaC.R:44). */
int sharp_wmetac(sharp_ctx *ctx, int N, int C, const int32_t *labels, const sharp_hc_params *prm, int32_t *finalc,
                 int *ncluster, double *x0, int max_x0_cols, double *w1);

/* ---- a15: sMetaC   (R/sMetaC.R:17-210) ------------------------------------------------------------
 * labels[ncells]: integer codes of rerowColor (any codes; unique() order = first appearance);
 * se1 row-major ncells x p.  finalcolor[ncells]; tf[nC] (may be NULL); *nc = nC. */
int sharp_smetac(sharp_ctx *ctx, int64_t ncells, int p, const int32_t *labels, const double *se1,
                 const sharp_hc_params *prm, int32_t *finalcolor, int32_t *tf, int *nc);

/* ---- a16: SHARP_small / SHARP_large compute, fused on the device -----------------------------------
 * Replaces the bodies of SHARP_small (R/SHARP.R:343-416) and SHARP_large (R/SHARP.R:502-783) (and SHARP_fpart,
 * R/SHARP_unlimited2.R:316-520, through logkind/round_digits): projection of every (member, block), per-block
 * getrowColor, gather of enrp / enE, per-block wMetaC, sMetaC across blocks, un-shuffle.  What stays in R:
 * argument defaults, prep, seeds / ranM / sample(), the small-cluster merge and the final relabel
 * (R/SHARP.R:816-843), the result list. */
typedef struct {
    int large;            /* 0 = SHARP_small (one block, no sMetaC), 1 = SHARP_large */
    int logflag;          /* `flag` */
    int logkind;          /* 2 or 10 (0 = 2) */
    int round_digits;     /* <0 none */
    int partition_ncells; /* ng */
    int n_cluster;        /* N.cluster      (0 = NULL) */
    int enp_n_cluster;    /* enpN.cluster   (0 = NULL) */
    int ind_n_cluster;    /* indN.cluster   (0 = NULL) */
    sharp_hc_params hc;   /* hmethod, minN, maxN, sil.thre, height.Ntimes (n_cluster ignored) */
    int normalize;        /* as in sharp_rp_project */
    double norm_mul;
    int skip_smetac;      /* large = 1 only: 1 = stop after the per-block wMetaC like SHARP_fpart (R/SHARP_unlimited2.R:477-531):
                             labels[] = an injective code of fColor = "<finalC>en<t>" (1 + position of the (block, meta-cluster)
                             pair in block order), un-shuffled; no x0.  0 = SHARP_large (sMetaC across the blocks) */
    int block_max_n;      /* > 0: maxN.cluster of the per-(member, block) clusterings only (SHARP_fpart assigns 40 inside its
                             worker, R/SHARP_unlimited2.R:421; its wMetaC keeps the caller's value, :481); 0 = hc.max_n */
    int shard;            /* large = 1 with a communicator on the context (sharp_comm_init): 1 = the cell blocks of THIS matrix
                             are dealt over the ranks (the K x T nested loop of R/SHARP.R:554-618 and the per-block wMetaC,
                             :692-709, run on the rank that owns the block); block-level results and the rows of enE / K are
                             allgathered and every rank finishes with the whole matrix's labels / viE.  Every rank makes the
                             same call; no x0.  Un-shuffled host data (n >= 1e5): only the rank's columns are uploaded */
    int shard_rotate;     /* rotates which ranks get the larger shares when the blocks do not divide evenly */
} sharp_run_params;

/* reind: 1-based permutation from `set.seed(50); sample(ncells)` or NULL; applied iff ncells < 1e5
 * (R/SHARP.R:504).  labels[n]: SrowColor after the un-shuffle, as integers (R/SHARP.R:775-778) -- for
 * large=1 with one block, the position of the cluster in unique(fColor).  vie (may be NULL): row-major n x p =
 * enE/K.  x0 (may be NULL): row-major n x *x0_cols, room for n x max_x0_cols. */
int sharp_run(sharp_ctx *ctx, int m, int64_t n, const double *dense, const int64_t *colptr, const int32_t *rowidx,
              const double *val, const double *colsum, const sharp_rm_dev *rm, const int64_t *reind,
              const sharp_run_params *prm, int32_t *labels, double *vie, double *x0, int *x0_cols, int max_x0_cols);

/* ---- device-resident variant (bench `value`, SHARP_unlimited parts) ---------------------------------
 * sharp_expr_upload copies one expression matrix (or part) to the device (asynchronously on the context's
 * stream when the host buffers are pinned); sharp_run_dev is sharp_run on it and additionally keeps viE on
 * the device to feed sharp_centroids (the colMeans of R/sMetaC.R:58-63 for SHARP_unlimited's global sMetaC). */
typedef struct sharp_expr_dev sharp_expr_dev;
int sharp_expr_upload(sharp_ctx *ctx, int m, int64_t n, const double *dense, const int64_t *colptr,
                      const int32_t *rowidx, const double *val, sharp_expr_dev **out);
void sharp_expr_free(sharp_expr_dev *e);
int sharp_run_dev(sharp_ctx *ctx, const sharp_expr_dev *e, const double *colsum, const sharp_rm_dev *rm,
                  const int64_t *reind, const sharp_run_params *prm, int32_t *labels, double *vie, double *x0,
                  int *x0_cols, int max_x0_cols);
/* centroids of the LAST sharp_run/sharp_run_dev on this context: for cluster ids 1..nclust (values of `labels`,
 * e.g. pred_clusters after the host relabel), cen row-major nclust x p = colMeans(viE[labels == c, ]). */
int sharp_centroids(sharp_ctx *ctx, int64_t n, const int32_t *labels, int nclust, double *cen, int64_t *counts);

/* ---- a17: the loop over parts of SHARP_unlimited / SHARP_unlimited3, fused ------------------------------
 * Replaces  for (i in 1:nnp) y[[i]] = SHARP(scExp[[i]], reduced.ndim = p, prep = FALSE, logflag = FALSE, rM = rM, ...)
 * (R/SHARP_unlimited.R:125-149, R/SHARP_unlimited3.R:103-131) for parts that take the SHARP_large path, plus the
 * colMeans of every part-level cluster that the global sMetaC needs (R/sMetaC.R:58-63).  The reference runs the parts
 * one after the other; here `group` parts share every block-clustering launch (their (member, block) problems are
 * independent) and `lanes` groups are in flight on different streams, so that the latency-bound agglomeration of one
 * group overlaps the tensor-pipe-bound distance kernel of another.  Per part: pred[n] = pred_clusters of SHARP()
 * (after the merge of clusters with < small_thre cells when N.cluster is NULL and n > 1e4, R/SHARP.R:816-825, and the
 * first-appearance relabel, :828-832), nclust = N.pred_cluster, cen = row-major nclust x p centroids of viE (room for
 * cen_cap rows), counts[cen_cap]. */
typedef struct {
    int64_t n;                 /* cells of this part */
    const sharp_expr_dev *dev; /* device-resident part (sharp_expr_upload), or NULL and host slots below */
    const double *dense;       /* host: dense column-major m x n, or */
    const int64_t *colptr;     /*       dgCMatrix slots p / i / x */
    const int32_t *rowidx;
    const double *val;
    const int64_t *reind;      /* 1-based `set.seed(50); sample(n)` of this part or NULL (applied iff n < 1e5) */
    int32_t *pred;             /* out [n] */
    int nclust;                /* out */
    double *cen;               /* out (may be NULL) */
    int64_t *counts;           /* out (may be NULL) */
    int sharded;               /* 1 (needs sharp_comm_init): EVERY rank passes this part, in the same order relative to the other
                                  sharded parts, and its blocks are dealt over the ranks (sharp_run_params.shard); every rank
                                  receives its full outputs.  0: the part belongs to the calling rank alone */
} sharp_part;
int sharp_run_parts(sharp_ctx *ctx, int m, int nparts, sharp_part *parts, const sharp_rm_dev *rm,
                    const sharp_run_params *prm, int small_thre, int cen_cap, int group, int lanes);
/* Optional head start for sharp_run_parts on HOST buffers: enqueues the upload of the FIRST PART of the same `parts`
 * array (same group / lanes arguments) and returns at once, so that the copy overlaps host work the caller still has
 * to do before sharp_run_parts (SHARP_unlimited generates its ranM matrices there, R/SHARP_unlimited.R:97-104).  Takes
 * effect from the second call on a context (the staging buffers must already have their size); otherwise a no-op.
 * The buffers must stay unchanged until sharp_run_parts returns. */
int sharp_parts_prefetch(sharp_ctx *ctx, int m, int nparts, sharp_part *parts, int group, int lanes);
/* How sharp_run_parts / sharp_parts_prefetch split `nparts` parts into groups (pure host logic, no device needed):
 * gstart[0..*ngroups] receives the group boundaries (gstart[g] .. gstart[g+1]-1 are the parts of group g).
 * group / lanes <= 0 select the defaults; host_data != 0: the parts are host buffers (half-size first group). */
int sharp_plan_groups(int nparts, int host_data, int group, int lanes, int *gstart, int cap, int *ngroups, int *group_used,
                      int *lanes_used);
/* cap (GB) of the distance-matrix workspace per context; the (member, block) problems of a group run in waves of
 * as many problems as fit (default 48) */
int sharp_ctx_set_block_budget(sharp_ctx *ctx, int gigabytes);
/* on = 1: sharp_run_parts enqueues every stage of every part on the context's own stream (no overlap between
 * kernels), so that the per-kernel profile (sharp_prof_*) times each launch alone -- same launches and grids as the
 * concurrent run.  Measurement aid; default 0. */
int sharp_ctx_set_serial(sharp_ctx *ctx, int on);

/* Per-member results of the LAST sharp_run/sharp_run_dev on this context (SHARP_small's `allrpinfo`,
 * R/SHARP.R:366-385): rowcolor[n] = the member's colour index per cell (getrowColor), inde[n*p] row-major = the
 * member's projection (`tmp$mat`).  Either may be NULL.  Rows are in the order the run clustered them, i.e. the
 * input order for SHARP_small (SHARP_large shuffles; it has no allrpinfo). */
int sharp_last_member(sharp_ctx *ctx, int k, int64_t n, int32_t *rowcolor, double *inde);
/* viE = enE/K of the LAST run (row-major n x p, un-shuffled), for callers that did not ask sharp_run for it. */
int sharp_last_vie(sharp_ctx *ctx, int64_t n, int p, double *vie);

/* sMetaC given precomputed centroids (what SHARP_unlimited's global step needs when parts live on several
 * GPUs): cen row-major nC x p in unique(fColor) order; ncells_total drives the k-range tweak (R/sMetaC.R:101-119).
 * tf[nC]. */
int sharp_smetac_centroids(sharp_ctx *ctx, int nC, int p, const double *cen, int64_t ncells_total,
                           const sharp_hc_params *prm, int32_t *tf);

/* The label tail of SHARP_unlimited on the host (R/SHARP_unlimited.R:166-183), pure host code: part t's cells are
 * pred[part_start[t] .. part_start[t+1]) with 1-based part-level cluster ids; out[i] = tf[part_off[t] + pred[i] - 1];
 * clusters with fewer than merge_thre cells are merged into the smallest such id (:168-176; merge_thre <= 0: skipped);
 * ids are renumbered 1.. by decreasing size, ties in the string order of the ids (:180-183).  counts (optional, ntf
 * entries) receives the sizes of the new ids 1..*n_labels. */
int sharp_labels_combine(int nparts, const int64_t *part_start, const int32_t *part_off, const int32_t *pred,
                         const int32_t *tf, int ntf, int merge_thre, int32_t *out, int64_t *counts, int *n_labels);

/* clusterID = match(y, unique(y)) (R/SHARP.R:429-432, 828-832) for ids in 0..nvals-1, pure host code: codes 1.. in order
 * of first appearance; uniq (optional, nvals entries) receives unique(y).  Returns the number of distinct ids, -1 for an
 * id out of range. */
int sharp_first_appearance_codes(const int32_t *y, int64_t n, int nvals, int32_t *codes, int32_t *uniq);

/* ---- streaming ingestion for SHARP_unlimited3 -------------------------------------------------------
 * Replaces  mat = readRDS(allfiles[i])  (R/SHARP_unlimited3.R:105, freed at :124-125) for parts kept in the raw dgCMatrix
 * container SHCSC001 (64-byte header: "SHCSC001", int32 m, int32 0, int64 n, int64 nnz; then the slots p (int64[n+1]),
 * i (int32[nnz]), x (double[nnz]), every section padded to 64 bytes, little-endian).  The reader fills caller buffers --
 * pinned ones from sharp_host_alloc make the following H2D copy asynchronous -- with pread() from `threads` threads, so a
 * host thread can read part i + 1 while sharp_run_parts works on part i.  Pure host code apart from the pinned allocator. */
int sharp_host_alloc(void **ptr, size_t bytes);
void sharp_host_free(void *ptr);
int sharp_csc_file_info(const char *path, int *m, int64_t *n, int64_t *nnz);
int sharp_csc_file_read(const char *path, int64_t *colptr, int32_t *rowidx, double *val, int threads);

/* Binds the calling thread (and threads created after) to the CPUs next to the context's GPU, read from
 * /sys/bus/pci/devices/<bus id>/local_cpulist -- what `numactl --cpunodebind` does for a launcher that knows the
 * topology; host buffers first touched afterwards live on the GPU's NUMA node.  One process per GPU calls it once,
 * before allocating its input buffers.  cpulist (optional) receives the list applied ("" = topology not exposed, no-op). */
int sharp_ctx_bind_host(sharp_ctx *ctx, char *cpulist, int cpulist_len);

/* ---- multi-GPU: one process -- or one host thread -- per GPU, NCCL over NVLink behind this ABI -----------
 * Replaces the gather side of foreach / doParallel (`.combine`, R/SHARP.R:554, 627-635, 692; R/SHARP_unlimited3.R:137-147)
 * across GPUs.  The path shards by parts and cell blocks with no data-path collective; what is exchanged are block-level
 * labels, cluster counts and reduced-space rows / centroids in front of the meta-clustering steps.  NCCL is loaded at
 * run time (dlopen), so single-GPU use does not need it.
 *   rank 0:     sharp_comm_unique_id(id)          -> 128 bytes, carried to the other ranks by the caller (socket, file, MPI ...)
 *   every rank: sharp_comm_init(ctx, id, rank, world)   attaches an ncclComm_t to the context (and its sub-contexts)
 * The collectives below work on HOST buffers (staged through the device); the sharded runs use the communicator directly
 * on device buffers. */
#define SHARP_COMM_ID_BYTES 128
int sharp_comm_unique_id(unsigned char *id, int id_len);
int sharp_comm_init(sharp_ctx *ctx, const unsigned char *id, int rank, int world);
int sharp_comm_destroy(sharp_ctx *ctx);
int sharp_comm_info(sharp_ctx *ctx, int *rank, int *world, int *nccl_version);
/* bytes[r] bytes from rank r (same array on every rank); recv gets the concatenation in rank order */
int sharp_comm_allgatherv(sharp_ctx *ctx, const void *send, const int64_t *bytes, void *recv);
int sharp_comm_bcast(sharp_ctx *ctx, void *buf, int64_t bytes, int root);
int sharp_comm_barrier(sharp_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif
