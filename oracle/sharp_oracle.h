/*
 * sharp_oracle.h -- C ABI of the CPU ORACLE for the SHARP hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a CPU restatement of the reference's algorithm
 * (shibiaowan/SHARP, pure R) used as the checker for the CUDA path.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product (sharp_b200/, include/sharp_b200.h) never links, imports or calls it.
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or known-answer data for this
 * path and there is no R interpreter in the authoring container, so this
 * restatement is pinned only by (i) hand-computed micro known-answers for each R primitive,
 * (ii) scipy / scikit-learn cross-checks (tests/test_oracle_*.py) and (iii) the well-known
 * set.seed(42); runif(3) values for the RNG emulation used by the host.
 *
 * Arithmetic note: R on x86 accumulates sum()/rowSums()/colMeans()/cor() in 80-bit long double;
 * R built with --disable-long-double (and every platform without an extended type) uses double.
 * The oracle restates the double configuration by default (so that the tie-sensitive small
 * computations can be reproduced bit-for-bit on a GPU); compile with -DORACLE_LDOUBLE to get the
 * extended-precision accumulators.
 *
 * Conventions: all matrices are dense, fp64.  "rowmajor n x p" means element (i,j) at [i*p+j].
 * Labels are 1-based like R.  Every function returns 0 on success, a negative code on failure and
 * leaves a message retrievable through oracle_last_error().
 */
#ifndef SHARP_ORACLE_H
#define SHARP_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* hclust method codes, in the order of stats::hclust's METHODS vector (iOpt = code). */
enum { ORC_WARD_D = 1, ORC_SINGLE = 2, ORC_COMPLETE = 3, ORC_AVERAGE = 4, ORC_MCQUITTY = 5,
       ORC_MEDIAN = 6, ORC_CENTROID = 7, ORC_WARD_D2 = 8 };

typedef struct {
    int hmethod;          /* ORC_* (default ORC_WARD_D)                          R/get_opt_hclust.R:36-38 */
    int n_cluster;        /* 0 = NULL (automatic), else the fixed N.cluster      R/get_opt_hclust.R:90    */
    int min_n;            /* minN.cluster                                        R/get_opt_hclust.R:41-43 */
    int max_n;            /* maxN.cluster                                        R/get_opt_hclust.R:46-48 */
    double sil_thre;      /* sil.thre                                            R/get_opt_hclust.R:51-53 */
    double height_ntimes; /* height.Ntimes                                       R/get_opt_hclust.R:56-58 */
} orc_hc_params;

const char *oracle_last_error(void);
int oracle_num_threads(void);

/* ---- a2/a3/a4: projection  (R/RPmat.R:32, R/SHARP.R:567-585, R/SHARP.R:113, R/SHARP_unlimited2.R:391,410)
 * E: genes x cells, either dense column-major (e_val, e_rowidx = e_colptr = NULL) or CSC
 *    (dgCMatrix slots: e_colptr[n+1], e_rowidx[nnz], e_val[nnz]).
 * cells: optional 0-based source column for every output cell (E[, tind]); NULL = identity; ncell outputs.
 * colsum: optional per-source-column divisor: value := value / colsum[c] * norm_mul  (R/SHARP.R:113).
 * logkind: 0 none, 2 log2(x+1), 10 log10(x+1).  round_digits: <0 none, else round(., digits).
 * rm_*: the m x p ranM matrix as dgCMatrix slots (p+1 colptr, rowidx ascending per column, x values).
 * out: rowmajor ncell x p  ( = t(1/sqrt(p) * t(rM) %*% inE) ).
 */
int oracle_rp_project(int m, int n, const double *e_val, const int32_t *e_rowidx, const int64_t *e_colptr,
                      const int64_t *cells, int64_t ncell, const double *colsum, double norm_mul, int logkind,
                      int round_digits, int p, const int32_t *rm_colptr, const int32_t *rm_rowidx,
                      const double *rm_x, double *out);

/* ---- stats::as.dist(1 - cor(t(scale-rows(mat))))   (R/get_opt_hclust.R:71-72)
 * mat rowmajor n x p -> zmat (rowmajor n x p, the row z-scored matrix) and dist (full symmetric n x n, diag 0). */
int oracle_zscore_corrdist(int n, int p, const double *mat, double *zmat, double *dist);

/* ---- stats::hclust(d, method)  (Fortran hclust.f restated)
 * dist: full symmetric n x n (only i<j read).  ia/ib: 1-based cluster representatives merged at each of the
 * n-1 steps (I2<J2), crit: heights. */
int oracle_hclust(int n, const double *dist, int method, int32_t *ia, int32_t *ib, double *crit);

/* cutree(h, k) from the (ia, ib) sequence: labels numbered by first appearance. */
int oracle_cutree_k(int n, const int32_t *ia, const int32_t *ib, int k, int32_t *labels);

/* median(silhouette(labels, dist)[,3])  (cluster::sildist restated) */
int oracle_silhouette_median(int n, const double *dist, const int32_t *labels, int k, double *sil_out /* n or NULL */,
                             double *median_out);

/* clues::get_CH(y, mem, disMethod = "1-corr")  (restated from memory; parity unpinned) */
int oracle_get_ch(int n, int p, const double *y, const int32_t *labels, int k, double *ch_out);

/* ---- get_opt_hclust  (R/get_opt_hclust.R:33-244)
 * mat rowmajor nrow x ncol.  symmetric: 1 = isSymmetric branch (d = 1 - mat), 0 = feature branch, -1 = decide
 * like isSymmetric() (square and all.equal(mat, t(mat), tol = 100 eps)).
 * outputs (any may be NULL): f[nrow]; v rowmajor nrow x nlev (nlev = *nlev_out, levels min_n..min(max_n,nrow-1),
 * or 1 column in the fixed-k branch); msil[nlev]; chind[nlev]; height[nrow-1]; *optn; *maxsil; *oind (1-based).
 * v/msil/chind must have room for nrow x (max_n-min_n+1). */
int oracle_opt_hclust(int nrow, int ncol, const double *mat, int symmetric, const orc_hc_params *prm, int32_t *f,
                      int32_t *v, int *nlev_out, double *msil, double *chind, double *height, int *optn,
                      double *maxsil, int *oind);

/* ---- getrowColor  (R/getrowColor.R:17-121): labels = colour index 1..40 (ids wrap modulo 40). */
int oracle_getrowcolor(int n, int p, const double *emat, const orc_hc_params *prm, int32_t *color, double *maxsil);

/* ---- wMetaC  (R/wMetaC.R:15-226)
 * labels: column-major N x C int32 (any integer codes; equality is all that matters).
 * finalc[N]: meta-cluster id (the numeric value R stores as a string); *ncluster = length(unique(finalC));
 * x0: rowmajor N x *ncluster (caller provides N x max_x0_cols), columns in unique(finalC) order;
 * w1 (optional, N).  Returns -20 where R itself would stop (NA in the one-cluster fallback, R/wMetaC.R:152). */
int oracle_wmetac(int N, int C, const int32_t *labels, const orc_hc_params *prm, int32_t *finalc, int *ncluster,
                  double *x0, int max_x0_cols, double *w1_out);

/* ---- sMetaC  (R/sMetaC.R:17-210)
 * labels[ncells]: integer codes of rerowColor; se1 rowmajor ncells x p.
 * finalcolor[ncells]; tf[nC] (nC = number of unique labels, first-appearance order); *nc_out = nC. */
int oracle_smetac(int64_t ncells, int p, const int32_t *labels, const double *se1, const orc_hc_params *prm,
                  int32_t *finalcolor, int32_t *tf, int *nc_out);

/* sMetaC from precomputed centroids (cen rowmajor nC x p, unique(rerowColor) order); ncells_total drives the k-range
 * tweak of R/sMetaC.R:101-119.  tf[nC]. */
int oracle_smetac_centroids(int nC, int p, const double *cen, int64_t ncells_total, const orc_hc_params *prm, int32_t *tf);

/* ---- SHARP_small / SHARP_large compute for ONE expression matrix  (R/SHARP.R:339-454, 478-851)
 * Everything the R drivers decide up front is an input: `large` (0 = SHARP_small, 1 = SHARP_large), the log flag,
 * K ranM matrices (dgCMatrix slots, concatenated; rm_nnz_off[K+1] offsets into rm_rowidx/rm_x, rm_colptr K x (p+1)),
 * reind (1-based permutation from set.seed(50); sample(ncells), or NULL), partition_ncells, ncluster settings.
 * Output pred[ncells] = enresults$pred_clusters, vie rowmajor ncells x p (may be NULL), x0 rowmajor ncells x *x0_cols
 * (may be NULL; room for ncells x max_x0_cols).
 */
typedef struct {
    int large;             /* 0 SHARP_small, 1 SHARP_large */
    int logflag;           /* `flag` */
    int ensize_k;
    int p;                 /* reduced.ndim */
    int partition_ncells;
    int n_cluster;         /* N.cluster (0 = NULL) */
    int enp_n_cluster;     /* enpN.cluster (0 = NULL) */
    int ind_n_cluster;     /* indN.cluster (0 = NULL) */
    orc_hc_params hc;      /* hmethod, minN, maxN, sil.thre, height.Ntimes (n_cluster field ignored) */
    int logkind;           /* 2 (SHARP_small/large) or 10 (SHARP_fpart) */
    int round_digits;      /* <0 none; 1 in SHARP_fpart */
} orc_sharp_params;

int oracle_sharp(int m, int64_t n, const double *e_val, const int32_t *e_rowidx, const int64_t *e_colptr,
                 const double *colsum, double norm_mul, const orc_sharp_params *prm, const int32_t *rm_colptr,
                 const int32_t *rm_rowidx, const double *rm_x, const int64_t *rm_nnz_off, const int64_t *reind,
                 int32_t *pred, int *npred, double *vie, double *x0, int *x0_cols, int max_x0_cols);

/* ---- SHARP_unlimited final stage (R/SHARP_unlimited.R:151-183): global sMetaC on the stacked per-part labels
 * and viE, small-cluster merge, relabel by decreasing size.  part_of[ncells] = 1-based part id, pred[ncells] =
 * per-part pred_clusters. */
int oracle_unlimited_combine(int64_t ncells, int p, const int32_t *part_of, const int32_t *pred, const double *e1,
                             const orc_hc_params *prm, int n_cluster, int32_t *final_labels, int *nfinal);

#ifdef __cplusplus
}
#endif
#endif
