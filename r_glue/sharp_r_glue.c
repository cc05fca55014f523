/*
 * sharp_r_glue.c -- `.Call` entry points that bind libsharpb200's C ABI (include/sharp_b200.h) into R.
 *
 * NOT BUILT IN THIS IMAGE: there is no R toolchain here (no R, no Rinternals.h).  A maintainer of the reference
 * builds it inside the SHARP package with
 *     R CMD SHLIB sharp_r_glue.c -I<repo>/include -L<repo>/sharp_b200 -lsharpb200
 * (or drops it into SHARP/src/ with PKG_LIBS = -lsharpb200), adds `useDynLib(SHARP, .registration = TRUE)` to
 * NAMESPACE and replaces the R closure bodies as shown in INTEGRATION.md.  Only the stable R C API is used.
 *
 * Conventions: every SEXP input is R-owned and read-only; outputs are freshly allocated and
 * PROTECTed; device handles are EXTPTRSXPs with finalizers; errors are raised with Rf_error() only after all native
 * resources of the call are released (the C ABI returns status codes and never longjmps); labels are 1-based like R.
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <stdint.h>
#include <string.h>

#include "sharp_b200.h"

/* ---- context: one per GPU, created lazily, kept for the R session ------------------------------------------- */
static sharp_ctx *g_ctx[16];

static sharp_ctx *ctx_for(int device) {
    if (device < 0 || device >= 16) Rf_error("sharp_b200: device index out of range");
    if (!g_ctx[device] && sharp_ctx_create(device, &g_ctx[device]) != SHARP_OK)
        Rf_error("sharp_b200: %s", sharp_last_error());
    return g_ctx[device];
}

static void rm_finalizer(SEXP ptr) {
    sharp_rm_dev *rm = (sharp_rm_dev *)R_ExternalPtrAddr(ptr);
    if (rm) { sharp_rm_free(rm); R_ClearExternalPtr(ptr); }
}

static sharp_hc_params hc_from(SEXP hmethod, SEXP ncluster, SEXP minN, SEXP maxN, SEXP silthre, SEXP heightN) {
    static const char *names[] = {"ward.D", "single", "complete", "average", "mcquitty", "median", "centroid", "ward.D2"};
    sharp_hc_params p;
    const char *h = CHAR(STRING_ELT(hmethod, 0));
    p.hmethod = 0;
    for (int i = 0; i < 8; i++) if (!strcmp(h, names[i])) p.hmethod = i + 1;
    if (!p.hmethod) Rf_error("invalid clustering method '%s'", h);
    p.n_cluster = Rf_isNull(ncluster) || !Rf_isNumeric(ncluster) ? 0 : Rf_asInteger(ncluster);
    p.min_n = Rf_asInteger(minN);
    p.max_n = Rf_asInteger(maxN);
    p.sil_thre = Rf_asReal(silthre);
    p.height_ntimes = Rf_asReal(heightN);
    return p;
}

/* .Call("sharp_R_device_info", device) -> list(name, sm_count, cc, total_mem)      replaces detectCores() */
SEXP sharp_R_device_info(SEXP device) {
    char name[256];
    int sm, maj, mnr;
    size_t mem;
    if (sharp_device_info(Rf_asInteger(device), name, 256, &sm, &maj, &mnr, &mem) != SHARP_OK)
        Rf_error("sharp_b200: %s", sharp_last_error());
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 4));
    SET_VECTOR_ELT(out, 0, Rf_mkString(name));
    SET_VECTOR_ELT(out, 1, Rf_ScalarInteger(sm));
    SET_VECTOR_ELT(out, 2, Rf_ScalarReal(maj + mnr / 10.0));
    SET_VECTOR_ELT(out, 3, Rf_ScalarReal((double)mem));
    UNPROTECT(1);
    return out;
}

/* .Call("sharp_R_rm_upload", rM, device): rM = list of K dgCMatrix (ranM() results) -> external pointer.
 * The dgCMatrix slots are read in place: @i (int), @p (int), @x (double), @Dim. */
SEXP sharp_R_rm_upload(SEXP rM, SEXP device) {
    const int K = Rf_length(rM);
    if (K < 1) Rf_error("rM must be a non-empty list of dgCMatrix");
    SEXP first = VECTOR_ELT(rM, 0);
    const int *dim = INTEGER(R_do_slot(first, Rf_install("Dim")));
    const int m = dim[0], p = dim[1];
    int64_t *off = (int64_t *)R_alloc(K + 1, sizeof(int64_t));
    off[0] = 0;
    for (int k = 0; k < K; k++) off[k + 1] = off[k] + Rf_length(R_do_slot(VECTOR_ELT(rM, k), Rf_install("i")));
    int32_t *colptr = (int32_t *)R_alloc((size_t)K * (p + 1), sizeof(int32_t));
    int32_t *rowidx = (int32_t *)R_alloc((size_t)off[K] + 1, sizeof(int32_t));
    double *x = (double *)R_alloc((size_t)off[K] + 1, sizeof(double));
    for (int k = 0; k < K; k++) {
        SEXP mk = VECTOR_ELT(rM, k);
        memcpy(colptr + (size_t)k * (p + 1), INTEGER(R_do_slot(mk, Rf_install("p"))), sizeof(int32_t) * (p + 1));
        memcpy(rowidx + off[k], INTEGER(R_do_slot(mk, Rf_install("i"))), sizeof(int32_t) * (off[k + 1] - off[k]));
        memcpy(x + off[k], REAL(R_do_slot(mk, Rf_install("x"))), sizeof(double) * (off[k + 1] - off[k]));
    }
    sharp_rm_dev *rm = NULL;
    if (sharp_rm_upload(ctx_for(Rf_asInteger(device)), m, p, K, colptr, rowidx, x, off, &rm) != SHARP_OK)
        Rf_error("sharp_b200: %s", sharp_last_error());
    SEXP ptr = PROTECT(R_MakeExternalPtr(rm, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(ptr, rm_finalizer, TRUE);
    UNPROTECT(1);
    return ptr;
}

/* expression argument: a numeric matrix (dense, column-major = R's own layout) or a dgCMatrix */
typedef struct { int m; int64_t n; const double *dense; int64_t *colptr; const int32_t *rowidx; const double *val; } expr_t;

static expr_t expr_from(SEXP E) {
    expr_t e;
    memset(&e, 0, sizeof e);
    if (Rf_isMatrix(E) && Rf_isReal(E)) {
        e.m = Rf_nrows(E);
        e.n = Rf_ncols(E);
        e.dense = REAL(E);
    } else if (Rf_inherits(E, "dgCMatrix")) {
        const int *dim = INTEGER(R_do_slot(E, Rf_install("Dim")));
        e.m = dim[0];
        e.n = dim[1];
        const int *p32 = INTEGER(R_do_slot(E, Rf_install("p")));
        e.colptr = (int64_t *)R_alloc((size_t)e.n + 1, sizeof(int64_t)); /* the ABI takes 64-bit column pointers */
        for (int64_t c = 0; c <= e.n; c++) e.colptr[c] = p32[c];
        e.rowidx = INTEGER(R_do_slot(E, Rf_install("i")));
        e.val = REAL(R_do_slot(E, Rf_install("x")));
    } else Rf_error("scExp must be a numeric matrix or a dgCMatrix");
    return e;
}

/* .Call("sharp_R_rp_project", E, rm_ptr, cells, normalize, logkind, round_digits, device)
 *   -> numeric array ncell x p x K (each slice = t(projmat) of one member)           replaces R/RPmat.R:32 etc. */
SEXP sharp_R_rp_project(SEXP E, SEXP rm_ptr, SEXP cells, SEXP normalize, SEXP logkind, SEXP round_digits, SEXP device,
                        SEXP K_, SEXP p_) {
    expr_t e = expr_from(E);
    sharp_rm_dev *rm = (sharp_rm_dev *)R_ExternalPtrAddr(rm_ptr);
    if (!rm) Rf_error("rm handle was freed");
    const int K = Rf_asInteger(K_), p = Rf_asInteger(p_);
    int64_t ncell = e.n, *cidx = NULL;
    if (!Rf_isNull(cells)) { /* 1-based R indices (E[, tind]) */
        ncell = Rf_length(cells);
        cidx = (int64_t *)R_alloc((size_t)ncell, sizeof(int64_t));
        for (int64_t i = 0; i < ncell; i++) cidx[i] = (int64_t)INTEGER(cells)[i] - 1;
    }
    double *tmp = (double *)R_alloc((size_t)K * ncell * p, sizeof(double)); /* row-major per member */
    int rc = sharp_rp_project(ctx_for(Rf_asInteger(device)), e.m, e.n, e.dense, e.colptr, e.rowidx, e.val, cidx, ncell,
                              Rf_asInteger(normalize), NULL, 1e6, Rf_asInteger(logkind), Rf_asInteger(round_digits), rm, tmp);
    if (rc != SHARP_OK) Rf_error("sharp_b200: %s", sharp_last_error());
    SEXP out = PROTECT(Rf_alloc3DArray(REALSXP, (int)ncell, p, K));
    double *o = REAL(out);
    for (int k = 0; k < K; k++)
        for (int64_t i = 0; i < ncell; i++)
            for (int j = 0; j < p; j++) o[((size_t)k * p + j) * ncell + i] = tmp[((size_t)k * ncell + i) * p + j];
    UNPROTECT(1);
    return out;
}

/* .Call("sharp_R_run", E, rm_ptr, reind, large, flag, logkind, round_digits, partition.ncells, N.cluster,
 *        enpN.cluster, indN.cluster, hmethod, minN, maxN, sil.thre, height.Ntimes, normalize, forview, device, p)
 *   -> list(labels = integer(ncells), viE = matrix ncells x p or NULL, x0 = matrix or NULL)
 * replaces the compute of SHARP_small (R/SHARP.R:343-416) / SHARP_large (R/SHARP.R:502-783) / SHARP_fpart. */
SEXP sharp_R_run(SEXP E, SEXP rm_ptr, SEXP reind, SEXP large, SEXP flag, SEXP logkind, SEXP round_digits, SEXP ng,
                 SEXP Ncl, SEXP enpN, SEXP indN, SEXP hmethod, SEXP minN, SEXP maxN, SEXP silthre, SEXP heightN,
                 SEXP normalize, SEXP forview, SEXP device, SEXP p_, SEXP opts) {
    /* opts: NULL or integer(4) = c(skip_smetac, block_max_n, shard, shard_rotate) -- SHARP_fpart's two-level scheme
     * (R/SHARP_unlimited2.R:421, 477-494: per-block maxN.cluster = 40, no cross-block sMetaC) and block-sharded runs on
     * a context that carries a communicator (sharp_R_comm_init) */
    expr_t e = expr_from(E);
    sharp_rm_dev *rm = (sharp_rm_dev *)R_ExternalPtrAddr(rm_ptr);
    if (!rm) Rf_error("rm handle was freed");
    sharp_run_params q;
    memset(&q, 0, sizeof q);
    q.large = Rf_asLogical(large);
    q.logflag = Rf_asLogical(flag);
    q.logkind = Rf_asInteger(logkind);
    q.round_digits = Rf_asInteger(round_digits);
    q.partition_ncells = Rf_asInteger(ng);
    q.n_cluster = Rf_isNull(Ncl) ? 0 : Rf_asInteger(Ncl);
    q.enp_n_cluster = Rf_isNull(enpN) ? 0 : Rf_asInteger(enpN);
    q.ind_n_cluster = Rf_isNull(indN) ? 0 : Rf_asInteger(indN);
    q.hc = hc_from(hmethod, R_NilValue, minN, maxN, silthre, heightN);
    q.normalize = Rf_asInteger(normalize);
    q.norm_mul = 1e6;
    if (!Rf_isNull(opts) && Rf_length(opts) >= 4) {
        q.skip_smetac = INTEGER(opts)[0];
        q.block_max_n = INTEGER(opts)[1];
        q.shard = INTEGER(opts)[2];
        q.shard_rotate = INTEGER(opts)[3];
    }
    const int p = Rf_asInteger(p_), view = Rf_asLogical(forview);
    int64_t *re = NULL;
    if (!Rf_isNull(reind)) {
        re = (int64_t *)R_alloc((size_t)e.n, sizeof(int64_t));
        for (int64_t i = 0; i < e.n; i++) re[i] = INTEGER(reind)[i];
    }
    const int maxc = (q.hc.max_n > 63 ? q.hc.max_n : 63) + 1;
    SEXP labels = PROTECT(Rf_allocVector(INTSXP, (R_xlen_t)e.n));
    double *vie = view ? (double *)R_alloc((size_t)e.n * p, sizeof(double)) : NULL;
    double *x0 = view ? (double *)R_alloc((size_t)e.n * maxc, sizeof(double)) : NULL;
    int x0c = 0;
    int rc = sharp_run(ctx_for(Rf_asInteger(device)), e.m, e.n, e.dense, e.colptr, e.rowidx, e.val, NULL, rm, re, &q,
                       INTEGER(labels), vie, x0, &x0c, maxc);
    if (rc != SHARP_OK) { UNPROTECT(1); Rf_error("%s", sharp_last_error()); } /* SHARP_E_RSTOP quotes the R error */
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 3));
    SET_VECTOR_ELT(out, 0, labels);
    if (view) { /* row-major C -> column-major R */
        SEXP v = PROTECT(Rf_allocMatrix(REALSXP, (int)e.n, p));
        for (int64_t i = 0; i < e.n; i++) for (int j = 0; j < p; j++) REAL(v)[(size_t)j * e.n + i] = vie[(size_t)i * p + j];
        SEXP x = PROTECT(Rf_allocMatrix(REALSXP, (int)e.n, x0c));
        for (int64_t i = 0; i < e.n; i++) for (int j = 0; j < x0c; j++) REAL(x)[(size_t)j * e.n + i] = x0[(size_t)i * x0c + j];
        SET_VECTOR_ELT(out, 1, v);
        SET_VECTOR_ELT(out, 2, x);
        UNPROTECT(2);
    }
    UNPROTECT(2);
    return out;
}

/* .Call("sharp_R_run_parts", parts, rm_ptr, reinds, flag, partition.ncells, enpN.cluster, indN.cluster, hmethod, maxN,
 *        sil.thre, height.Ntimes, normalize, device, p)
 *   parts: the LIST of genes x cells matrices SHARP_unlimited was given; reinds: list of `set.seed(50); sample(n)` per part
 *   (NULL entries for parts of 1e5 cells or more)
 *   -> list(pred = list of integer vectors (clusterID per part, R/SHARP.R:828-832), cen = list of nclust x p matrices
 *      (colMeans of viE per cluster, for the global sMetaC))
 * replaces the loop `y[[i]] = SHARP(scExp[[i]], reduced.ndim = p, prep = FALSE, rM = rM, ...)` of
 * R/SHARP_unlimited.R:125-149 by ONE device call when every part takes the SHARP_large path (sharp_run_parts: groups of
 * parts share the block-clustering launches, uploads run one group ahead).  sharp_parts_prefetch may be called with
 * the same list first (e.g. before the ranM matrices are drawn) to start the first uploads early. */
SEXP sharp_R_run_parts(SEXP parts, SEXP rm_ptr, SEXP reinds, SEXP flag, SEXP ng, SEXP enpN, SEXP indN, SEXP hmethod,
                       SEXP maxN, SEXP silthre, SEXP heightN, SEXP normalize, SEXP device, SEXP p_) {
    sharp_rm_dev *rm = (sharp_rm_dev *)R_ExternalPtrAddr(rm_ptr);
    if (!rm) Rf_error("rm handle was freed");
    const int np = Rf_length(parts), p = Rf_asInteger(p_);
    sharp_run_params q;
    memset(&q, 0, sizeof q);
    q.large = 1;
    q.logflag = Rf_asLogical(flag);
    q.logkind = 2;
    q.round_digits = -1;
    q.partition_ncells = Rf_asInteger(ng);
    q.enp_n_cluster = Rf_isNull(enpN) ? 0 : Rf_asInteger(enpN);
    q.ind_n_cluster = Rf_isNull(indN) ? 0 : Rf_asInteger(indN);
    q.hc = hc_from(hmethod, R_NilValue, R_NilValue, maxN, silthre, heightN);
    q.hc.min_n = 2; /* SHARP() inside the loop runs with its default minN.cluster */
    q.normalize = Rf_asInteger(normalize);
    q.norm_mul = 1e6;
    const int cen_cap = (q.hc.max_n > 63 ? q.hc.max_n : 63) + 1;
    sharp_part *P = (sharp_part *)R_alloc((size_t)np, sizeof(sharp_part));
    memset(P, 0, (size_t)np * sizeof(sharp_part));
    SEXP pred = PROTECT(Rf_allocVector(VECSXP, np));
    int m = 0;
    for (int i = 0; i < np; i++) {
        expr_t e = expr_from(VECTOR_ELT(parts, i));
        m = e.m;
        P[i].n = e.n;
        P[i].dense = e.dense;
        P[i].colptr = e.colptr;
        P[i].rowidx = e.rowidx;
        P[i].val = e.val;
        SEXP re = Rf_isNull(reinds) ? R_NilValue : VECTOR_ELT(reinds, i);
        if (!Rf_isNull(re)) {
            int64_t *r64 = (int64_t *)R_alloc((size_t)e.n, sizeof(int64_t));
            for (int64_t c = 0; c < e.n; c++) r64[c] = INTEGER(re)[c];
            P[i].reind = r64;
        }
        SEXP lab = PROTECT(Rf_allocVector(INTSXP, (R_xlen_t)e.n));
        SET_VECTOR_ELT(pred, i, lab);
        UNPROTECT(1);
        P[i].pred = INTEGER(lab);
        P[i].cen = (double *)R_alloc((size_t)cen_cap * p, sizeof(double));
        P[i].counts = (int64_t *)R_alloc((size_t)cen_cap, sizeof(int64_t));
    }
    int rc = sharp_run_parts(ctx_for(Rf_asInteger(device)), m, np, P, rm, &q, 10, cen_cap, 0, 0);
    if (rc != SHARP_OK) { UNPROTECT(1); Rf_error("%s", sharp_last_error()); }
    SEXP cen = PROTECT(Rf_allocVector(VECSXP, np));
    for (int i = 0; i < np; i++) { /* row-major C -> column-major R */
        const int nc = P[i].nclust;
        SEXP c = PROTECT(Rf_allocMatrix(REALSXP, nc, p));
        for (int a = 0; a < nc; a++) for (int j = 0; j < p; j++) REAL(c)[(size_t)j * nc + a] = P[i].cen[(size_t)a * p + j];
        SET_VECTOR_ELT(cen, i, c);
        UNPROTECT(1);
    }
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
    SET_VECTOR_ELT(out, 0, pred);
    SET_VECTOR_ELT(out, 1, cen);
    UNPROTECT(3);
    return out;
}

/* .Call("sharp_R_opt_hclust", mat, symmetric, hmethod, N.cluster, minN, maxN, sil.thre, height.Ntimes, device)
 *   -> list(f, v, maxsil, msil, CHind, height, optN.cluster)                   replaces R/get_opt_hclust.R:66-243 */
SEXP sharp_R_opt_hclust(SEXP mat, SEXP symmetric, SEXP hmethod, SEXP Ncl, SEXP minN, SEXP maxN, SEXP silthre,
                        SEXP heightN, SEXP device) {
    const int n = Rf_nrows(mat), pc = Rf_ncols(mat);
    sharp_hc_params prm = hc_from(hmethod, Ncl, minN, maxN, silthre, heightN);
    double *rm = (double *)R_alloc((size_t)n * pc, sizeof(double)); /* column-major R -> row-major */
    for (int i = 0; i < n; i++) for (int j = 0; j < pc; j++) rm[(size_t)i * pc + j] = REAL(mat)[(size_t)j * n + i];
    const int maxlev = prm.n_cluster ? 1 : (prm.max_n - prm.min_n + 1 > 1 ? prm.max_n - prm.min_n + 1 : 1);
    SEXP f = PROTECT(Rf_allocVector(INTSXP, n));
    int *v = (int *)R_alloc((size_t)n * maxlev, sizeof(int));
    double *msil = (double *)R_alloc(maxlev, sizeof(double)), *ch = (double *)R_alloc(maxlev, sizeof(double));
    SEXP height = PROTECT(Rf_allocVector(REALSXP, n - 1));
    int nlev = 0, optn = 0, oind = 0;
    double maxsil = 0;
    int rc = sharp_opt_hclust(ctx_for(Rf_asInteger(device)), n, pc, rm, Rf_asLogical(symmetric), 1, &prm, INTEGER(f), v,
                              &nlev, msil, ch, REAL(height), &optn, &maxsil, &oind);
    if (rc != SHARP_OK) { UNPROTECT(2); Rf_error("%s", sharp_last_error()); }
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 7));
    SEXP vm = PROTECT(Rf_allocMatrix(INTSXP, n, nlev));
    for (int i = 0; i < n; i++) for (int l = 0; l < nlev; l++) INTEGER(vm)[(size_t)l * n + i] = v[(size_t)i * nlev + l];
    SEXP ms = PROTECT(Rf_allocVector(REALSXP, nlev)), chs = PROTECT(Rf_allocVector(REALSXP, nlev));
    memcpy(REAL(ms), msil, sizeof(double) * nlev);
    memcpy(REAL(chs), ch, sizeof(double) * nlev);
    SET_VECTOR_ELT(out, 0, f);
    SET_VECTOR_ELT(out, 1, vm);
    SET_VECTOR_ELT(out, 2, Rf_ScalarReal(maxsil));
    SET_VECTOR_ELT(out, 3, ms);
    SET_VECTOR_ELT(out, 4, chs);
    SET_VECTOR_ELT(out, 5, height);
    SET_VECTOR_ELT(out, 6, Rf_ScalarInteger(optn));
    UNPROTECT(6);
    return out;
}

/* .Call("sharp_R_wmetac", codes, hmethod, enN.cluster, minN, maxN, sil.thre, height.Ntimes, device)
 *   codes = integer matrix N x C of match(nC[, c], unique(nC[, c]))  -> list(finalC, x0)  replaces R/wMetaC.R:19-225 */
SEXP sharp_R_wmetac(SEXP codes, SEXP hmethod, SEXP enN, SEXP minN, SEXP maxN, SEXP silthre, SEXP heightN, SEXP device) {
    const int N = Rf_nrows(codes), C = Rf_ncols(codes);
    sharp_hc_params prm = hc_from(hmethod, enN, minN, maxN, silthre, heightN);
    const int maxc = (prm.max_n > prm.n_cluster ? prm.max_n : prm.n_cluster) + 1;
    SEXP fc = PROTECT(Rf_allocVector(INTSXP, N));
    double *x0 = (double *)R_alloc((size_t)N * maxc, sizeof(double));
    int nc = 0;
    int rc = sharp_wmetac(ctx_for(Rf_asInteger(device)), N, C, INTEGER(codes), &prm, INTEGER(fc), &nc, x0, maxc, NULL);
    if (rc != SHARP_OK) { UNPROTECT(1); Rf_error("%s", sharp_last_error()); }
    SEXP x = PROTECT(Rf_allocMatrix(REALSXP, N, nc));
    for (int i = 0; i < N; i++) for (int j = 0; j < nc; j++) REAL(x)[(size_t)j * N + i] = x0[(size_t)i * nc + j];
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
    SET_VECTOR_ELT(out, 0, fc);
    SET_VECTOR_ELT(out, 1, x);
    UNPROTECT(3);
    return out;
}

/* .Call("sharp_R_smetac", codes, sE1, hmethod, finalN.cluster, minN, maxN, sil.thre, height.Ntimes, device)
 *   codes = match(rerowColor, unique(rerowColor)) -> list(finalColor, tf)                replaces R/sMetaC.R:21-208 */
SEXP sharp_R_smetac(SEXP codes, SEXP sE1, SEXP hmethod, SEXP Ncl, SEXP minN, SEXP maxN, SEXP silthre, SEXP heightN,
                    SEXP device) {
    const int64_t n = Rf_nrows(sE1);
    const int p = Rf_ncols(sE1);
    sharp_hc_params prm = hc_from(hmethod, Ncl, minN, maxN, silthre, heightN);
    double *rm = (double *)R_alloc((size_t)n * p, sizeof(double));
    for (int64_t i = 0; i < n; i++) for (int j = 0; j < p; j++) rm[(size_t)i * p + j] = REAL(sE1)[(size_t)j * n + i];
    SEXP fc = PROTECT(Rf_allocVector(INTSXP, (R_xlen_t)n));
    int *tf = (int *)R_alloc((size_t)n, sizeof(int));
    int nC = 0;
    int rc = sharp_smetac(ctx_for(Rf_asInteger(device)), n, p, INTEGER(codes), rm, &prm, INTEGER(fc), tf, &nC);
    if (rc != SHARP_OK) { UNPROTECT(1); Rf_error("%s", sharp_last_error()); }
    SEXP t = PROTECT(Rf_allocVector(INTSXP, nC));
    memcpy(INTEGER(t), tf, sizeof(int) * nC);
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
    SET_VECTOR_ELT(out, 0, fc);
    SET_VECTOR_ELT(out, 1, t);
    UNPROTECT(3);
    return out;
}

/* SHARP_unlimited's global step when the parts were clustered on the device: centroids of the last run's viE and
 * sMetaC on stacked centroids (R/SHARP_unlimited.R:151-164 without moving ncells x p doubles through R). */
SEXP sharp_R_centroids(SEXP labels, SEXP nclust, SEXP p_, SEXP device) {
    const int64_t n = Rf_length(labels);
    const int nc = Rf_asInteger(nclust), p = Rf_asInteger(p_);
    double *cen = (double *)R_alloc((size_t)nc * p, sizeof(double));
    if (sharp_centroids(ctx_for(Rf_asInteger(device)), n, INTEGER(labels), nc, cen, NULL) != SHARP_OK)
        Rf_error("sharp_b200: %s", sharp_last_error());
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, nc, p));
    for (int c = 0; c < nc; c++) for (int j = 0; j < p; j++) REAL(out)[(size_t)j * nc + c] = cen[(size_t)c * p + j];
    UNPROTECT(1);
    return out;
}

SEXP sharp_R_smetac_centroids(SEXP cen, SEXP ncells, SEXP hmethod, SEXP Ncl, SEXP minN, SEXP maxN, SEXP silthre,
                              SEXP heightN, SEXP device) {
    const int nC = Rf_nrows(cen), p = Rf_ncols(cen);
    sharp_hc_params prm = hc_from(hmethod, Ncl, minN, maxN, silthre, heightN);
    double *rm = (double *)R_alloc((size_t)nC * p, sizeof(double));
    for (int i = 0; i < nC; i++) for (int j = 0; j < p; j++) rm[(size_t)i * p + j] = REAL(cen)[(size_t)j * nC + i];
    SEXP tf = PROTECT(Rf_allocVector(INTSXP, nC));
    int rc = sharp_smetac_centroids(ctx_for(Rf_asInteger(device)), nC, p, rm, (int64_t)Rf_asReal(ncells), &prm, INTEGER(tf));
    if (rc != SHARP_OK) { UNPROTECT(1); Rf_error("%s", sharp_last_error()); }
    UNPROTECT(1);
    return tf;
}

/* ---- multi-GPU: one R process per GPU; the NCCL communicator lives on the context ------------------------------ */
/* .Call("sharp_R_comm_unique_id") -> raw(128); rank 0 calls it and hands the bytes to the other processes (socket,
 * file, MPI -- the library never does the rendezvous itself) */
SEXP sharp_R_comm_unique_id(void) {
    SEXP id = PROTECT(Rf_allocVector(RAWSXP, SHARP_COMM_ID_BYTES));
    if (sharp_comm_unique_id(RAW(id), SHARP_COMM_ID_BYTES) != SHARP_OK) { UNPROTECT(1); Rf_error("%s", sharp_last_error()); }
    UNPROTECT(1);
    return id;
}

/* .Call("sharp_R_comm_init", id, rank, world, device): every process, same id; replaces registerDoParallel(n.cores) */
SEXP sharp_R_comm_init(SEXP id, SEXP rank, SEXP world, SEXP device) {
    if (TYPEOF(id) != RAWSXP || Rf_length(id) < SHARP_COMM_ID_BYTES) Rf_error("sharp_R_comm_init: id must be raw(128)");
    if (sharp_comm_init(ctx_for(Rf_asInteger(device)), RAW(id), Rf_asInteger(rank), Rf_asInteger(world)) != SHARP_OK)
        Rf_error("%s", sharp_last_error());
    return R_NilValue;
}

/* .Call("sharp_R_comm_allgather", x, device): x raw vector (serialize() of this rank's labels / centroids) -> list of
 * raw vectors, one per rank: the `.combine` of foreach (R/SHARP_unlimited3.R:137-147) across processes */
SEXP sharp_R_comm_allgather(SEXP x, SEXP device) {
    sharp_ctx *c = ctx_for(Rf_asInteger(device));
    int rank = 0, world = 1;
    sharp_comm_info(c, &rank, &world, NULL);
    int64_t *sizes = (int64_t *)R_alloc((size_t)world, sizeof(int64_t));
    int64_t *eight = (int64_t *)R_alloc((size_t)world, sizeof(int64_t));
    for (int r = 0; r < world; r++) eight[r] = 8;
    int64_t mine = (int64_t)XLENGTH(x);
    if (sharp_comm_allgatherv(c, &mine, eight, sizes) != SHARP_OK) Rf_error("%s", sharp_last_error());
    int64_t total = 0;
    for (int r = 0; r < world; r++) total += sizes[r];
    unsigned char *all = (unsigned char *)R_alloc((size_t)(total > 0 ? total : 1), 1);
    if (sharp_comm_allgatherv(c, RAW(x), sizes, all) != SHARP_OK) Rf_error("%s", sharp_last_error());
    SEXP out = PROTECT(Rf_allocVector(VECSXP, world));
    int64_t off = 0;
    for (int r = 0; r < world; r++) {
        SEXP v = PROTECT(Rf_allocVector(RAWSXP, (R_xlen_t)sizes[r]));
        memcpy(RAW(v), all + off, (size_t)sizes[r]);
        SET_VECTOR_ELT(out, r, v);
        UNPROTECT(1);
        off += sizes[r];
    }
    UNPROTECT(1);
    return out;
}

/* ---- streaming ingestion (SHARP_unlimited3, R/SHARP_unlimited3.R:103-131: readRDS per part) ----------------------- */
/* .Call("sharp_R_csc_read", path) -> list(Dim, p, i, x): the slots of the dgCMatrix stored in an SHCSC001 file (written
 * once with writeBin -- INTEGRATION.md 5), read by the library's threaded pread.  `p` comes back as doubles (a part
 * can hold more than 2^31 non-zeros); new("dgCMatrix", ...) takes as.integer(p) when it fits. */
SEXP sharp_R_csc_read(SEXP path) {
    const char *f = CHAR(STRING_ELT(path, 0));
    int m = 0;
    int64_t n = 0, nnz = 0;
    if (sharp_csc_file_info(f, &m, &n, &nnz) != SHARP_OK) Rf_error("%s", sharp_last_error());
    int64_t *cp = (int64_t *)R_alloc((size_t)(n + 1), sizeof(int64_t));
    SEXP ri = PROTECT(Rf_allocVector(INTSXP, (R_xlen_t)nnz)), xv = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)nnz));
    if (sharp_csc_file_read(f, cp, INTEGER(ri), REAL(xv), 4) != SHARP_OK) { UNPROTECT(2); Rf_error("%s", sharp_last_error()); }
    SEXP pp = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)(n + 1)));
    for (int64_t j = 0; j <= n; j++) REAL(pp)[j] = (double)cp[j];
    SEXP dim = PROTECT(Rf_allocVector(INTSXP, 2));
    INTEGER(dim)[0] = m; INTEGER(dim)[1] = (int)n;
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 4));
    SET_VECTOR_ELT(out, 0, dim); SET_VECTOR_ELT(out, 1, pp); SET_VECTOR_ELT(out, 2, ri); SET_VECTOR_ELT(out, 3, xv);
    UNPROTECT(5);
    return out;
}

static const R_CallMethodDef call_methods[] = {
    {"sharp_R_device_info", (DL_FUNC)&sharp_R_device_info, 1},
    {"sharp_R_rm_upload", (DL_FUNC)&sharp_R_rm_upload, 2},
    {"sharp_R_rp_project", (DL_FUNC)&sharp_R_rp_project, 9},
    {"sharp_R_run", (DL_FUNC)&sharp_R_run, 21},
    {"sharp_R_run_parts", (DL_FUNC)&sharp_R_run_parts, 14},
    {"sharp_R_opt_hclust", (DL_FUNC)&sharp_R_opt_hclust, 9},
    {"sharp_R_wmetac", (DL_FUNC)&sharp_R_wmetac, 8},
    {"sharp_R_smetac", (DL_FUNC)&sharp_R_smetac, 9},
    {"sharp_R_centroids", (DL_FUNC)&sharp_R_centroids, 4},
    {"sharp_R_smetac_centroids", (DL_FUNC)&sharp_R_smetac_centroids, 9},
    {"sharp_R_comm_unique_id", (DL_FUNC)&sharp_R_comm_unique_id, 0},
    {"sharp_R_comm_init", (DL_FUNC)&sharp_R_comm_init, 4},
    {"sharp_R_comm_allgather", (DL_FUNC)&sharp_R_comm_allgather, 2},
    {"sharp_R_csc_read", (DL_FUNC)&sharp_R_csc_read, 1},
    {NULL, NULL, 0}};

void R_init_SHARP(DllInfo *dll) {
    R_registerRoutines(dll, NULL, call_methods, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}
