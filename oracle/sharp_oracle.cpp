/*
 * sharp_oracle.cpp -- CPU ORACLE: a literal restatement of the reference's hot path
 * (shibiaowan/SHARP, R) plus the base-R / CRAN primitives it reaches.
 *
 * TEST INFRASTRUCTURE ONLY (see sharp_oracle.h).  PARITY UNPINNED (no reference golden vectors exist).
 * Every function cites the reference file:line it follows.  Loops deliberately follow the order
 * of operations of the R / C / Fortran code they restate (sequential sums in ascending index order,
 * first-strict-minimum tie-breaks), not the fastest order.
 *
 * Build: g++ -O2 -fopenmp -ffp-contract=off -shared -fPIC (see oracle/Makefile).  -ffp-contract=off
 * matters: R's binaries for generic x86-64 do not fuse multiply-adds.
 */
#include "sharp_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <numeric>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifdef ORACLE_LDOUBLE
typedef long double acc_t; /* R's LDOUBLE on x86 */
#else
typedef double acc_t; /* R built with --disable-long-double */
#endif

namespace {

std::string g_err;

int fail(int code, const std::string &msg) {
#pragma omp critical(orc_err)
    g_err = msg;
    return code;
}

const double kInf = std::numeric_limits<double>::infinity();

/* ------------------------------------------------------------------------------------------------
 * Projection: 1/sqrt(p) * t(rM) %*% E'   (R/RPmat.R:32; R/SHARP.R:579; R/SHARP_unlimited2.R:401)
 * Operator precedence in R: (1/sqrt(p)) * t(x) is formed first (a sparse matrix with entries
 * x/sqrt(p)), then multiplied by the dense matrix; the sparse x dense product accumulates
 * out[j] += a[j,i] * E'[i] over genes i in ascending order.
 * ---------------------------------------------------------------------------------------------- */
double r_round_digits(double x, int digits) {
    /* base::round(x, digits) -- R >= 4.0.0 "closest representable candidate, ties to even" (approximation;
     * used only by the SHARP_fpart variant, R/SHARP_unlimited2.R:410). */
    if (digits < 0 || x == 0.0 || !std::isfinite(x)) return x;
    double p10 = std::pow(10.0, (double)digits);
    double xd = x * p10;
    double fl = std::floor(xd), ce = std::ceil(xd);
    double lo = fl / p10, hi = ce / p10;
    double dl = x - lo, dh = hi - x;
    if (dl < dh) return lo;
    if (dh < dl) return hi;
    return (std::fmod(fl, 2.0) == 0.0) ? lo : hi;
}

inline double transform_value(double v, double cs, bool has_cs, double norm_mul, int logkind) {
    if (has_cs) v = v / cs * norm_mul; /* t(t(x)/colSums(x)) * 1e6          R/SHARP.R:113 */
    if (logkind == 2) v = std::log2(v + 1.0);        /* log2(newE + 1)      R/SHARP.R:344,570 */
    else if (logkind == 10) v = std::log10(v + 1.0); /* log10(newE + 1)     R/SHARP_unlimited2.R:391 */
    return v;
}

void project_one_cell(int m, const double *dense_col, const int32_t *ridx, const double *rval, int64_t nnz,
                      double cs, bool has_cs, double norm_mul, int logkind, int round_digits, int p,
                      const int32_t *rm_colptr, const int32_t *rm_rowidx, const double *rm_x, double inv_sqrt_p,
                      std::vector<double> &col, double *out) {
    /* densify the transformed column (data.matrix(newE), R/SHARP.R:574) */
    if (dense_col) {
        for (int i = 0; i < m; i++) col[i] = transform_value(dense_col[i], cs, has_cs, norm_mul, logkind);
    } else {
        double z = transform_value(0.0, cs, has_cs, norm_mul, logkind);
        for (int i = 0; i < m; i++) col[i] = z;
        for (int64_t q = 0; q < nnz; q++) col[ridx[q]] = transform_value(rval[q], cs, has_cs, norm_mul, logkind);
    }
    for (int j = 0; j < p; j++) {
        double s = 0.0;
        for (int32_t q = rm_colptr[j]; q < rm_colptr[j + 1]; q++) {
            double a = inv_sqrt_p * rm_x[q]; /* entry of 1/sqrt(p) * t(rM) */
            s += a * col[rm_rowidx[q]];
        }
        out[j] = (round_digits >= 0) ? r_round_digits(s, round_digits) : s;
    }
}

/* ------------------------------------------------------------------------------------------------
 * scale rows + 1 - cor   (R/get_opt_hclust.R:71-72; base::scale.default; stats C cov_complete1)
 * ---------------------------------------------------------------------------------------------- */
void zscore_rows(int n, int p, const double *mat, double *z) {
    for (int i = 0; i < n; i++) {
        const double *x = mat + (size_t)i * p;
        double *o = z + (size_t)i * p;
        acc_t s = 0;
        for (int j = 0; j < p; j++) s += x[j];
        double mean = (double)(s / p); /* colMeans(t(mat)) */
        for (int j = 0; j < p; j++) o[j] = x[j] - mean;
        acc_t ss = 0;
        for (int j = 0; j < p; j++) ss += (acc_t)(o[j] * o[j]); /* sum(v^2) */
        double sd = std::sqrt((double)ss / (double)std::max(1, p - 1));
        for (int j = 0; j < p; j++) o[j] = o[j] / sd;
    }
}

void cor_rows_to_dist(int n, int p, const double *z, double *dist) {
    /* cor(t(z)): pearson between rows, complete obs (cov_complete1 with cor = TRUE), then 1 - r. */
    std::vector<double> xm(n);
    for (int i = 0; i < n; i++) {
        const double *xx = z + (size_t)i * p;
        acc_t sum = 0;
        for (int k = 0; k < p; k++) sum += xx[k];
        acc_t tmp = sum / p;
        if (std::isfinite((double)tmp)) {
            sum = 0;
            for (int k = 0; k < p; k++) sum += (xx[k] - tmp);
            tmp = tmp + sum / p;
        }
        xm[i] = (double)tmp;
    }
    const int n1 = p - 1;
    std::vector<double> cov((size_t)n * n);
#pragma omp parallel for schedule(dynamic, 8)
    for (int i = 0; i < n; i++) {
        const double *xx = z + (size_t)i * p;
        double xxm = xm[i];
        for (int j = 0; j <= i; j++) {
            const double *yy = z + (size_t)j * p;
            double yym = xm[j];
            acc_t sum = 0;
            for (int k = 0; k < p; k++) sum += (acc_t)((xx[k] - xxm) * (yy[k] - yym));
            cov[(size_t)i * n + j] = cov[(size_t)j * n + i] = (double)(sum / n1);
        }
    }
    std::vector<double> sd(n);
    for (int i = 0; i < n; i++) sd[i] = std::sqrt(cov[(size_t)i * n + i]);
    for (int i = 0; i < n; i++) {
        for (int j = 0; j < i; j++) {
            double r;
            if (sd[i] == 0 || sd[j] == 0) r = std::numeric_limits<double>::quiet_NaN();
            else {
                r = cov[(size_t)i * n + j] / (sd[i] * sd[j]);
                if (r > 1.) r = 1.;
                if (r < -1.) r = -1.;
            }
            dist[(size_t)i * n + j] = dist[(size_t)j * n + i] = 1.0 - r;
        }
        dist[(size_t)i * n + i] = 0.0;
    }
}

/* ------------------------------------------------------------------------------------------------
 * stats::hclust  (Fortran hclust.f, F. Murtagh)
 * ---------------------------------------------------------------------------------------------- */
int hclust_core(int n, const double *dist, int iopt, int32_t *ia, int32_t *ib, double *crit) {
    if (n < 2) return fail(-10, "hclust: must have n >= 2 objects to cluster");
    if (iopt < 1 || iopt > 8) return fail(-11, "hclust: invalid clustering method");
    std::vector<double> W((size_t)n * n);
    for (int i = 0; i < n; i++)
        for (int j = i + 1; j < n; j++) {
            double d = dist[(size_t)i * n + j];
            if (std::isnan(d)) return fail(-12, "hclust: NA/NaN/Inf in foreign function call (arg 10)");
            W[(size_t)i * n + j] = (iopt == 8) ? d * d : d;
        }
    auto D = [&](int i, int j) -> double & { return W[(size_t)i * n + j]; }; /* requires i < j */
    std::vector<int> nn(n, 0);
    std::vector<double> disnn(n, kInf), membr(n, 1.0);
    std::vector<char> flag(n, 1);
    int jj = 0, jm = 0, im = 0;
    for (int i = 0; i < n - 1; i++) {
        double dmin = kInf;
        for (int j = i + 1; j < n; j++)
            if (dmin > D(i, j)) { dmin = D(i, j); jm = j; }
        nn[i] = jm;
        disnn[i] = dmin;
    }
    int ncl = n;
    while (ncl > 1) {
        double dmin = kInf;
        for (int i = 0; i < n - 1; i++)
            if (flag[i] && disnn[i] < dmin) { dmin = disnn[i]; im = i; jm = nn[i]; }
        ncl--;
        int i2 = std::min(im, jm), j2 = std::max(im, jm);
        ia[n - ncl - 1] = i2 + 1;
        ib[n - ncl - 1] = j2 + 1;
        if (iopt == 8) dmin = std::sqrt(dmin);
        crit[n - ncl - 1] = dmin;
        flag[j2] = 0;
        dmin = kInf;
        for (int k = 0; k < n; k++) {
            if (!flag[k] || k == i2) continue;
            double &d1ref = (i2 < k) ? D(i2, k) : D(k, i2);
            double dind1 = d1ref;
            double dind2 = (j2 < k) ? D(j2, k) : D(k, j2);
            double d12 = D(i2, j2);
            double mi = membr[i2], mj = membr[j2], mk = membr[k];
            double r;
            switch (iopt) {
            case 1: case 8: {
                double t1 = (mi + mk) * dind1;
                double t2 = (mj + mk) * dind2;
                double t3 = mk * d12;
                r = (t1 + t2) - t3;
                r = r / (mi + mj + mk);
            } break;
            case 2: r = std::min(dind1, dind2); break;
            case 3: r = std::max(dind1, dind2); break;
            case 4: r = (mi * dind1 + mj * dind2) / (mi + mj); break;
            case 5: r = (dind1 + dind2) / 2; break;
            case 6: r = ((dind1 + dind2) - d12 / 2) / 2; break;
            default: /* 7 */ r = (mi * dind1 + mj * dind2 - mi * mj * d12 / (mi + mj)) / (mi + mj); break;
            }
            d1ref = r;
            if (i2 < k) {
                if (r < dmin) { dmin = r; jj = k; }
            } else {
                if (r < disnn[k]) { disnn[k] = r; nn[k] = i2; }
            }
        }
        membr[i2] += membr[j2];
        disnn[i2] = dmin;
        nn[i2] = jj;
        for (int i = 0; i < n - 1; i++) {
            if (flag[i] && (nn[i] == i2 || nn[i] == j2)) {
                dmin = kInf;
                for (int j = i + 1; j < n; j++)
                    if (flag[j] && D(i, j) < dmin) { dmin = D(i, j); jj = j; }
                nn[i] = jj;
                disnn[i] = dmin;
            }
        }
    }
    return 0;
}

/* cutree(h, k = ...) for a descending range of k, via the partition after n-k merges; ids by first appearance
 * (stats C cutree: observation 1 is in cluster 1, new ids in scan order).
 * v is rowmajor n x nlev with column c holding k = kmin + c. */
void cutree_levels(int n, const int32_t *ia, const int32_t *ib, int kmin, int kmax, int32_t *v, int nlev) {
    std::vector<int> cl(n);
    std::iota(cl.begin(), cl.end(), 0);
    int done = 0;
    std::vector<int> id(n);
    for (int k = kmax; k >= kmin; k--) {
        int need = n - k;
        for (; done < need; done++) {
            int a = ia[done] - 1, b = ib[done] - 1;
            for (int x = 0; x < n; x++)
                if (cl[x] == b) cl[x] = a;
        }
        std::fill(id.begin(), id.end(), 0);
        int next = 0;
        for (int x = 0; x < n; x++) {
            if (id[cl[x]] == 0) id[cl[x]] = ++next;
            v[(size_t)x * nlev + (k - kmin)] = id[cl[x]];
        }
    }
}

/* cluster::silhouette(labels, dist) C sildist + median. */
double r_median(std::vector<double> x) {
    size_t n = x.size();
    if (n == 0) return std::numeric_limits<double>::quiet_NaN();
    std::sort(x.begin(), x.end());
    size_t half = (n + 1) / 2;
    if (n % 2 == 1) return x[half - 1];
    double a = x[half - 1], b = x[half];
    /* mean(c(a, b)): R's mean does a sum/n pass and a refinement pass */
    acc_t s = ((acc_t)a + (acc_t)b) / 2;
    acc_t t = ((acc_t)a - s) + ((acc_t)b - s);
    s += t / 2;
    return (double)s;
}

void silhouette_widths(int n, const double *dist, const int32_t *lab, int k, std::vector<double> &si) {
    std::vector<double> diC((size_t)n * k, 0.0);
    std::vector<int> counts(k, 0);
    for (int i = 0; i < n; i++) {
        int ci = lab[i] - 1;
        counts[ci]++;
        for (int j = i + 1; j < n; j++) {
            int cj = lab[j] - 1;
            double d = dist[(size_t)i * n + j];
            diC[(size_t)k * i + cj] += d;
            diC[(size_t)k * j + ci] += d;
        }
    }
    si.assign(n, 0.0);
    for (int i = 0; i < n; i++) {
        size_t ki = (size_t)k * i;
        int ci = lab[i] - 1;
        bool computeSi = true;
        for (int j = 0; j < k; j++) {
            if (j == ci) {
                if (counts[j] == 1) computeSi = false;
                else diC[ki + j] /= (counts[j] - 1);
            } else diC[ki + j] /= counts[j];
        }
        double a_i = diC[ki + ci], b_i;
        if (ci == 0) b_i = diC[ki + 1];
        else b_i = diC[ki];
        for (int j = 1; j < k; j++)
            if (j != ci && b_i > diC[ki + j]) b_i = diC[ki + j];
        si[i] = (computeSi && (b_i != a_i)) ? (b_i - a_i) / std::max(a_i, b_i) : 0.;
    }
}

/* clues::get_CH(y, mem, disMethod = "1-corr")  -- restated from memory (parity unpinned):
 * rows standardised (mean 0, sd 1) when ncol > 1, then CH = [B/(g-1)] / [W/(n-g)]. */
double ch_index_core(int n, int p, const double *y, const int32_t *lab, int g, bool standardise) {
    std::vector<double> ys;
    const double *Y = y;
    if (standardise && p > 1) {
        ys.resize((size_t)n * p);
        for (int i = 0; i < n; i++) {
            const double *x = y + (size_t)i * p;
            acc_t s = 0;
            for (int j = 0; j < p; j++) s += x[j];
            double mean = (double)(s / p);
            acc_t ss = 0;
            for (int j = 0; j < p; j++) ss += (acc_t)((x[j] - mean) * (x[j] - mean));
            double sd = std::sqrt((double)ss / (double)(p - 1));
            for (int j = 0; j < p; j++) ys[(size_t)i * p + j] = (x[j] - mean) / sd;
        }
        Y = ys.data();
    }
    std::vector<double> cen((size_t)g * p, 0.0), tot(p, 0.0);
    std::vector<int> cnt(g, 0);
    for (int i = 0; i < n; i++) {
        int c = lab[i] - 1;
        cnt[c]++;
        for (int j = 0; j < p; j++) {
            cen[(size_t)c * p + j] += Y[(size_t)i * p + j];
            tot[j] += Y[(size_t)i * p + j];
        }
    }
    for (int c = 0; c < g; c++)
        for (int j = 0; j < p; j++) cen[(size_t)c * p + j] /= cnt[c];
    for (int j = 0; j < p; j++) tot[j] /= n;
    double B = 0.0, Wt = 0.0;
    for (int c = 0; c < g; c++) {
        double s = 0.0;
        for (int j = 0; j < p; j++) {
            double d = cen[(size_t)c * p + j] - tot[j];
            s += d * d;
        }
        B += cnt[c] * s;
    }
    for (int i = 0; i < n; i++) {
        int c = lab[i] - 1;
        double s = 0.0;
        for (int j = 0; j < p; j++) {
            double d = Y[(size_t)i * p + j] - cen[(size_t)c * p + j];
            s += d * d;
        }
        Wt += s;
    }
    return (B / (g - 1)) / (Wt / (n - g));
}

bool is_symmetric_like_r(int nrow, int ncol, const double *mat) {
    /* isSymmetric.matrix: square and all.equal(object, t(object), tolerance = 100 * .Machine$double.eps) */
    if (nrow != ncol) return false;
    double sum_abs = 0.0, sum_diff = 0.0;
    for (int i = 0; i < nrow; i++)
        for (int j = 0; j < ncol; j++) {
            double a = mat[(size_t)i * ncol + j], b = mat[(size_t)j * ncol + i];
            sum_abs += std::fabs(a);
            sum_diff += std::fabs(a - b);
        }
    double xn = sum_abs / ((double)nrow * ncol);
    double xy = sum_diff / ((double)nrow * ncol);
    const double tol = 100 * std::numeric_limits<double>::epsilon();
    if (std::isfinite(xn) && xn > tol) xy /= xn; /* all.equal.numeric: relative unless mean(|target|) tiny */
    return !(xy > tol);
}

/* get_opt_hclust (R/get_opt_hclust.R:33-244). */
struct OptHclust {
    std::vector<int32_t> f;
    std::vector<int32_t> v; /* rowmajor n x nlev */
    int nlev = 0;
    std::vector<double> msil, chind, height;
    int optn = 0, oind = 0;
    double maxsil = 0;
};

int opt_hclust_core(int nrow, int ncol, const double *mat, int symmetric, const orc_hc_params &P, OptHclust &R) {
    const int n = nrow;
    int hm = P.hmethod ? P.hmethod : ORC_WARD_D;
    int minN = P.min_n > 0 ? P.min_n : 2;
    int maxN = P.max_n > 0 ? P.max_n : 40;
    bool sym = symmetric < 0 ? is_symmetric_like_r(nrow, ncol, mat) : (symmetric != 0);
    if (sym && nrow != ncol) return fail(-13, "get_opt_hclust: symmetric input must be square");
    std::vector<double> dist((size_t)n * n), zmat;
    const double *my = mat; /* `my = mat` -- R/get_opt_hclust.R:111 (after the scale() reassignment at :71) */
    int myp = ncol;
    if (sym) {
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++) /* as.dist() keeps the lower triangle (row > col) */
                dist[(size_t)i * n + j] = (i == j) ? 0.0 : 1.0 - mat[(size_t)std::max(i, j) * n + std::min(i, j)];
    } else {
        zmat.resize((size_t)n * ncol);
        zscore_rows(n, ncol, mat, zmat.data());
        cor_rows_to_dist(n, ncol, zmat.data(), dist.data());
        my = zmat.data();
    }
    std::vector<int32_t> ia(n - 1 > 0 ? n - 1 : 1), ib(n - 1 > 0 ? n - 1 : 1);
    R.height.assign(n - 1 > 0 ? n - 1 : 0, 0.0);
    int rc = hclust_core(n, dist.data(), hm, ia.data(), ib.data(), R.height.data());
    if (rc) return rc;

    if (P.n_cluster != 0) { /* fixed-k branch, R/get_opt_hclust.R:90-107 */
        int k = P.n_cluster;
        if (k < 2) return fail(-14, "The given N.cluster is less than 2, which is not suitable for clustering!");
        if (k > n) return fail(-15, "cutree: elements of 'k' must be between 1 and n");
        R.nlev = 1;
        R.v.assign(n, 0);
        cutree_levels(n, ia.data(), ib.data(), k, k, R.v.data(), 1);
        R.f = R.v;
        if (k >= n) return fail(-16, "silhouette: k >= n gives NA (R error at sil[, 3])");
        std::vector<double> si;
        silhouette_widths(n, dist.data(), R.f.data(), k, si);
        R.msil.assign(1, r_median(si));
        /* clusterCrit::intCriteria(data.matrix(mat), v, "Calinski_Harabasz") -- plain CH, recorded only */
        R.chind.assign(1, ch_index_core(n, myp, my, R.f.data(), k, false));
        R.optn = k;
        R.oind = 1;
        R.maxsil = R.msil[0];
        return 0;
    }

    int upper = std::min(maxN, n - 1);
    if (upper < minN) return fail(-21, "get_opt_hclust: fewer than minN.cluster+1 objects (R: nc = minN:upper runs backwards and silhouette() returns NA)");
    int nlev = upper - minN + 1;
    R.nlev = nlev;
    R.v.assign((size_t)n * nlev, 0);
    cutree_levels(n, ia.data(), ib.data(), minN, upper, R.v.data(), nlev);
    R.msil.assign(nlev, 0.0);
    R.chind.assign(nlev, 0.0);
    int err = 0;
#pragma omp parallel for schedule(dynamic, 1) if (n >= 512)
    for (int i = 0; i < nlev; i++) {
        std::vector<int32_t> lab(n);
        for (int x = 0; x < n; x++) lab[x] = R.v[(size_t)x * nlev + i];
        std::vector<double> si;
        silhouette_widths(n, dist.data(), lab.data(), minN + i, si);
        R.msil[i] = r_median(si);
        R.chind[i] = ch_index_core(n, myp, my, lab.data(), minN + i, true);
    }
    if (err) return err;
    /* selection rule, R/get_opt_hclust.R:162-217 */
    double mx = R.msil[0];
    for (int i = 1; i < nlev; i++)
        if (R.msil[i] > mx) mx = R.msil[i];
    std::vector<int> tmp;
    for (int i = 0; i < nlev; i++)
        if (R.msil[i] == mx) tmp.push_back(i + 1);
    int oind;
    if (tmp.size() > 1) oind = tmp[(tmp.size() + 1) / 2 - 1]; /* tmp[ceiling(length(tmp)/2)] */
    else if (tmp.size() == 1) oind = tmp[0];
    else return fail(-17, "get_opt_hclust: silhouette medians are NaN");
    if (mx <= P.sil_thre) {
        oind = 1; /* which.max(CHind): first maximum, NaN skipped */
        {
            int best = -1;
            for (int i = 0; i < nlev; i++)
                if (!std::isnan(R.chind[i]) && (best < 0 || R.chind[i] > R.chind[best])) best = i;
            if (best < 0) return fail(-18, "get_opt_hclust: all CH indices are NaN");
            oind = best + 1;
        }
        if (oind == 1) {
            int nh = (int)R.height.size();
            int nt = std::min(10, nh);
            const double *t = R.height.data() + (nh - nt); /* tail(h$height, n = 10) */
            int pind = -1;
            for (int i = 0; i + 1 < nt; i++) {
                double dif = t[i + 1] - t[i];
                if (dif > (P.height_ntimes - 1) * t[i]) { pind = i; break; } /* which.max(flag) */
            }
            if (pind >= 0) {
                double opth = (t[pind] + t[pind + 1]) / 2;
                /* cutree(h, h = opth): k = n + 1 - which.max(c(height, Inf) > opth); needs sorted heights */
                for (int i = 0; i + 1 < nh; i++)
                    if (R.height[i + 1] < R.height[i])
                        return fail(-19, "cutree: the 'height' component of 'tree' is not sorted (increasingly)");
                int idx = nh; /* 0-based position of Inf */
                for (int i = 0; i < nh; i++)
                    if (R.height[i] > opth) { idx = i; break; }
                int kcut = n + 1 - (idx + 1);
                oind = kcut - 1; /* length(unique(optv)) - 1, used as a COLUMN index (quirk B2) */
            }
        }
    }
    if (oind < 1 || oind > nlev) return fail(-22, "get_opt_hclust: subscript out of bounds (oind, quirk B2)");
    R.oind = oind;
    R.f.resize(n);
    for (int x = 0; x < n; x++) R.f[x] = R.v[(size_t)x * nlev + (oind - 1)];
    {
        std::vector<int32_t> u(R.f);
        std::sort(u.begin(), u.end());
        R.optn = (int)(std::unique(u.begin(), u.end()) - u.begin());
    }
    R.maxsil = mx;
    return 0;
}

/* getrowColor (R/getrowColor.R:35-68): cluster id -> colour index, wrapping modulo 40. */
int getrowcolor_core(int n, int p, const double *emat, const orc_hc_params &P, int32_t *color, double *maxsil) {
    OptHclust R;
    int rc = opt_hclust_core(n, p, emat, 0, P, R);
    if (rc) return rc;
    /* unf = unique(as.character(f)); the j-th unique value gets colour j (wrapped) */
    std::vector<int> first_rank;
    std::vector<int32_t> seen;
    for (int x = 0; x < n; x++) {
        int32_t fv = R.f[x];
        int j = -1;
        for (size_t q = 0; q < seen.size(); q++)
            if (seen[q] == fv) { j = (int)q; break; }
        if (j < 0) { seen.push_back(fv); j = (int)seen.size() - 1; }
        int c = j + 1;
        if (c > 40) { c = c % 40; if (c == 0) c = 40; }
        color[x] = c;
    }
    if (maxsil) *maxsil = R.maxsil;
    return 0;
}

/* decimal-string order of two non-negative ints: table() on a character vector sorts its levels as strings */
bool str_less(int a, int b) {
    char sa[16], sb[16];
    snprintf(sa, sizeof sa, "%d", a);
    snprintf(sb, sizeof sb, "%d", b);
    return strcmp(sa, sb) < 0;
}

/* names(sort(table(d), decreasing = TRUE)): distinct values ordered by count desc, ties by string order */
void sorted_table(const std::vector<int> &d, std::vector<std::pair<int, int>> &out /* (value,count) */) {
    std::vector<int> u(d);
    std::sort(u.begin(), u.end(), str_less);
    u.erase(std::unique(u.begin(), u.end()), u.end());
    out.clear();
    for (int val : u) out.push_back({val, (int)std::count(d.begin(), d.end(), val)});
    std::stable_sort(out.begin(), out.end(),
                     [](const std::pair<int, int> &a, const std::pair<int, int> &b) { return a.second > b.second; });
}

/* wMetaC (R/wMetaC.R:15-226) */
int wmetac_core(int N, int C, const int32_t *labels, const orc_hc_params &P, std::vector<int32_t> &finalc,
                std::vector<int32_t> &uC, std::vector<double> &x0, std::vector<double> *w1_out) {
    /* getA + AA (R/wMetaC.R:24-25, 242-283): AA[i,j] = (#members in which i and j share a cluster) / C */
    std::vector<double> w1(N);
    for (int i = 0; i < N; i++) {
        double rs = 0.0; /* rowSums(newAA): ascending column order */
        for (int j = 0; j < N; j++) {
            int cnt = 0;
            for (int c = 0; c < C; c++) cnt += (labels[(size_t)c * N + i] == labels[(size_t)c * N + j]);
            if (cnt != 0) {
                double x = (double)cnt / (double)C;
                rs += x * (1 - x); /* R/wMetaC.R:31 */
            }
        }
        double w0 = 4.0 / N * rs;          /* R/wMetaC.R:41 */
        w1[i] = (w0 + 0.01) / (1 + 0.01);  /* R/wMetaC.R:43-44 */
    }
    if (w1_out) *w1_out = w1;
    /* x = paste(nC[,i], "_", i); R = unique(x)  (R/wMetaC.R:60-67): global cluster ids by first appearance */
    std::vector<int> gid((size_t)N * C);
    std::vector<std::vector<int>> members; /* ascending cell indices per global cluster */
    for (int c = 0; c < C; c++) {
        std::vector<std::pair<int32_t, int>> seen; /* (label, gid) */
        for (int i = 0; i < N; i++) {
            int32_t l = labels[(size_t)c * N + i];
            int g = -1;
            for (auto &s : seen)
                if (s.first == l) { g = s.second; break; }
            if (g < 0) {
                g = (int)members.size();
                members.emplace_back();
                seen.push_back({l, g});
            }
            gid[(size_t)c * N + i] = g;
            members[g].push_back(i);
        }
    }
    const int allC = (int)members.size();
    if (allC < 2) return fail(-23, "wMetaC: combn(allC, 2) needs at least 2 clusters");
    /* S via getss (R/wMetaC.R:70-77, 299-320) */
    std::vector<double> S((size_t)allC * allC, 0.0);
    std::vector<int> col_of(allC);
    for (int c = 0; c < C; c++)
        for (int i = 0; i < N; i++) col_of[gid[(size_t)c * N + i]] = c;
    for (int k = 0; k < allC; k++) {
        S[(size_t)k * allC + k] = 1.0;
        for (int j = k + 1; j < allC; j++) {
            const std::vector<int> &a = members[k], &b = members[j];
            int cj = col_of[j], ck = col_of[k];
            /* intersect(a, b): elements of a that are in b, in a's order */
            acc_t si = 0;
            int ni = 0;
            for (int i : a)
                if (gid[(size_t)cj * N + i] == j) { si += w1[i]; ni++; }
            double ss = 0;
            if (ni != 0) {
                /* union(a, b) = unique(c(a, b)): a, then the elements of b not in a */
                acc_t su = 0;
                for (int i : a) su += w1[i];
                for (int i : b)
                    if (gid[(size_t)ck * N + i] != k) su += w1[i];
                ss = (double)si / (double)su;
            }
            S[(size_t)k * allC + j] = S[(size_t)j * allC + k] = ss;
        }
    }
    OptHclust H;
    int rc = opt_hclust_core(allC, allC, S.data(), 1, P, H);
    if (rc) return rc;
    /* newnC[] <- tf[match(q, R)]; finalC = names(sort(table(d), decreasing = TRUE)[1])  (R/wMetaC.R:141-143).
     * newnC stays a CHARACTER matrix, so table() orders its levels as strings. */
    finalc.assign(N, 0);
    std::vector<std::vector<std::pair<int, int>>> tabs(N);
    for (int i = 0; i < N; i++) {
        std::vector<int> d(C);
        for (int c = 0; c < C; c++) d[c] = H.f[gid[(size_t)c * N + i]];
        sorted_table(d, tabs[i]);
        finalc[i] = tabs[i][0].first;
    }
    auto uniq = [&](const std::vector<int32_t> &v) {
        std::vector<int32_t> u;
        for (int32_t x : v)
            if (std::find(u.begin(), u.end(), x) == u.end()) u.push_back(x);
        return u;
    };
    uC = uniq(finalc);
    if ((int)uC.size() == 1) { /* R/wMetaC.R:148-161 (quirk B7: n0 is always 1) */
        for (int i = 0; i < N; i++) {
            if (tabs[i].size() < 2)
                return fail(-20, "wMetaC: missing value where TRUE/FALSE needed (one-cluster fallback, R/wMetaC.R:152)");
            finalc[i] = (tabs[i][1].second >= 1 * 0.5) ? tabs[i][1].first : tabs[i][0].first;
        }
        uC = uniq(finalc);
    }
    const int NC = (int)uC.size();
    /* x0 (R/wMetaC.R:180-208) */
    x0.assign((size_t)N * NC, 0.0);
    for (int i = 0; i < N; i++) {
        std::vector<double> y0(NC, 0.0);
        for (int c = 0; c < C; c++) {
            int t = H.f[gid[(size_t)c * N + i]];
            for (int q = 0; q < NC; q++)
                if (uC[q] == t) y0[q] += 1;
        }
        int xind = (int)(std::find(uC.begin(), uC.end(), finalc[i]) - uC.begin());
        x0[(size_t)i * NC + xind] = 1;
        for (int q = 0; q < NC; q++)
            if (q != xind && y0[q] != 0) x0[(size_t)i * NC + q] = 0.5 * y0[q] / y0[xind];
    }
    return 0;
}

/* stats::cor(x, y) for two vectors (C cov_complete2) */
double cor_vec(int n, const double *x, const double *y) {
    auto mean2 = [&](const double *v) {
        acc_t sum = 0;
        for (int k = 0; k < n; k++) sum += v[k];
        acc_t tmp = sum / n;
        if (std::isfinite((double)tmp)) {
            sum = 0;
            for (int k = 0; k < n; k++) sum += (v[k] - tmp);
            tmp = tmp + sum / n;
        }
        return (double)tmp;
    };
    double xm = mean2(x), ym = mean2(y);
    int n1 = n - 1;
    acc_t sum = 0;
    for (int k = 0; k < n; k++) sum += (acc_t)((x[k] - xm) * (y[k] - ym));
    double ans = (double)(sum / n1);
    auto sdev = [&](const double *v, double vm) {
        acc_t s = 0;
        for (int k = 0; k < n; k++) s += (acc_t)((v[k] - vm) * (v[k] - vm));
        return std::sqrt((double)(s / n1));
    };
    double sx = sdev(x, xm), sy = sdev(y, ym);
    if (sx == 0 || sy == 0) return std::numeric_limits<double>::quiet_NaN();
    double r = ans / (sx * sy);
    if (r > 1.) r = 1.;
    if (r < -1.) r = -1.;
    return r;
}

/* sMetaC (R/sMetaC.R:17-210).  labels: arbitrary int codes; R = unique() by first appearance. */
int smetac_from_centroids(int nC, int p, const std::vector<double> &aG, int64_t ncells, orc_hc_params P, std::vector<int32_t> &tf);

int smetac_core(int64_t ncells, int p, const int32_t *labels, const double *se1, orc_hc_params P,
                std::vector<int32_t> &finalcolor, std::vector<int32_t> &tf) {
    /* R = unique(rerowColor) */
    std::vector<int32_t> Rl;
    std::vector<int> code(ncells);
    {
        std::vector<std::pair<int32_t, int>> sorted; /* map label -> index, kept sorted for speed */
        for (int64_t i = 0; i < ncells; i++) {
            int32_t l = labels[i];
            auto it = std::lower_bound(sorted.begin(), sorted.end(), std::make_pair(l, -1));
            if (it == sorted.end() || it->first != l) {
                it = sorted.insert(it, {l, (int)Rl.size()});
                Rl.push_back(l);
            }
            code[i] = it->second;
        }
    }
    const int nC = (int)Rl.size();
    if (nC < 2) return fail(-24, "sMetaC: combn(nC, 2) needs at least 2 clusters");
    /* aG[t, ] = colMeans(sE1[cluster t, ])  (R/sMetaC.R:58-63): ascending row order */
    std::vector<double> aG((size_t)nC * p, 0.0);
    {
        std::vector<acc_t> sum((size_t)nC * p, 0);
        std::vector<int64_t> cnt(nC, 0);
        for (int64_t i = 0; i < ncells; i++) {
            int c = code[i];
            cnt[c]++;
            const double *row = se1 + (size_t)i * p;
            acc_t *s = sum.data() + (size_t)c * p;
            for (int j = 0; j < p; j++) s[j] += row[j];
        }
        for (int c = 0; c < nC; c++)
            for (int j = 0; j < p; j++) aG[(size_t)c * p + j] = (double)(sum[(size_t)c * p + j] / cnt[c]);
    }
    int rc0 = smetac_from_centroids(nC, p, aG, ncells, P, tf);
    if (rc0) return rc0;
    finalcolor.resize(ncells);
    for (int64_t i = 0; i < ncells; i++) finalcolor[i] = tf[code[i]];
    return 0;
}

/* everything of sMetaC after the centroids (R/sMetaC.R:67-182); ncells drives the k-range tweak (:101-119) */
int smetac_from_centroids(int nC, int p, const std::vector<double> &aG, int64_t ncells, orc_hc_params P, std::vector<int32_t> &tf) {
    /* S[i,j] = cor(aG[i,], aG[j,])  (R/sMetaC.R:67-85) */
    std::vector<double> S((size_t)nC * nC, 0.0);
#pragma omp parallel for schedule(dynamic, 4)
    for (int i = 0; i < nC; i++) {
        S[(size_t)i * nC + i] = 1.0;
        for (int j = i + 1; j < nC; j++) {
            double r = cor_vec(p, aG.data() + (size_t)i * p, aG.data() + (size_t)j * p);
            S[(size_t)i * nC + j] = S[(size_t)j * nC + i] = r;
        }
    }
    /* k-range tweak (R/sMetaC.R:101-119) */
    int minN = P.min_n, maxN = P.max_n;
    int mm = (int)(ncells / 10000);
    if (ncells < 1000000) {
        int baseN = std::min(std::max(mm, 2), 10);
        if (minN == 2 && std::min(maxN, nC) - baseN >= 3) minN = baseN;
    } else {
        int mm3 = (int)(ncells / 50000), mm2 = (int)(ncells / 5000);
        maxN = std::max(maxN, mm2);
        minN = std::max(minN, mm3);
    }
    P.min_n = minN;
    P.max_n = maxN;
    OptHclust H;
    int rc = opt_hclust_core(nC, nC, S.data(), 1, P, H);
    if (rc) return rc;
    /* "second best if 2 clusters" rule (R/sMetaC.R:139-151) */
    int n = (int)H.msil.size();
    std::vector<int32_t> uf(H.f);
    std::sort(uf.begin(), uf.end());
    int nuf = (int)(std::unique(uf.begin(), uf.end()) - uf.begin());
    tf.assign(nC, 0);
    if (n > 1 && nuf == 2 && H.maxsil > P.sil_thre) {
        std::vector<double> s0(H.msil);
        std::sort(s0.begin(), s0.end());
        double s1 = s0[n - 2]; /* sort(s0, partial = n-1)[n-1] */
        int s2 = -1;
        for (int i = 0; i < n; i++)
            if (H.msil[i] == s1) { s2 = i; break; } /* quirk B4: first match */
        for (int x = 0; x < nC; x++) tf[x] = H.v[(size_t)x * H.nlev + s2];
    } else {
        tf = H.f;
    }
    return 0;
}

/* merge clusters with < 10 cells into min(as.numeric(s))  (R/SHARP.R:816-825, R/SHARP_unlimited.R:168-177) */
void merge_small_clusters(std::vector<int32_t> &lab) {
    std::vector<int32_t> u(lab);
    std::sort(u.begin(), u.end());
    u.erase(std::unique(u.begin(), u.end()), u.end());
    std::vector<int32_t> small;
    for (int32_t val : u)
        if (std::count(lab.begin(), lab.end(), val) < 10) small.push_back(val);
    if (small.empty()) return;
    int32_t target = *std::min_element(small.begin(), small.end());
    for (auto &l : lab)
        if (std::find(small.begin(), small.end(), l) != small.end()) l = target;
}

/* clusterID = match(y, unique(y))  (R/SHARP.R:429-432, 828-831) */
int relabel_first_appearance(const std::vector<int32_t> &y, int32_t *out) {
    std::vector<int32_t> uy;
    for (size_t i = 0; i < y.size(); i++) {
        auto it = std::find(uy.begin(), uy.end(), y[i]);
        if (it == uy.end()) { uy.push_back(y[i]); out[i] = (int32_t)uy.size(); }
        else out[i] = (int32_t)(it - uy.begin()) + 1;
    }
    return (int)uy.size();
}

/* folds (R/SHARP.R:513-536): block id (1-based) of every (shuffled) position */
void make_folds(int64_t ncells, int ng, std::vector<int> &folds, int &T) {
    T = (int)((ncells + ng - 1) / ng);
    folds.assign(ncells, 1);
    if (T > 1) {
        /* cut(seq(1, T*ng), breaks = T) -> fold j covers (j-1)*ng+1 .. j*ng */
        std::vector<int> f((size_t)T * ng);
        for (int64_t i = 0; i < (int64_t)T * ng; i++) f[i] = (int)(i / ng) + 1;
        int64_t nt = ncells - (int64_t)(T - 2) * ng;
        /* nind = which(folds == T-1); folds[nind[floor(nt/2) + 1:ng]] = T (NA subscripts ignored) */
        int64_t base = (int64_t)(T - 2) * ng;
        for (int64_t q = nt / 2; q < ng; q++) f[base + q] = T;
        for (int64_t i = 0; i < ncells; i++) folds[i] = f[i];
    }
}

} // namespace

/* ================================================================================================
 * C ABI
 * ============================================================================================== */
extern "C" {

const char *oracle_last_error(void) {
    return g_err.c_str();
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

int oracle_rp_project(int m, int n, const double *e_val, const int32_t *e_rowidx, const int64_t *e_colptr,
                      const int64_t *cells, int64_t ncell, const double *colsum, double norm_mul, int logkind,
                      int round_digits, int p, const int32_t *rm_colptr, const int32_t *rm_rowidx,
                      const double *rm_x, double *out) {
    if (m <= 0 || p <= 0) return fail(-1, "rp_project: bad dimensions");
    const double inv_sqrt_p = 1.0 / std::sqrt((double)p);
    const bool csc = (e_colptr != nullptr);
    int bad = 0;
#pragma omp parallel
    {
        std::vector<double> col(m);
#pragma omp for schedule(dynamic, 16)
        for (int64_t c = 0; c < ncell; c++) {
            int64_t src = cells ? cells[c] : c;
            if (src < 0 || src >= n) { bad = 1; continue; }
            double cs = colsum ? colsum[src] : 1.0;
            if (csc)
                project_one_cell(m, nullptr, e_rowidx + e_colptr[src], e_val + e_colptr[src],
                                 e_colptr[src + 1] - e_colptr[src], cs, colsum != nullptr, norm_mul, logkind,
                                 round_digits, p, rm_colptr, rm_rowidx, rm_x, inv_sqrt_p, col, out + (size_t)c * p);
            else
                project_one_cell(m, e_val + (size_t)src * m, nullptr, nullptr, 0, cs, colsum != nullptr, norm_mul,
                                 logkind, round_digits, p, rm_colptr, rm_rowidx, rm_x, inv_sqrt_p, col,
                                 out + (size_t)c * p);
        }
    }
    if (bad) return fail(-2, "rp_project: cell index out of range");
    return 0;
}

int oracle_zscore_corrdist(int n, int p, const double *mat, double *zmat, double *dist) {
    std::vector<double> z;
    double *zp = zmat;
    if (!zp) { z.resize((size_t)n * p); zp = z.data(); }
    zscore_rows(n, p, mat, zp);
    cor_rows_to_dist(n, p, zp, dist);
    return 0;
}

int oracle_hclust(int n, const double *dist, int method, int32_t *ia, int32_t *ib, double *crit) {
    return hclust_core(n, dist, method, ia, ib, crit);
}

int oracle_cutree_k(int n, const int32_t *ia, const int32_t *ib, int k, int32_t *labels) {
    if (k < 1 || k > n) return fail(-15, "cutree: elements of 'k' must be between 1 and n");
    cutree_levels(n, ia, ib, k, k, labels, 1);
    return 0;
}

int oracle_silhouette_median(int n, const double *dist, const int32_t *labels, int k, double *sil_out,
                             double *median_out) {
    if (k <= 1 || k >= n) return fail(-16, "silhouette: needs 2 <= k <= n-1");
    std::vector<double> si;
    silhouette_widths(n, dist, labels, k, si);
    if (sil_out) std::copy(si.begin(), si.end(), sil_out);
    if (median_out) *median_out = r_median(si);
    return 0;
}

int oracle_get_ch(int n, int p, const double *y, const int32_t *labels, int k, double *ch_out) {
    *ch_out = ch_index_core(n, p, y, labels, k, true);
    return 0;
}

int oracle_opt_hclust(int nrow, int ncol, const double *mat, int symmetric, const orc_hc_params *prm, int32_t *f,
                      int32_t *v, int *nlev_out, double *msil, double *chind, double *height, int *optn,
                      double *maxsil, int *oind) {
    OptHclust R;
    int rc = opt_hclust_core(nrow, ncol, mat, symmetric, *prm, R);
    if (rc) return rc;
    if (f) std::copy(R.f.begin(), R.f.end(), f);
    if (v) std::copy(R.v.begin(), R.v.end(), v);
    if (nlev_out) *nlev_out = R.nlev;
    if (msil) std::copy(R.msil.begin(), R.msil.end(), msil);
    if (chind) std::copy(R.chind.begin(), R.chind.end(), chind);
    if (height) std::copy(R.height.begin(), R.height.end(), height);
    if (optn) *optn = R.optn;
    if (maxsil) *maxsil = R.maxsil;
    if (oind) *oind = R.oind;
    return 0;
}

int oracle_getrowcolor(int n, int p, const double *emat, const orc_hc_params *prm, int32_t *color, double *maxsil) {
    return getrowcolor_core(n, p, emat, *prm, color, maxsil);
}

int oracle_wmetac(int N, int C, const int32_t *labels, const orc_hc_params *prm, int32_t *finalc, int *ncluster,
                  double *x0, int max_x0_cols, double *w1_out) {
    std::vector<int32_t> fc, uC;
    std::vector<double> x0v, w1;
    int rc = wmetac_core(N, C, labels, *prm, fc, uC, x0v, &w1);
    if (rc) return rc;
    std::copy(fc.begin(), fc.end(), finalc);
    int NC = (int)uC.size();
    if (ncluster) *ncluster = NC;
    if (x0) {
        if (NC > max_x0_cols) return fail(-3, "wmetac: x0 buffer too small");
        std::copy(x0v.begin(), x0v.end(), x0);
    }
    if (w1_out) std::copy(w1.begin(), w1.end(), w1_out);
    return 0;
}

int oracle_smetac(int64_t ncells, int p, const int32_t *labels, const double *se1, const orc_hc_params *prm,
                  int32_t *finalcolor, int32_t *tf, int *nc_out) {
    std::vector<int32_t> fc, t;
    int rc = smetac_core(ncells, p, labels, se1, *prm, fc, t);
    if (rc) return rc;
    std::copy(fc.begin(), fc.end(), finalcolor);
    if (tf) std::copy(t.begin(), t.end(), tf);
    if (nc_out) *nc_out = (int)t.size();
    return 0;
}

/* sMetaC given the cluster centroids (what SHARP_unlimited's global step looks like when the parts live elsewhere) */
int oracle_smetac_centroids(int nC, int p, const double *cen, int64_t ncells_total, const orc_hc_params *prm, int32_t *tf) {
    if (nC < 2) return fail(-24, "sMetaC: combn(nC, 2) needs at least 2 clusters");
    std::vector<double> aG(cen, cen + (size_t)nC * p);
    std::vector<int32_t> t;
    int rc = smetac_from_centroids(nC, p, aG, ncells_total, *prm, t);
    if (rc) return rc;
    std::copy(t.begin(), t.end(), tf);
    return 0;
}

int oracle_sharp(int m, int64_t n, const double *e_val, const int32_t *e_rowidx, const int64_t *e_colptr,
                 const double *colsum, double norm_mul, const orc_sharp_params *prm, const int32_t *rm_colptr,
                 const int32_t *rm_rowidx, const double *rm_x, const int64_t *rm_nnz_off, const int64_t *reind,
                 int32_t *pred, int *npred, double *vie, double *x0, int *x0_cols, int max_x0_cols) {
    const orc_sharp_params &Q = *prm;
    const int K = Q.ensize_k, p = Q.p;
    const int64_t ncells = n;
    const bool shuffle = Q.large && reind && ncells < 100000; /* R/SHARP.R:504 */
    std::vector<int> folds;
    int T = 1;
    if (Q.large) make_folds(ncells, Q.partition_ncells, folds, T);
    else folds.assign(ncells, 1);
    std::vector<int64_t> start(T + 1, 0);
    for (int64_t i = 0; i < ncells; i++) start[folds[i]]++;
    for (int t = 1; t <= T; t++) start[t] += start[t - 1];
    /* source column of every (shuffled) position: E = E[, reind] */
    std::vector<int64_t> src(ncells);
    for (int64_t i = 0; i < ncells; i++) src[i] = shuffle ? reind[i] - 1 : i;

    orc_hc_params ind = Q.hc;
    ind.n_cluster = Q.ind_n_cluster;
    const int logkind = Q.logflag ? (Q.logkind ? Q.logkind : 2) : 0;

    std::vector<int32_t> enrp((size_t)ncells * K); /* column-major ncells x K colour ids */
    std::vector<double> proj((size_t)K * ncells * p);
    int err = 0;
    /* enlist = foreach(k) %:% foreach(t) %dopar% {...}   (R/SHARP.R:554-618 / 350-387) */
#pragma omp parallel for collapse(2) schedule(dynamic, 1)
    for (int k = 0; k < K; k++) {
        for (int t = 0; t < T; t++) {
            if (err) continue;
            int64_t s = start[t], e = start[t + 1];
            int nt = (int)(e - s);
            double *pe = proj.data() + ((size_t)k * ncells + s) * p;
            int rc = oracle_rp_project(m, (int)n, e_val, e_rowidx, e_colptr, src.data() + s, nt, colsum, norm_mul,
                                       logkind, Q.round_digits, p, rm_colptr + (size_t)k * (p + 1),
                                       rm_rowidx + rm_nnz_off[k], rm_x + rm_nnz_off[k], pe);
            if (!rc) rc = getrowcolor_core(nt, p, pe, ind, enrp.data() + (size_t)k * ncells + s, nullptr);
            if (rc) {
#pragma omp atomic write
                err = rc;
            }
        }
    }
    if (err) return err;
    /* enE = sum over k (in k order) of pE1   (R/SHARP.R:629-635 / 393-399) */
    std::vector<double> enE((size_t)ncells * p, 0.0);
    for (int k = 0; k < K; k++)
        for (size_t q = 0; q < (size_t)ncells * p; q++) enE[q] = enE[q] + proj[(size_t)k * ncells * p + q];
    proj.clear();
    proj.shrink_to_fit();
    std::vector<double> E1((size_t)ncells * p);
    for (size_t q = 0; q < (size_t)ncells * p; q++) E1[q] = enE[q] / K;

    /* per-block wMetaC (R/SHARP.R:692-709) or one wMetaC (R/SHARP.R:401) */
    orc_hc_params wp = Q.hc;
    wp.n_cluster = Q.large ? Q.enp_n_cluster : Q.n_cluster;
    std::vector<std::vector<int32_t>> bfc(T), buC(T);
    std::vector<std::vector<double>> bx0(T);
#pragma omp parallel for schedule(dynamic, 1)
    for (int t = 0; t < T; t++) {
        if (err) continue;
        int64_t s = start[t], e = start[t + 1];
        int nt = (int)(e - s);
        std::vector<int32_t> lab((size_t)nt * K);
        for (int k = 0; k < K; k++)
            for (int i = 0; i < nt; i++) lab[(size_t)k * nt + i] = enrp[(size_t)k * ncells + s + i];
        int rc = wmetac_core(nt, K, lab.data(), wp, bfc[t], buC[t], bx0[t], nullptr);
        if (rc) {
#pragma omp atomic write
            err = rc;
        }
    }
    if (err) return err;

    std::vector<int32_t> Srow(ncells);
    std::vector<double> x0mat; /* rowmajor ncells x ncol_x0, shuffled order */
    int ncol_x0 = 0;
    if (!Q.large) {
        for (int64_t i = 0; i < ncells; i++) Srow[i] = bfc[0][i];
        ncol_x0 = (int)buC[0].size();
        x0mat = bx0[0];
    } else {
        /* fColor = paste(finalC, "en", t); uC = unique(fColor)  (R/SHARP.R:712-731) */
        std::vector<int32_t> fcode(ncells);
        std::vector<int> off(T + 1, 0);
        for (int t = 0; t < T; t++) off[t + 1] = off[t] + (int)buC[t].size();
        int lenuC = off[T];
        for (int t = 0; t < T; t++)
            for (int64_t i = start[t]; i < start[t + 1]; i++) {
                int32_t f = bfc[t][i - start[t]];
                int q = (int)(std::find(buC[t].begin(), buC[t].end(), f) - buC[t].begin());
                fcode[i] = off[t] + q; /* index into uC (first-appearance order, block by block) */
            }
        std::vector<int32_t> stf;
        if (T == 1) {
            for (int64_t i = 0; i < ncells; i++) Srow[i] = fcode[i];
            ncol_x0 = lenuC;
        } else {
            orc_hc_params sp = Q.hc;
            sp.n_cluster = Q.n_cluster;
            std::vector<int32_t> fc;
            int rc = smetac_core(ncells, p, fcode.data(), E1.data(), sp, fc, stf);
            if (rc) return rc;
            for (int64_t i = 0; i < ncells; i++) Srow[i] = fc[i];
            std::vector<int32_t> u(stf);
            std::sort(u.begin(), u.end());
            ncol_x0 = (int)(std::unique(u.begin(), u.end()) - u.begin());
        }
        if (x0) {
            if (ncol_x0 > max_x0_cols) return fail(-3, "sharp: x0 buffer too small");
            x0mat.assign((size_t)ncells * ncol_x0, 0.0);
            for (int t = 0; t < T; t++) {
                int q = (int)buC[t].size();
                for (int64_t i = start[t]; i < start[t + 1]; i++)
                    for (int c = 0; c < q; c++) {
                        double val = bx0[t][(size_t)(i - start[t]) * q + c];
                        if (T == 1) x0mat[(size_t)i * ncol_x0 + off[t] + c] = val;
                        else x0mat[(size_t)i * ncol_x0 + (stf[off[t] + c] - 1)] += val; /* rowSums(sx0[, si]) */
                    }
            }
        }
    }
    /* un-shuffle (R/SHARP.R:775-783) */
    std::vector<int32_t> finalrow(ncells);
    for (int64_t i = 0; i < ncells; i++) finalrow[src[i]] = Srow[i];
    if (vie)
        for (int64_t i = 0; i < ncells; i++)
            std::copy(E1.begin() + (size_t)i * p, E1.begin() + (size_t)(i + 1) * p, vie + (size_t)src[i] * p);
    if (x0) {
        if (ncol_x0 > max_x0_cols) return fail(-3, "sharp: x0 buffer too small");
        for (int64_t i = 0; i < ncells; i++)
            std::copy(x0mat.begin() + (size_t)i * ncol_x0, x0mat.begin() + (size_t)(i + 1) * ncol_x0,
                      x0 + (size_t)src[i] * ncol_x0);
    }
    if (x0_cols) *x0_cols = ncol_x0;
    /* small-cluster merge + relabel (R/SHARP.R:418-432, 816-832) */
    if (Q.n_cluster == 0 && ncells > 10000) merge_small_clusters(finalrow);
    int np = relabel_first_appearance(finalrow, pred);
    if (npred) *npred = np;
    return 0;
}

int oracle_unlimited_combine(int64_t ncells, int p, const int32_t *part_of, const int32_t *pred, const double *e1,
                             const orc_hc_params *prm, int n_cluster, int32_t *final_labels, int *nfinal) {
    /* fColor = paste(pred_clusters, "s", i): a distinct code per (part, cluster) */
    std::vector<int32_t> code(ncells);
    for (int64_t i = 0; i < ncells; i++) code[i] = part_of[i] * 100000 + pred[i];
    orc_hc_params sp = *prm;
    sp.n_cluster = n_cluster;
    std::vector<int32_t> fc, tf;
    int rc = smetac_core(ncells, p, code.data(), e1, sp, fc, tf);
    if (rc) return rc;
    if (n_cluster == 0 && ncells > 10000) merge_small_clusters(fc);
    /* x = sort(table(finalrowColor), decreasing = TRUE); map names(x) -> 1..  (R/SHARP_unlimited.R:180-183) */
    std::vector<int> d(fc.begin(), fc.end());
    std::vector<int> u(d);
    std::sort(u.begin(), u.end(), str_less);
    u.erase(std::unique(u.begin(), u.end()), u.end());
    std::vector<std::pair<int, int64_t>> tab;
    for (int val : u) tab.push_back({val, (int64_t)std::count(d.begin(), d.end(), val)});
    std::stable_sort(tab.begin(), tab.end(),
                     [](const std::pair<int, int64_t> &a, const std::pair<int, int64_t> &b) { return a.second > b.second; });
    for (int64_t i = 0; i < ncells; i++) {
        int q = 0;
        while (tab[q].first != fc[i]) q++;
        final_labels[i] = q + 1;
    }
    if (nfinal) *nfinal = (int)tab.size();
    return 0;
}

} /* extern "C" */
