// rp_common.cuh -- device helpers shared by the projection kernels (rp_project.cu, rp_project_v3.cu).
#pragma once
#include "devutil.cuh"
#include "internal.cuh"

namespace sharp {

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
    double *out;
};

__device__ __forceinline__ uint32_t atoms_add(uint32_t addr, uint32_t v) {
    uint32_t old;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ void reds_add_if(uint32_t addr, uint32_t v, uint32_t pred) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p red.shared.add.u32 [%0], %1;\n\t}" ::"r"(addr), "r"(v), "r"(pred) : "memory");
}


}  // namespace sharp
