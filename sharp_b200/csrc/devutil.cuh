// devutil.cuh -- small device helpers shared by the kernels.
#pragma once
#include <cuda_runtime.h>

#include <climits>
#include <cstdint>

namespace sharp {

#define SHARP_INF (__longlong_as_double(0x7ff0000000000000LL))

// (value, index) pair with the order "smaller value first, then smaller index" -- a parallel reduction
// under this order returns exactly what a sequential scan with a strict `<` returns (first minimum).
struct DI {
    double d;
    int i;
};

__device__ __forceinline__ DI di_better(DI a, DI b) { return (b.d < a.d || (b.d == a.d && b.i < a.i)) ? b : a; }

__device__ __forceinline__ DI warp_argmin(DI v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        DI w;
        w.d = __shfl_xor_sync(0xffffffffu, v.d, o);
        w.i = __shfl_xor_sync(0xffffffffu, v.i, o);
        v = di_better(v, w);
    }
    return v;
}

// Same result as warp_argmin (all lanes receive it) with three REDUX instructions instead of five shuffle rounds:
// the value is mapped to a 64-bit key whose unsigned order is the numeric order, the minimum key is found high word
// first, then the smallest index among the lanes that hold it.  Values must not be NaN (callers only keep values that
// passed a `<` test); -0.0 is canonicalised to +0.0 so that equal values tie on the index like di_better.
__device__ __forceinline__ DI warp_argmin_redux(DI v) {
    const unsigned full = 0xffffffffu;
    unsigned long long b = (unsigned long long)__double_as_longlong(v.d + 0.0);
    b ^= (b >> 63) ? ~0ull : 0x8000000000000000ull;
    const unsigned hi = (unsigned)(b >> 32), lo = (unsigned)b;
    const unsigned mh = __reduce_min_sync(full, hi);
    const unsigned ml = __reduce_min_sync(full, hi == mh ? lo : 0xffffffffu);
    const bool win = (hi == mh) && (lo == ml);
    const unsigned mi = __reduce_min_sync(full, win ? (unsigned)v.i : 0xffffffffu);
    unsigned long long k = ((unsigned long long)mh << 32) | ml;
    k ^= (k >> 63) ? 0x8000000000000000ull : ~0ull;
    DI r;
    r.d = __longlong_as_double((long long)k);
    r.i = (int)mi;
    return r;
}

// All threads of the block receive the block-wide argmin.  `scratch` is shared memory with one slot per warp.
template <int THREADS>
__device__ __forceinline__ DI block_argmin(DI v, DI *scratch) {
    v = warp_argmin_redux(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    DI r;
    if (lane < THREADS / 32) r = scratch[lane];
    else { r.d = SHARP_INF; r.i = INT_MAX; }
    r = warp_argmin_redux(r);
    __syncthreads();
    return r;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// deterministic block sum (fixed tree), result in all threads
template <int THREADS>
__device__ __forceinline__ double block_sum(double v, double *scratch) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    double r = (lane < THREADS / 32) ? scratch[lane] : 0.0;
    r = warp_sum(r);
    __syncthreads();
    return r;
}

// Exclusive prefix sum of an int array in shared memory (in place), length n, by the whole block.
// tmp: shared int[THREADS].  Returns the total in all threads.
template <int THREADS>
__device__ __forceinline__ int block_exclusive_scan(int *a, int n, int *tmp) {
    const int tid = threadIdx.x;
    const int per = (n + THREADS - 1) / THREADS;
    const int lo = tid * per, hi = min(n, lo + per);
    int s = 0;
    for (int i = lo; i < hi; i++) s += a[i];
    tmp[tid] = s;
    __syncthreads();
    // Hillis-Steele inclusive scan over THREADS partial sums
    for (int o = 1; o < THREADS; o <<= 1) {
        int v = (tid >= o) ? tmp[tid - o] : 0;
        __syncthreads();
        tmp[tid] += v;
        __syncthreads();
    }
    int run = (tid == 0) ? 0 : tmp[tid - 1];
    const int total = tmp[THREADS - 1];
    for (int i = lo; i < hi; i++) {
        int v = a[i];
        a[i] = run;
        run += v;
    }
    __syncthreads();
    return total;
}

// In-place ascending bitonic sort of a shared-memory array of length P2 (a power of two) by the whole block.
template <int THREADS>
__device__ __forceinline__ void block_bitonic_sort(double *a, int P2) {
    for (int k = 2; k <= P2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < P2 / 2; t += THREADS) {
                // index of the lower element of the t-th compare-exchange pair at distance j
                int i = 2 * t - (t & (j - 1));
                int l = i + j;
                bool up = ((i & k) == 0);
                double x = a[i], y = a[l];
                if ((x > y) == up) {
                    a[i] = y;
                    a[l] = x;
                }
            }
            __syncthreads();
        }
    }
}

// stats::median on a sorted array (R: mean of the two middle values for even n; mean() = sum/2 plus a
// refinement pass).  Plain (unfused) double operations so that the CPU oracle gets the same bits.
__device__ __forceinline__ double median_sorted(const double *a, int n) {
    const int half = (n + 1) / 2;
    if (n & 1) return a[half - 1];
    const double x = a[half - 1], y = a[half];
    double s = __ddiv_rn(__dadd_rn(x, y), 2.0);
    double t = __dadd_rn(__dsub_rn(x, s), __dsub_rn(y, s));
    s = __dadd_rn(s, __ddiv_rn(t, 2.0));
    return s;
}

__device__ __forceinline__ int next_pow2(int n) {
    int p = 1;
    while (p < n) p <<= 1;
    return p;
}

}  // namespace sharp
