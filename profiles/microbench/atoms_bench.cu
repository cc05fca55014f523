// microbenchmark: shared-memory atomic throughput with random addresses (what bounds the projection scatter)
#include <cstdio>
#include <cuda_runtime.h>
constexpr int W = 5120;  // words per array
__device__ __forceinline__ unsigned nxt(unsigned &s) { s = s * 1664525u + 1013904223u; return (s >> 8) % 2540u; }
template <int MODE>
__global__ void __launch_bounds__(128, 8) k(unsigned *out, int iters) {
    __shared__ unsigned a[W];
    for (int i = threadIdx.x; i < W; i += 128) a[i] = 0;
    __syncthreads();
    unsigned s = blockIdx.x * 131u + threadIdx.x * 7919u + 1u;
    unsigned *lo = a, *hi = a + 2560;
    for (int it = 0; it < iters; it++) {
        unsigned col[8], old[8];
#pragma unroll
        for (int e = 0; e < 8; e++) col[e] = nxt(s);
        if (MODE == 0) {  // returning lo + dependent hi (current kernel)
#pragma unroll
            for (int e = 0; e < 8; e++) old[e] = atomicAdd(lo + col[e], 0x9e3779b9u);
#pragma unroll
            for (int e = 0; e < 8; e++) atomicAdd(hi + col[e], 3u + ((old[e] + 0x9e3779b9u) < 0x9e3779b9u ? 1u : 0u));
        } else if (MODE == 1) {  // one non-returning
#pragma unroll
            for (int e = 0; e < 8; e++) atomicAdd(lo + col[e], 0x100u);
        } else if (MODE == 2) {  // three non-returning (21-bit limbs, no carries)
#pragma unroll
            for (int e = 0; e < 8; e++) { atomicAdd(lo + col[e], 5u); atomicAdd(hi + col[e], 7u); atomicAdd(a + 2 * 2560 - 2560 + ((col[e] * 3u) % 2560u), 9u); }
        } else if (MODE == 3) {  // racy plain RMW 32-bit
#pragma unroll
            for (int e = 0; e < 8; e++) old[e] = lo[col[e]];
#pragma unroll
            for (int e = 0; e < 8; e++) lo[col[e]] = old[e] + 1u;
        } else if (MODE == 4) {  // racy plain RMW 64-bit
            unsigned long long *q = reinterpret_cast<unsigned long long *>(a);
            unsigned long long o[8];
#pragma unroll
            for (int e = 0; e < 8; e++) o[e] = q[col[e]];
#pragma unroll
            for (int e = 0; e < 8; e++) q[col[e]] = o[e] + 1ull;
        } else if (MODE == 5) {  // two non-returning, independent
#pragma unroll
            for (int e = 0; e < 8; e++) { atomicAdd(lo + col[e], 5u); atomicAdd(hi + col[e], 7u); }
        } else if (MODE == 6) {  // conflict-free returning atomics (lane-strided)
#pragma unroll
            for (int e = 0; e < 8; e++) old[e] = atomicAdd(lo + ((col[e] & ~31u) | (threadIdx.x & 31)), 1u);
            s += old[3];
        } else if (MODE == 7) {  // conflict-free non-returning
#pragma unroll
            for (int e = 0; e < 8; e++) atomicAdd(lo + ((col[e] & ~31u) | (threadIdx.x & 31)), 1u);
        }
    }
    __syncthreads();
    unsigned t = 0;
    for (int i = threadIdx.x; i < W; i += 128) t += a[i];
    out[blockIdx.x * 128 + threadIdx.x] = t + s;
}
template <int MODE>
void run(const char *name, int adds_per_it, unsigned *out) {
    int sms = 148, grid = sms * 8, iters = 4000;
    k<MODE><<<grid, 128>>>(out, 10);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k<MODE><<<grid, 128>>>(out, iters);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double adds = (double)grid * 128 * iters * adds_per_it;
    printf("%-44s %8.3f ms  %7.2f Gadd/s  %6.2f add/clk/SM (1.965 GHz)\n", name, ms, adds / ms / 1e6, adds / (ms * 1e-3) / sms / 1.965e9);
}
int main() {
    unsigned *out; cudaMalloc(&out, 148 * 8 * 128 * 4);
    run<0>("lo returning + dependent hi (2 atomics/add)", 8, out);
    run<1>("1 non-returning atomic/add", 8, out);
    run<5>("2 independent non-returning/add", 8, out);
    run<2>("3 independent non-returning/add", 8, out);
    run<3>("racy LDS+STS 32-bit", 8, out);
    run<4>("racy LDS+STS 64-bit", 8, out);
    run<6>("conflict-free returning atomic", 8, out);
    run<7>("conflict-free non-returning atomic", 8, out);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
