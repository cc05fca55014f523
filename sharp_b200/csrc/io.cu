// io.cu -- streaming ingestion of dgCMatrix parts for SHARP_unlimited3.
//
// Replaces  mat = readRDS(allfiles[i])  (R/SHARP_unlimited3.R:105; freed again at :124-125) for parts stored in a raw
// dgCMatrix container, so that reading part i + 1 overlaps the H2D copy and the clustering of part i:
//
//   "SHCSC001" | int32 m | int32 flags (0) | int64 n | int64 nnz | 32 reserved bytes          (64-byte header)
//   int64 colptr[n + 1]   (slot `p`)        -- section padded to a multiple of 64 bytes
//   int32 rowidx[nnz]     (slot `i`, ascending inside a column)   -- padded to 64
//   double val[nnz]       (slot `x`)
//   little-endian throughout.  An R maintainer writes it with writeBin() from the three slots.
//
// The reader fills CALLER buffers (pinned host memory from sharp_host_alloc, so the following H2D copy is asynchronous)
// with pread() from several threads; nothing here touches the GPU except the pinned allocator.
#include <fcntl.h>
#include <sched.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cctype>
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "internal.cuh"

namespace {

struct CscHeader {
    char magic[8];
    int32_t m, flags;
    int64_t n, nnz;
    char reserved[32];
};
static_assert(sizeof(CscHeader) == 64, "header layout");

inline int64_t pad64(int64_t b) { return (b + 63) & ~(int64_t)63; }

int read_all(int fd, void *dst, int64_t bytes, int64_t off, int threads) {
    if (bytes <= 0) return 0;
    threads = std::max(1, std::min<int>(threads, (int)(bytes >> 24) + 1)); /* at least 16 MB per thread */
    std::vector<std::thread> th;
    std::vector<int> err((size_t)threads, 0);
    const int64_t per = (bytes + threads - 1) / threads;
    for (int t = 0; t < threads; t++)
        th.emplace_back([=, &err]() {
            int64_t a = (int64_t)t * per, b = std::min(bytes, a + per);
            unsigned char *p = reinterpret_cast<unsigned char *>(dst);
            while (a < b) {
                const ssize_t got = pread(fd, p + a, (size_t)std::min<int64_t>(b - a, (int64_t)1 << 30), off + a);
                if (got <= 0) { err[t] = got == 0 ? EIO : errno; return; }
                a += got;
            }
        });
    for (auto &x : th) x.join();
    for (int e : err)
        if (e) return e;
    return 0;
}

}  // namespace

using namespace sharp;

extern "C" {

int sharp_host_alloc(void **ptr, size_t bytes) {
    if (!ptr) return set_error(SHARP_E_ARG, "host_alloc: null output");
    cudaError_t e = cudaMallocHost(ptr, std::max<size_t>(bytes, 64));
    if (e != cudaSuccess) {
        *ptr = nullptr;
        cudaGetLastError();
        return set_error(e == cudaErrorMemoryAllocation ? SHARP_E_NOMEM : SHARP_E_CUDA, "cudaMallocHost(%zu bytes): %s", bytes, cudaGetErrorString(e));
    }
    return 0;
}

void sharp_host_free(void *ptr) {
    if (ptr) cudaFreeHost(ptr);
}

int sharp_csc_file_info(const char *path, int *m, int64_t *n, int64_t *nnz) {
    if (!path) return set_error(SHARP_E_ARG, "csc_file_info: null path");
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return set_error(SHARP_E_ARG, "cannot open %s: %s", path, strerror(errno));
    CscHeader h;
    const ssize_t got = pread(fd, &h, sizeof h, 0);
    struct stat st;
    fstat(fd, &st);
    close(fd);
    if (got != (ssize_t)sizeof h || memcmp(h.magic, "SHCSC001", 8) != 0)
        return set_error(SHARP_E_ARG, "%s is not a SHCSC001 dgCMatrix file", path);
    if (h.m <= 0 || h.n < 0 || h.nnz < 0) return set_error(SHARP_E_ARG, "%s: bad dimensions", path);
    const int64_t need = 64 + pad64((h.n + 1) * 8) + pad64(h.nnz * 4) + h.nnz * 8;
    if ((int64_t)st.st_size < need) return set_error(SHARP_E_ARG, "%s is truncated (%lld of %lld bytes)", path, (long long)st.st_size, (long long)need);
    if (m) *m = h.m;
    if (n) *n = h.n;
    if (nnz) *nnz = h.nnz;
    return 0;
}

// colptr[n + 1], rowidx[nnz], val[nnz]: caller buffers of at least the sizes sharp_csc_file_info reported
int sharp_csc_file_read(const char *path, int64_t *colptr, int32_t *rowidx, double *val, int threads) {
    int m;
    int64_t n, nnz;
    SHARP_TRY(sharp_csc_file_info(path, &m, &n, &nnz));
    if (!colptr || (nnz > 0 && (!rowidx || !val))) return set_error(SHARP_E_ARG, "csc_file_read: null buffer");
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return set_error(SHARP_E_ARG, "cannot open %s: %s", path, strerror(errno));
    const int64_t o1 = 64, o2 = o1 + pad64((n + 1) * 8), o3 = o2 + pad64(nnz * 4);
    int e = read_all(fd, colptr, (n + 1) * 8, o1, 1);
    if (!e) e = read_all(fd, rowidx, nnz * 4, o2, threads);
    if (!e) e = read_all(fd, val, nnz * 8, o3, threads);
    close(fd);
    if (e) return set_error(SHARP_E_ARG, "reading %s: %s", path, strerror(e));
    if (colptr[0] != 0 || colptr[n] != nnz) return set_error(SHARP_E_ARG, "%s: colptr does not match nnz", path);
    return 0;
}


// Binds the CALLING thread (and the threads it creates later) to the CPUs next to the context's GPU
// (/sys/bus/pci/devices/<bus id>/local_cpulist), so that host buffers allocated and first touched afterwards live on the
// GPU's NUMA node: with one process per GPU on a two-socket host the uploads of all ranks otherwise cross the socket
// link.  What `numactl --cpunodebind` does for a launcher that knows the topology.  cpulist (optional) receives the
// list that was applied ("" when the topology is not exposed: the call is then a no-op).
int sharp_ctx_bind_host(sharp_ctx *c, char *cpulist, int cpulist_len) {
    if (!c) return set_error(SHARP_E_ARG, "null context");
    if (cpulist && cpulist_len > 0) cpulist[0] = 0;
    char bus[64] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, c->device) != cudaSuccess) { cudaGetLastError(); return 0; }
    for (char *q = bus; *q; q++) *q = (char)tolower(*q);
    char path[160];
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/local_cpulist", bus);
    FILE *f = fopen(path, "r");
    if (!f) return 0;
    char line[4096] = {0};
    const bool got = fgets(line, sizeof line, f) != nullptr;
    fclose(f);
    if (!got) return 0;
    cpu_set_t set;
    CPU_ZERO(&set);
    int ncpu = 0;
    for (char *q = line; *q && *q != '\n';) { /* "0-31,64-95" */
        char *end;
        long a = strtol(q, &end, 10);
        if (end == q) break;
        long b = a;
        if (*end == '-') { q = end + 1; b = strtol(q, &end, 10); }
        for (long v = a; v <= b && v < CPU_SETSIZE; v++) { CPU_SET((int)v, &set); ncpu++; }
        q = (*end == ',') ? end + 1 : end;
        if (*end != ',' ) break;
    }
    if (ncpu == 0) return 0;
    if (sched_setaffinity(0, sizeof set, &set) != 0) return 0;
    if (cpulist && cpulist_len > 0) {
        size_t n = strcspn(line, "\n");
        if (n >= (size_t)cpulist_len) n = (size_t)cpulist_len - 1;
        memcpy(cpulist, line, n);
        cpulist[n] = 0;
    }
    return 0;
}

}  // extern "C"
