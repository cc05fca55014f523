// comm.cu -- multi-GPU plumbing behind the C ABI: one process (or host thread) per GPU, NCCL over NVLink / NVSwitch.
//
// Replaces the role of foreach / doParallel's `.combine` gather (R/SHARP.R:554, 627-635, 692; R/SHARP_unlimited3.R:137-147)
// across GPUs: the path shards by cell blocks and parts with NO data-path collective; the only exchanges are allgathers
// of block-level labels, cluster counts and reduced-space rows / centroids before the meta-clustering steps.
//
// NCCL is loaded at run time (dlopen "libnccl.so.2"): single-GPU use needs no NCCL at all, and a host process that has
// already loaded an NCCL (PyTorch bundles one) shares it by soname.  The unique id of the communicator travels through
// the caller (R: any socket / file mechanism; sharp_b200/comm.py: a TCP rendezvous on MASTER_ADDR), never through this
// library.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>
#include <vector>

#include "internal.cuh"

namespace sharp {

struct NcclApi {
    void *h = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclBroadcast) Broadcast = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
};

static NcclApi g_nccl;
static std::mutex g_nccl_mu;

static int nccl_load() {
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (g_nccl.h) return 0;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return set_error(SHARP_E_CUDA, "multi-GPU: cannot load NCCL (libnccl.so.2): %s", dlerror());
    NcclApi a;
    a.h = h;
#define SHARP_SYM(name)                                                                          \
    a.name = reinterpret_cast<decltype(a.name)>(dlsym(h, "nccl" #name));                          \
    if (!a.name) return set_error(SHARP_E_CUDA, "multi-GPU: libnccl has no symbol nccl" #name)
    SHARP_SYM(GetUniqueId);
    SHARP_SYM(CommInitRank);
    SHARP_SYM(CommDestroy);
    SHARP_SYM(AllGather);
    SHARP_SYM(Broadcast);
    SHARP_SYM(GroupStart);
    SHARP_SYM(GroupEnd);
    SHARP_SYM(GetErrorString);
    SHARP_SYM(GetVersion);
#undef SHARP_SYM
    g_nccl = a;
    return 0;
}

#define SHARP_NCCL(expr)                                                                                  \
    do {                                                                                                  \
        ncclResult_t _r = (expr);                                                                         \
        if (_r != ncclSuccess)                                                                            \
            return ::sharp::set_error(SHARP_E_CUDA, "%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,  \
                                      g_nccl.GetErrorString(_r));                                         \
    } while (0)

// allgather of per-rank segments of different sizes, device buffers, on `st`: bytes[r] from rank r land at recv + off[r]
// (one broadcast per rank inside a group: NCCL fuses them)
int comm_allgatherv_dev(sharp_ctx *c, const void *send, void *recv, const int64_t *bytes, cudaStream_t st) {
    if (!c->comm) return set_error(SHARP_E_ARG, "no communicator on this context (sharp_comm_init)");
    ncclComm_t comm = reinterpret_cast<ncclComm_t>(c->comm);
    int64_t off = 0;
    SHARP_NCCL(g_nccl.GroupStart());
    for (int r = 0; r < c->comm_world; r++) {
        if (bytes[r] > 0) {
            unsigned char *dst = reinterpret_cast<unsigned char *>(recv) + off;
            const void *src = (r == c->comm_rank) ? send : dst;
            ncclResult_t rr = g_nccl.Broadcast(src, dst, (size_t)bytes[r], ncclChar, r, comm, st);
            if (rr != ncclSuccess) {
                g_nccl.GroupEnd();
                return set_error(SHARP_E_CUDA, "ncclBroadcast: %s", g_nccl.GetErrorString(rr));
            }
        }
        off += bytes[r];
    }
    SHARP_NCCL(g_nccl.GroupEnd());
    return 0;
}

void comm_destroy(sharp_ctx *c) {
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(reinterpret_cast<ncclComm_t>(c->comm));
    c->comm = nullptr;
    c->comm_world = 1;
    c->comm_rank = 0;
}

}  // namespace sharp

using namespace sharp;

extern "C" {

int sharp_comm_unique_id(unsigned char *id, int id_len) {
    if (!id || id_len < SHARP_COMM_ID_BYTES) return set_error(SHARP_E_ARG, "comm_unique_id: the buffer must hold %d bytes", SHARP_COMM_ID_BYTES);
    static_assert(sizeof(ncclUniqueId) == SHARP_COMM_ID_BYTES, "ncclUniqueId size");
    SHARP_TRY(nccl_load());
    ncclUniqueId u;
    SHARP_NCCL(g_nccl.GetUniqueId(&u));
    memcpy(id, &u, sizeof u);
    return 0;
}

int sharp_comm_init(sharp_ctx *c, const unsigned char *id, int rank, int world) {
    if (!c || !id || world < 1 || rank < 0 || rank >= world) return set_error(SHARP_E_ARG, "comm_init: bad arguments");
    SHARP_CUDA(cudaSetDevice(c->device));
    if (c->comm) comm_destroy(c);
    if (world == 1) return 0;
    SHARP_TRY(nccl_load());
    ncclUniqueId u;
    memcpy(&u, id, sizeof u);
    ncclComm_t comm = nullptr;
    SHARP_NCCL(g_nccl.CommInitRank(&comm, world, u, rank));
    c->comm = comm;
    c->comm_rank = rank;
    c->comm_world = world;
    return 0;
}

int sharp_comm_destroy(sharp_ctx *c) {
    if (!c) return set_error(SHARP_E_ARG, "null context");
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    comm_destroy(c);
    return 0;
}

int sharp_comm_info(sharp_ctx *c, int *rank, int *world, int *nccl_version) {
    if (!c) return set_error(SHARP_E_ARG, "null context");
    if (rank) *rank = c->comm_rank;
    if (world) *world = c->comm_world;
    if (nccl_version) {
        *nccl_version = 0;
        if (g_nccl.GetVersion) g_nccl.GetVersion(nccl_version);
    }
    return 0;
}

// host buffers: bytes[r] bytes from rank r; recv receives the concatenation in rank order (every rank)
int sharp_comm_allgatherv(sharp_ctx *c, const void *send, const int64_t *bytes, void *recv) {
    if (!c || !bytes || !recv) return set_error(SHARP_E_ARG, "comm_allgatherv: bad arguments");
    if (c->comm_world == 1) {
        if (bytes[0] > 0) memcpy(recv, send, (size_t)bytes[0]);
        return 0;
    }
    SHARP_CUDA(cudaSetDevice(c->device));
    int64_t total = 0;
    for (int r = 0; r < c->comm_world; r++) {
        if (bytes[r] < 0) return set_error(SHARP_E_ARG, "comm_allgatherv: negative size");
        total += bytes[r];
    }
    const int64_t mine = bytes[c->comm_rank];
    if (mine > 0 && !send) return set_error(SHARP_E_ARG, "comm_allgatherv: null send buffer");
    const int64_t total4 = (total + 3) & ~(int64_t)3, mine4 = (mine + 3) & ~(int64_t)3;
    SHARP_TRY(c->ws[WS_COMM_SLOT].reserve((size_t)std::max<int64_t>(total4 + mine4, 16)));
    unsigned char *dev = c->ws[WS_COMM_SLOT].as<unsigned char>();
    unsigned char *dsend = dev + total4;
    if (mine > 0) {
        /* host -> device through the page-locked arena and a copy KERNEL (up to 8 MB): the copy engine may be busy with a
           look-ahead upload of expression data queued earlier, and would make every rank wait for it */
        if (mine4 <= ((int64_t)8 << 20) && c->reserve_pinned((size_t)mine4) == 0) {
            memcpy(c->pinned, send, (size_t)mine);
            SHARP_TRY(h2d_by_kernel(c->stream, dsend, c->pinned, (size_t)mine4));
        } else SHARP_CUDA(cudaMemcpyAsync(dsend, send, (size_t)mine, cudaMemcpyHostToDevice, c->stream));
    }
    SHARP_TRY(comm_allgatherv_dev(c, dsend, dev, bytes, c->stream));
    if (total > 0) SHARP_CUDA(cudaMemcpyAsync(recv, dev, (size_t)total, cudaMemcpyDeviceToHost, c->stream));
    SHARP_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int sharp_comm_bcast(sharp_ctx *c, void *buf, int64_t bytes, int root) {
    if (!c || bytes < 0 || (bytes > 0 && !buf)) return set_error(SHARP_E_ARG, "comm_bcast: bad arguments");
    if (c->comm_world == 1 || bytes == 0) return 0;
    if (root < 0 || root >= c->comm_world) return set_error(SHARP_E_ARG, "comm_bcast: bad root");
    SHARP_CUDA(cudaSetDevice(c->device));
    SHARP_TRY(c->ws[WS_COMM_SLOT].reserve((size_t)bytes));
    void *dev = c->ws[WS_COMM_SLOT].ptr;
    if (c->comm_rank == root) SHARP_CUDA(cudaMemcpyAsync(dev, buf, (size_t)bytes, cudaMemcpyHostToDevice, c->stream));
    SHARP_NCCL(g_nccl.Broadcast(dev, dev, (size_t)bytes, ncclChar, root, reinterpret_cast<ncclComm_t>(c->comm), c->stream));
    if (c->comm_rank != root) SHARP_CUDA(cudaMemcpyAsync(buf, dev, (size_t)bytes, cudaMemcpyDeviceToHost, c->stream));
    SHARP_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int sharp_comm_barrier(sharp_ctx *c) {
    if (!c) return set_error(SHARP_E_ARG, "null context");
    if (c->comm_world == 1) return 0;
    unsigned char token[8] = {0};
    std::vector<int64_t> bytes((size_t)c->comm_world, 8);
    std::vector<unsigned char> all((size_t)c->comm_world * 8);
    return sharp_comm_allgatherv(c, token, bytes.data(), all.data());
}

}  // extern "C"
