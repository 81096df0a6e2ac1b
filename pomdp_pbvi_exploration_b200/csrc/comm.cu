// Collectives of the belief-sharded backup inside the C ABI (pbvi_comm_init / pbvi_allgather_*): a host in any
// language can drive the multi-GPU path -- one process (or thread) per GPU, each with its own pbvi_model handle and pbvi_comm.
// NCCL (over NVLink 5 / NVSwitch inside a box) is bound at run time with dlopen, so the library has no link-time dependency on it:
// single-GPU users never need NCCL, and a process that already loaded NCCL (PyTorch does) shares that copy.
//   exchange step of the sharded backup (north_star item 4):  pbvi_allgather_tuples  -> pbvi_group_record_blocks -> pbvi_backup_assemble
//   literal form (rows travel):                               pbvi_allgather_rows
//   compute_change (src/pomdp.py:2141-2169) across shards:    pbvi_allreduce_max
//   expansion on rank 0 (src/pomdp.py:2306-2320):             pbvi_broadcast_rows
#include <dlfcn.h>

#include <cstdlib>
#include <mutex>

#include "pbvi_common.cuh"

namespace {

// the slice of nccl.h this file needs (NCCL 2.x ABI: the enum values and the 128-byte unique id are stable across 2.x)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
constexpr int NCCL_INT32 = 2, NCCL_FLOAT64 = 8, NCCL_MAX = 2;

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi g_nccl;
std::mutex g_nccl_mutex;        // one thread per GPU is a supported way to drive the library: the binding is resolved once

int load_nccl() {
    std::lock_guard<std::mutex> lock(g_nccl_mutex);
    if (g_nccl.lib) return PBVI_OK;
    const char* env = std::getenv("PBVI_NCCL_LIB");
    const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
    void* lib = nullptr;
    for (const char* n : names) {
        if (!n) continue;
        lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (lib) break;
    }
    if (!lib) {
        pbvi::set_error("NCCL is not available: dlopen(libnccl.so.2) failed (%s); set PBVI_NCCL_LIB to its path", dlerror());
        return PBVI_ERR_UNSUPPORTED;
    }
    NcclApi api;
    api.lib = lib;
    bool ok = true;
    auto sym = [&](const char* name) { void* p = dlsym(lib, name); ok &= p != nullptr; return p; };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
    api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(sym("ncclBroadcast"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    if (!ok) {
        pbvi::set_error("the NCCL library lacks a required symbol");
        return PBVI_ERR_UNSUPPORTED;
    }
    g_nccl = api;
    return PBVI_OK;
}

}  // namespace

struct pbvi_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1, device = 0;
};

#define PBVI_NCCL(call)                                                                                   \
    do {                                                                                                  \
        ncclResult_t _r = (call);                                                                         \
        if (_r != 0) {                                                                                    \
            pbvi::set_error("%s failed: %s (%s:%d)", #call, g_nccl.GetErrorString(_r), __FILE__, __LINE__); \
            return PBVI_ERR_NCCL;                                                                         \
        }                                                                                                 \
    } while (0)

extern "C" int pbvi_comm_unique_id(unsigned char* id128) {
    PBVI_REQUIRE(id128 != nullptr, "id buffer is NULL");
    PBVI_TRY(load_nccl());
    ncclUniqueId id;
    PBVI_NCCL(g_nccl.GetUniqueId(&id));
    std::memcpy(id128, id.internal, 128);
    return PBVI_OK;
}

extern "C" int pbvi_comm_init(pbvi_model* m, const unsigned char* id128, int rank, int nranks, pbvi_comm** out) {
    PBVI_REQUIRE(out != nullptr, "out pointer is NULL");
    *out = nullptr;
    PBVI_REQUIRE(m != nullptr && id128 != nullptr, "NULL argument");
    PBVI_REQUIRE(nranks > 0 && rank >= 0 && rank < nranks, "need 0 <= rank < nranks");
    PBVI_TRY(load_nccl());
    PBVI_CUDA(cudaSetDevice(m->device));
    ncclUniqueId id;
    std::memcpy(id.internal, id128, 128);
    pbvi_comm* c = new pbvi_comm();
    c->rank = rank; c->nranks = nranks; c->device = m->device;
    ncclResult_t r = g_nccl.CommInitRank(&c->comm, nranks, id, rank);
    if (r != 0) {
        pbvi::set_error("ncclCommInitRank failed: %s", g_nccl.GetErrorString(r));
        delete c;
        return PBVI_ERR_NCCL;
    }
    *out = c;
    return PBVI_OK;
}

extern "C" int pbvi_comm_destroy(pbvi_comm* c) {
    if (!c) return PBVI_OK;
    if (c->comm && g_nccl.CommDestroy) {
        cudaSetDevice(c->device);
        g_nccl.CommDestroy(c->comm);
    }
    delete c;
    return PBVI_OK;
}

extern "C" int pbvi_comm_rank(const pbvi_comm* c, int* rank, int* nranks) {
    PBVI_REQUIRE(c != nullptr, "comm handle is NULL");
    if (rank) *rank = c->rank;
    if (nranks) *nranks = c->nranks;
    return PBVI_OK;
}

extern "C" int pbvi_allgather_tuples(pbvi_comm* c, const int32_t* d_block, int block_rows, int row_words, int32_t* d_gathered, void* stream) {
    PBVI_REQUIRE(c != nullptr && d_block && d_gathered, "NULL argument");
    PBVI_REQUIRE(block_rows > 0 && row_words > 0, "block shape must be positive");
    PBVI_CUDA(cudaSetDevice(c->device));
    PBVI_NCCL(g_nccl.AllGather(d_block, d_gathered, (size_t)block_rows * row_words, NCCL_INT32, c->comm, (cudaStream_t)stream));
    return PBVI_OK;
}

extern "C" int pbvi_allgather_rows(pbvi_comm* c, const double* d_rows, int n_rows, int row_len, double* d_gathered, void* stream) {
    PBVI_REQUIRE(c != nullptr && d_rows && d_gathered, "NULL argument");
    PBVI_REQUIRE(n_rows > 0 && row_len > 0, "block shape must be positive");
    PBVI_CUDA(cudaSetDevice(c->device));
    PBVI_NCCL(g_nccl.AllGather(d_rows, d_gathered, (size_t)n_rows * row_len, NCCL_FLOAT64, c->comm, (cudaStream_t)stream));
    return PBVI_OK;
}

extern "C" int pbvi_allreduce_max(pbvi_comm* c, double* d_values, int n, void* stream) {
    PBVI_REQUIRE(c != nullptr && d_values, "NULL argument");
    PBVI_REQUIRE(n > 0, "n must be positive");
    PBVI_CUDA(cudaSetDevice(c->device));
    PBVI_NCCL(g_nccl.AllReduce(d_values, d_values, (size_t)n, NCCL_FLOAT64, NCCL_MAX, c->comm, (cudaStream_t)stream));
    return PBVI_OK;
}

extern "C" int pbvi_broadcast_rows(pbvi_comm* c, double* d_rows, size_t count, int root, void* stream) {
    PBVI_REQUIRE(c != nullptr && d_rows, "NULL argument");
    PBVI_REQUIRE(root >= 0 && root < c->nranks, "no such root rank");
    if (count == 0) return PBVI_OK;
    PBVI_CUDA(cudaSetDevice(c->device));
    PBVI_NCCL(g_nccl.Broadcast(d_rows, d_rows, count, NCCL_FLOAT64, root, c->comm, (cudaStream_t)stream));
    return PBVI_OK;
}
