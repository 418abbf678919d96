// Row-strip sharding across the GPUs of one box: the only cross-GPU step on this path is finishing a
// reduction (min_max keys, mask counts) — at most 16 bytes per rank — with one NCCL all-reduce over
// NVLink/NVSwitch. One rank per process; libnccl is dlopen'ed on first use so single-GPU callers carry
// no NCCL dependency (in a torch process this resolves to the NCCL torch already loaded).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>

#include "ec_internal.hpp"

struct ec_comm {
    ncclComm_t comm;
    int n_ranks, rank;
    int64_t* dkeys;   // 2 device words for in-place all-reduce
    int64_t* pinned;  // host mirror
};

namespace ec {
struct Nccl {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static Nccl g_nccl;
static std::mutex g_nccl_mu;

static ec_status nccl_load() {
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (g_nccl.h) return EC_OK;
    const char* names[] = {getenv("EC_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names) {
        if (n && *n && (h = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
    }
    if (!h) { set_error("cannot load NCCL: %s", dlerror()); return EC_NCCL; }
#define SYM(field, name)                                                         \
    g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(h, name));     \
    if (!g_nccl.field) { set_error("NCCL symbol %s missing", name); return EC_NCCL; }
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(AllReduce, "ncclAllReduce")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    g_nccl.h = h;
    return EC_OK;
}
static ec_status nccl_fail(ncclResult_t r, const char* what) {
    set_error("%s: %s", what, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "nccl error");
    return EC_NCCL;
}
}  // namespace ec

using namespace ec;

extern "C" {

ec_status ec_comm_unique_id(void* id128) {
    if (ec_status s = nccl_load()) return s;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    if (ncclResult_t r = g_nccl.GetUniqueId(&id)) return nccl_fail(r, "ncclGetUniqueId");
    memcpy(id128, &id, 128);
    return EC_OK;
}
ec_status ec_comm_init_rank(const void* id128, int n_ranks, int rank, ec_comm** out) {
    if (ec_status s = nccl_load()) return s;
    if (ec_status s = ec_synchronize()) return s;  // binds the device
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ec_comm* c = new ec_comm{};
    c->n_ranks = n_ranks;
    c->rank = rank;
    if (ncclResult_t r = g_nccl.CommInitRank(&c->comm, n_ranks, id, rank)) { delete c; return nccl_fail(r, "ncclCommInitRank"); }
    if (cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&c->dkeys), 4 * sizeof(int64_t))) { delete c; return cuda_fail(e, "cudaMalloc"); }
    if (cudaError_t e = cudaMallocHost(reinterpret_cast<void**>(&c->pinned), 4 * sizeof(int64_t))) { delete c; return cuda_fail(e, "cudaMallocHost"); }
    *out = c;
    return EC_OK;
}
void ec_comm_destroy(ec_comm* c) {
    if (!c) return;
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    cudaFree(c->dkeys);
    cudaFreeHost(c->pinned);
    delete c;
}
ec_status ec_comm_allreduce_min_i64(ec_comm* c, int64_t* device_buf, size_t count) {
    if (ncclResult_t r = g_nccl.AllReduce(device_buf, device_buf, count, ncclInt64, ncclMin, c->comm, static_cast<cudaStream_t>(ec_get_stream())))
        return nccl_fail(r, "ncclAllReduce(min)");
    return EC_OK;
}
ec_status ec_comm_allreduce_sum_u64(ec_comm* c, uint64_t* device_buf, size_t count) {
    if (ncclResult_t r = g_nccl.AllReduce(device_buf, device_buf, count, ncclUint64, ncclSum, c->comm, static_cast<cudaStream_t>(ec_get_stream())))
        return nccl_fail(r, "ncclAllReduce(sum)");
    return EC_OK;
}
ec_status ec_buf_min_max_sharded(ec_comm* c, const ec_buf* shard, const ec_mask* mask_or_null, ec_value* mn, ec_value* mx) {
    if (ec_status s = ec_buf_min_max_keys(shard, mask_or_null, c->dkeys)) return s;
    if (ec_status s = ec_comm_allreduce_min_i64(c, c->dkeys, 2)) return s;
    cudaStream_t st = static_cast<cudaStream_t>(ec_get_stream());
    if (cudaError_t e = cudaMemcpyAsync(c->pinned, c->dkeys, 16, cudaMemcpyDeviceToHost, st)) return cuda_fail(e, "cudaMemcpyAsync(D2H)");
    if (cudaError_t e = cudaStreamSynchronize(st)) return cuda_fail(e, "cudaStreamSynchronize");
    return ec_min_max_from_keys(ec_buf_ctype(shard), c->pinned, mn, mx);
}
ec_status ec_mask_counts_sharded(ec_comm* c, const ec_mask* shard, size_t* data, size_t* nodata) {
    size_t d = 0, nd = 0;
    if (ec_status s = ec_mask_counts(shard, &d, &nd)) return s;
    cudaStream_t st = static_cast<cudaStream_t>(ec_get_stream());
    c->pinned[2] = static_cast<int64_t>(d);
    c->pinned[3] = static_cast<int64_t>(nd);
    if (cudaError_t e = cudaMemcpyAsync(c->dkeys + 2, c->pinned + 2, 16, cudaMemcpyHostToDevice, st)) return cuda_fail(e, "cudaMemcpyAsync(H2D)");
    if (ec_status s = ec_comm_allreduce_sum_u64(c, reinterpret_cast<uint64_t*>(c->dkeys + 2), 2)) return s;
    if (cudaError_t e = cudaMemcpyAsync(c->pinned + 2, c->dkeys + 2, 16, cudaMemcpyDeviceToHost, st)) return cuda_fail(e, "cudaMemcpyAsync(D2H)");
    if (cudaError_t e = cudaStreamSynchronize(st)) return cuda_fail(e, "cudaStreamSynchronize");
    *data = static_cast<size_t>(c->pinned[2]);
    *nodata = static_cast<size_t>(c->pinned[3]);
    return EC_OK;
}

}  // extern "C"
