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
#include "ec_reduce.cuh"

struct ec_comm {
    ncclComm_t comm;
    int n_ranks, rank;
    int64_t* dkeys;   // 4 device words for in-place all-reduce (+ 20 for the statistics limbs)
    int64_t* pinned;  // host mirror
    // NVLink peer exchange (see PeerExchange in ec_reduce.cuh): every rank's mailbox mapped into every rank
    bool peer_ok;
    unsigned long long* mailbox;            // ours: 2 epochs x n_ranks slots x 4 words
    unsigned long long** peer_ptrs_dev;     // device array [n_ranks]
    void* opened[64];                       // cudaIpcOpenMemHandle mappings to close
    unsigned long long epoch;
};

namespace ec {
struct Nccl {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
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
    SYM(CommInitAll, "ncclCommInitAll")
    SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(AllReduce, "ncclAllReduce")
    SYM(AllGather, "ncclAllGather")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    g_nccl.h = h;
    return EC_OK;
}
static ec_status nccl_fail(ncclResult_t r, const char* what) {
    set_error("%s: %s", what, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "nccl error");
    return EC_NCCL;
}

// One process driving several GPUs (ec_init_devices): communicators over its own devices, made on first use.
static ncclComm_t g_local_comms[kMaxDev];
static int g_local_n = 0;
ec_status local_allreduce_min_i64(const int* cuda_devices, int n, int64_t* const* device_bufs, const cudaStream_t* streams, size_t count) {
    if (ec_status s = nccl_load()) return s;
    {
        std::lock_guard<std::mutex> lk(g_nccl_mu);
        if (g_local_n == 0) {
            if (ncclResult_t r = g_nccl.CommInitAll(g_local_comms, n, cuda_devices)) return nccl_fail(r, "ncclCommInitAll");
            g_local_n = n;
        }
    }
    if (n != g_local_n) { set_error("invalid argument: the local communicators span %d devices", g_local_n); return EC_INVALID_ARG; }
    if (ncclResult_t r = g_nccl.GroupStart()) return nccl_fail(r, "ncclGroupStart");
    for (int g = 0; g < n; ++g)
        if (ncclResult_t r = g_nccl.AllReduce(device_bufs[g], device_bufs[g], count, ncclInt64, ncclMin, g_local_comms[g], streams[g])) {
            g_nccl.GroupEnd();
            return nccl_fail(r, "ncclAllReduce(min)");
        }
    if (ncclResult_t r = g_nccl.GroupEnd()) return nccl_fail(r, "ncclGroupEnd");
    return EC_OK;
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
// Map every rank's mailbox into this process: cudaIpc handles travel through an NCCL all-gather. Every rank runs the
// same collectives whatever happens locally — a rank that could not allocate, export or map sends a zeroed handle and
// votes 0 in the final MIN all-reduce, so all ranks agree on peer_ok and nobody is left waiting inside a collective.
static ec_status peer_setup(ec_comm* c) {
    c->peer_ok = false;
    if (c->n_ranks > 32 || env_int("EC_NO_PEER_EXCHANGE", 0)) return EC_OK;  // one warp folds the ranks; NCCL path otherwise (same decision on every rank)
    cudaStream_t st = static_cast<cudaStream_t>(ec_get_stream());
    // two regions: the 4-word min_max / counts messages, then the 20-word statistics-sums messages (2 epochs x n_ranks slots each)
    const size_t words = size_t(2) * c->n_ranks * 4 + size_t(2) * c->n_ranks * 2 * kSumWords;
    int ok = 1;
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof mine);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    if (cudaMalloc(reinterpret_cast<void**>(&c->mailbox), words * 8) != cudaSuccess || cudaMemset(c->mailbox, 0, words * 8) != cudaSuccess ||
        cudaIpcGetMemHandle(&mine, c->mailbox) != cudaSuccess) {
        cudaGetLastError();
        ok = 0;
        memset(&mine, 0, sizeof mine);
    }
    char* dh = nullptr;
    if (cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&dh), size_t(64) * (c->n_ranks + 1))) return cuda_fail(e, "cudaMalloc");
    struct Free { char* p; ~Free() { cudaFree(p); } } free_dh{dh};
    if (cudaError_t e = cudaMemcpyAsync(dh, &mine, 64, cudaMemcpyHostToDevice, st)) return cuda_fail(e, "cudaMemcpyAsync");
    if (ncclResult_t r = g_nccl.AllGather(dh, dh + 64, 64, ncclChar, c->comm, st)) return nccl_fail(r, "ncclAllGather(ipc handles)");
    cudaIpcMemHandle_t all[64];
    if (cudaError_t e = cudaMemcpyAsync(all, dh + 64, size_t(64) * c->n_ranks, cudaMemcpyDeviceToHost, st)) return cuda_fail(e, "cudaMemcpyAsync");
    if (cudaError_t e = cudaStreamSynchronize(st)) return cuda_fail(e, "cudaStreamSynchronize");
    unsigned long long* ptrs[64];
    const cudaIpcMemHandle_t zero{};
    for (int r = 0; r < c->n_ranks; ++r) {
        c->opened[r] = nullptr;
        ptrs[r] = nullptr;
        if (r == c->rank) { ptrs[r] = c->mailbox; continue; }
        if (!ok) continue;
        void* p = nullptr;
        if (memcmp(&all[r], &zero, sizeof zero) == 0 || cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; continue; }
        c->opened[r] = p;
        ptrs[r] = static_cast<unsigned long long*>(p);
    }
    // every rank must agree before anyone uses the exchange (a rank that could not map a peer vetoes it)
    c->pinned[0] = ok;
    if (cudaError_t e = cudaMemcpyAsync(c->dkeys, c->pinned, 8, cudaMemcpyHostToDevice, st)) return cuda_fail(e, "cudaMemcpyAsync");
    if (ncclResult_t r = g_nccl.AllReduce(c->dkeys, c->dkeys, 1, ncclInt64, ncclMin, c->comm, st)) return nccl_fail(r, "ncclAllReduce");
    if (cudaError_t e = cudaMemcpyAsync(c->pinned, c->dkeys, 8, cudaMemcpyDeviceToHost, st)) return cuda_fail(e, "cudaMemcpyAsync");
    if (cudaError_t e = cudaStreamSynchronize(st)) return cuda_fail(e, "cudaStreamSynchronize");
    if (c->pinned[0] != 1) {  // vetoed somewhere: give the mailbox and the mappings back, the NCCL path is used
        for (int r = 0; r < c->n_ranks; ++r)
            if (c->opened[r]) { cudaIpcCloseMemHandle(c->opened[r]); c->opened[r] = nullptr; }
        cudaFree(c->mailbox);
        c->mailbox = nullptr;
        return EC_OK;
    }
    if (cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&c->peer_ptrs_dev), sizeof(void*) * c->n_ranks)) return cuda_fail(e, "cudaMalloc");
    if (cudaError_t e = cudaMemcpy(c->peer_ptrs_dev, ptrs, sizeof(void*) * c->n_ranks, cudaMemcpyHostToDevice)) return cuda_fail(e, "cudaMemcpy");
    c->peer_ok = true;
    return EC_OK;
}
static PeerExchange next_exchange(ec_comm* c) {
    PeerExchange px;
    px.peers = c->peer_ptrs_dev;
    px.n_ranks = c->n_ranks;
    px.rank = c->rank;
    px.epoch = ++c->epoch;
    // A peer that never arrives (a rank skipped the collective, or died) must become an error, not a hung GPU:
    // give up after EC_PEER_WAIT_SECONDS (default 20) of SM clock ticks.
    static const unsigned long long limit = [] {
        int khz = 1900000, dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
        return static_cast<unsigned long long>(env_int("EC_PEER_WAIT_SECONDS", 20)) * 1000ull * static_cast<unsigned long long>(khz);
    }();
    px.spin_limit = limit;
    return px;
}

ec_status ec_comm_init_rank(const void* id128, int n_ranks, int rank, ec_comm** out) {
    if (ec_status s = nccl_load()) return s;
    if (ec_status s = ec_synchronize()) return s;  // binds the device
    if (n_ranks < 1 || n_ranks > 64 || rank < 0 || rank >= n_ranks) { set_error("invalid argument: rank / n_ranks"); return EC_INVALID_ARG; }
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ec_comm* c = new ec_comm{};
    c->n_ranks = n_ranks;
    c->rank = rank;
    if (ncclResult_t r = g_nccl.CommInitRank(&c->comm, n_ranks, id, rank)) { delete c; return nccl_fail(r, "ncclCommInitRank"); }
    if (cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&c->dkeys), 24 * sizeof(int64_t))) { delete c; return cuda_fail(e, "cudaMalloc"); }
    if (cudaError_t e = cudaMallocHost(reinterpret_cast<void**>(&c->pinned), 24 * sizeof(int64_t))) { delete c; return cuda_fail(e, "cudaMallocHost"); }
    if (ec_status s = peer_setup(c)) { ec_comm_destroy(c); return s; }
    *out = c;
    return EC_OK;
}
int ec_comm_peer_exchange(const ec_comm* c) { return c->peer_ok ? 1 : 0; }
void ec_comm_destroy(ec_comm* c) {
    if (!c) return;
    for (int r = 0; r < c->n_ranks && r < 64; ++r)
        if (c->opened[r]) cudaIpcCloseMemHandle(c->opened[r]);
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    cudaFree(c->peer_ptrs_dev);
    cudaFree(c->mailbox);
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
    if (c->peer_ok) {  // ONE kernel per GPU: shard reduction + exchange over NVLink peer memory + final fold
        uint64_t k[2];
        if (ec_status s = reduce_min_max_peer(shard, mask_or_null, next_exchange(c), &k[0], &k[1])) return s;
        const uint8_t ct = ec_buf_ctype(shard);
        memset(mn, 0, sizeof *mn); memset(mx, 0, sizeof *mx);
        mn->ct = mx->ct = ct;
        mn->bits = key_to_bits(ct, k[0]);
        mx->bits = key_to_bits(ct, k[1]);
        return EC_OK;
    }
    if (ec_status s = ec_buf_min_max_keys(shard, mask_or_null, c->dkeys)) return s;
    if (ec_status s = ec_comm_allreduce_min_i64(c, c->dkeys, 2)) return s;
    cudaStream_t st = static_cast<cudaStream_t>(ec_get_stream());
    if (cudaError_t e = cudaMemcpyAsync(c->pinned, c->dkeys, 16, cudaMemcpyDeviceToHost, st)) return cuda_fail(e, "cudaMemcpyAsync(D2H)");
    if (cudaError_t e = cudaStreamSynchronize(st)) return cuda_fail(e, "cudaStreamSynchronize");
    return ec_min_max_from_keys(ec_buf_ctype(shard), c->pinned, mn, mx);
}
ec_status ec_mask_counts_sharded(ec_comm* c, const ec_mask* shard, size_t* data, size_t* nodata) {
    if (c->peer_ok) {
        uint64_t ones, total;
        if (ec_status s = reduce_popcount_peer(shard, next_exchange(c), &ones, &total)) return s;
        *data = ones;
        *nodata = total - ones;
        return EC_OK;
    }
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
// Statistics of a row-strip sharded raster (extension): global min/max (above) -> the same plan on
// every rank -> this strip's exact moment sums -> one all-reduce(SUM) -> the same finish everywhere. The 128-bit
// sums travel as 32-bit limbs in 64-bit words, so adding up to 64 ranks cannot lose a carry.
ec_status ec_buf_statistics_sharded(ec_comm* c, const ec_buf* shard, const ec_mask* mask_or_null, ec_statistics* out) {
    ec_value mn, mx;
    if (ec_status s = ec_buf_min_max_sharded(c, shard, mask_or_null, &mn, &mx)) return s;
    int kind, e;
    double p;
    if (ec_status s = ec_statistics_plan(&mn, &mx, &kind, &p, &e)) return s;
    uint64_t raw[EC_MOMENT_WORDS] = {0};
    if (kind == EC_STATS_REGULAR && c->peer_ok) {
        // the sums ride the peer mailboxes like the min_max keys: statistics kernel + one exchanging CTA behind it, no NCCL
        uint64_t total[EC_MOMENT_WORDS];
        if (ec_status s = moments_exchange(shard, mask_or_null, p, e, next_exchange(c), size_t(2) * c->n_ranks * 4, total)) return s;
        return ec_statistics_finish(total, 1, &mn, &mx, out);
    }
    if (kind == EC_STATS_REGULAR) {
        if (ec_status s = ec_buf_moments(shard, mask_or_null, p, e, raw)) return s;
    } else if (mask_or_null) {
        size_t d = 0, nd = 0;
        if (ec_status s = ec_mask_counts(mask_or_null, &d, &nd)) return s;
        raw[0] = d;
    } else {
        raw[0] = ec_buf_len(shard);
    }
    uint64_t* limbs = reinterpret_cast<uint64_t*>(c->pinned + 4);
    uint64_t* dlimbs = reinterpret_cast<uint64_t*>(c->dkeys + 4);
    limbs[0] = raw[0];
    for (int k = 0; k < 8; ++k) { limbs[1 + 2 * k] = raw[1 + k] & 0xFFFFFFFFull; limbs[2 + 2 * k] = raw[1 + k] >> 32; }
    cudaStream_t st = static_cast<cudaStream_t>(ec_get_stream());
    if (cudaError_t err = cudaMemcpyAsync(dlimbs, limbs, 17 * sizeof(uint64_t), cudaMemcpyHostToDevice, st)) return cuda_fail(err, "cudaMemcpyAsync(H2D)");
    if (ec_status s = ec_comm_allreduce_sum_u64(c, dlimbs, 17)) return s;
    if (cudaError_t err = cudaMemcpyAsync(limbs, dlimbs, 17 * sizeof(uint64_t), cudaMemcpyDeviceToHost, st)) return cuda_fail(err, "cudaMemcpyAsync(D2H)");
    if (cudaError_t err = cudaStreamSynchronize(st)) return cuda_fail(err, "cudaStreamSynchronize");
    uint64_t total[EC_MOMENT_WORDS];
    total[0] = limbs[0];
    for (int a = 0; a < 4; ++a) {  // four 128-bit accumulators, each from four summed limbs, modulo 2^128
        unsigned __int128 v = 0;
        for (int j = 3; j >= 0; --j) v = (v << 32) + limbs[1 + 4 * a + j];
        total[1 + 2 * a] = static_cast<uint64_t>(v);
        total[2 + 2 * a] = static_cast<uint64_t>(v >> 64);
    }
    return ec_statistics_finish(total, 1, &mn, &mx, out);
}

}  // extern "C"
