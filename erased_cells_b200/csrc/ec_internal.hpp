// Host-side internals shared by the translation units of liberased_cells_b200.so.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <memory>
#include <string>
#include <tuple>
#include <type_traits>
#include <utility>

#include "../../include/erased_cells_b200.h"

#include <algorithm>
#include <atomic>
#include <functional>
#include <vector>

namespace ec {
struct Expr;       // a deferred op (lazy mode): evaluated, possibly fused with its children, on first use
struct CountSlot;  // pinned host slot a mask-producing kernel publishes its popcount into

// Refcounted device allocation. It belongs to one logical device and to the stream it was allocated for (its home
// stream): kernels that read or write it on another stream are ordered against the home stream by events, and the
// block goes back to the home stream's free list when its last user lets go (see dev_free in ec_api.cu).
struct DevBlock {
    void* p;
    int dev;
    cudaStream_t home;
    std::atomic<int> snapshots{0};       // pending lazy expressions that read this block: in-place mutation copies first
    std::atomic<int> views{0};           // live views of the block: a put through one is seen by all, so nothing about the cells is cached meanwhile
    std::vector<std::pair<cudaStream_t, int>> foreign;  // other (stream, logical device) that touched the block (guarded by the allocator mutex)
    DevBlock(void* q, int d, cudaStream_t s) : p(q), dev(d), home(s) {}
    DevBlock(const DevBlock&) = delete;
    DevBlock& operator=(const DevBlock&) = delete;
    ~DevBlock();
};
}  // namespace ec

// A CellBuffer is either plain — `len` cells at dptr on logical device `dev` — or row-strip sharded: parts[g] is a
// plain buffer on logical device g holding cells [offs[g], offs[g + 1]) (possibly none), dptr is null.
struct ec_buf {
    uint8_t ct = 0;
    bool owned = true;
    size_t len = 0;
    size_t capacity_bytes = 0;
    void* dptr = nullptr;         // null while `expr` is pending (and for a sharded buffer)
    cudaEvent_t ready = nullptr;  // set by ec_buf_from_host_async: readers on other streams wait on it
    std::shared_ptr<ec::DevBlock> blk;
    std::shared_ptr<ec::Expr> expr;
    int dev = 0;
    std::vector<ec_buf*> parts;
    std::vector<size_t> offs;
    // min_max of all cells (order keys), remembered once any unmasked min_max has run and dropped by put / extend: a later
    // min_max is free, and statistics of Float32 / 64-bit cells need no min_max pass of their own
    bool mm_known = false;
    uint64_t mm_k0 = 0, mm_k1 = 0;
    bool is_view = false;
};
// Validity bits, packed; plain or sharded like ec_buf. The words are refcounted (a clone shares them until one side is
// mutated). The number of set bits is cached: kernels that produce a mask count as they go and the last CTA
// publishes the total into `cnt_slot` (pinned host memory) tagged with `cnt_seq`.
struct ec_mask {
    size_t len = 0;
    size_t capacity_bytes = 0;
    uint32_t* words = nullptr;
    std::shared_ptr<ec::DevBlock> blk;
    int dev = 0;
    std::shared_ptr<ec::CountSlot> cnt;  // {ones, seq} in pinned host memory; shared by clones
    uint64_t cnt_seq = 0;                // the tag the producing kernel publishes with (0: nobody is counting)
    bool cnt_known = false;
    uint64_t ones = 0;
    std::vector<ec_mask*> parts;
    std::vector<size_t> offs;
};
struct ec_event {
    cudaEvent_t ev;
};

namespace ec {

struct ReduceScratch;
struct PeerExchange;
struct MaskCount;  // ec_common.cuh: where a mask-producing kernel publishes the number of set bits it wrote

constexpr int kMaxDev = 16;
enum : int { FINISH_HOST = 0, FINISH_PEER = 1, FINISH_NCCL = 2 };  // how a sharded reduction crosses the GPUs (ec_set_shard_finish)

// ---- logical devices and row-strip sharding inside one process (ec_api.cu / ec_shard.cu) ------------------------
int n_devices();
int device_phys(int dev);
cudaStream_t device_stream(int dev);  // the stream the calling thread's work on logical device `dev` goes to
// true when a fresh buffer of n cells is to be spread over the devices: several devices, n >= the threshold, and the
// caller is not already working on one strip
bool shard_policy(size_t n);
inline bool is_sharded(const ec_buf* b) { return !b->parts.empty(); }
inline bool is_sharded(const ec_mask* m) { return !m->parts.empty(); }
// run plain entry points on logical device `dev` (binds the thread to its CUDA device, turns the sharding policy off)
struct DevScope {
    int prev_dev, prev_phys = -1;
    explicit DevScope(int dev);
    ~DevScope();
    DevScope(const DevScope&) = delete;
    DevScope& operator=(const DevScope&) = delete;
};
// a reduction in flight whose finishing CTA publishes {r0, r1, seq, status} into mapped pinned host memory
struct PendingReduce {
    volatile uint64_t* pin = nullptr;
    uint64_t seq = 0;
    cudaStream_t stream = nullptr;
    int dev = 0;
};
ec_status reduce_end(const PendingReduce& p, uint64_t* r0, uint64_t* r1);
// statistics kernels in flight: raw accumulators arrive in pinned host memory once `stream` has drained
struct StatsPending {
    uint64_t* pin = nullptr;
    cudaStream_t stream = nullptr;
    int dev = 0;
    int words = 0;
};
// begin = enqueue on the calling thread's current device and return; end = wait for the result
ec_status min_max_begin(const ec_buf* b, const ec_mask* m, const PeerExchange* px, PendingReduce* pend, ReduceScratch* sc_out);
ec_status popcount_begin(const ec_mask* m, const PeerExchange* px, uint64_t second_word, PendingReduce* pend);
ec_status first_diff_begin(const ec_buf* l, const ec_buf* r, size_t n, PendingReduce* pend);
ec_status mask_first_diff(const ec_mask* l, const ec_mask* r, size_t n, uint64_t* bit_out);
ec_status int_stats_begin(const ec_buf* b, const ec_mask* m, StatsPending* p);
ec_status quant_stats_begin(const ec_buf* b, const ec_mask* m, int exp2, StatsPending* p);
ec_status moments_begin(const ec_buf* b, const ec_mask* m, double pivot, int exp2, StatsPending* p);
ec_status stats_end(const StatsPending& p, uint64_t* w);
bool stats_integer_route(uint8_t ct);
void int_stats_min_max(uint8_t ct, const uint64_t* w, ec_value* mn, ec_value* mx);
ec_status mask_ones(const ec_mask* m, uint64_t* ones);
// all-reduce(MIN) of `count` int64 per device over the GPUs of this process (ncclCommInitAll communicators, ec_comm.cu)
ec_status local_allreduce_min_i64(const int* cuda_devices, int n, int64_t* const* device_bufs, const cudaStream_t* streams, size_t count);

// ---- sharded flavours of the entry points (ec_shard.inc) ---------------------------------------------------------
using GenFn = std::function<ec_status(size_t off, size_t n, ec_buf** out)>;
using MaskGenFn = std::function<ec_status(size_t off, size_t n, ec_mask** out)>;
using Map1Fn = std::function<ec_status(const ec_buf*, ec_buf**)>;
using Map2Fn = std::function<ec_status(const ec_buf*, const ec_buf*, ec_buf**)>;
using MapBmFn = std::function<ec_status(const ec_buf*, const ec_mask*, ec_buf**)>;
using BufToMaskFn = std::function<ec_status(const ec_buf*, ec_mask**)>;
using MaskMap2Fn = std::function<ec_status(const ec_mask*, const ec_mask* /* may be null */, ec_mask**)>;
int sh_part_of(const std::vector<size_t>& offs, size_t index);
ec_status sh_generate(uint8_t ct, size_t len, const GenFn& fn, ec_buf** out);
ec_status shm_generate(size_t len, const MaskGenFn& fn, ec_mask** out);
// ---- pageable host memory (ec_ingest.inc): copies between a Vec<T> and HBM go through pinned staging chunks, moved
// by a pool of host threads while the DMA of the previous chunk runs
struct HostSeg { int dev; void* d; void* h; size_t bytes; cudaStream_t stream; };
bool staged_wanted(const void* host, size_t bytes);                    // large, and neither pinned nor registered
ec_status staged_transfer(const HostSeg* segs, int n, bool to_device);  // returns with every byte in place
ec_status sh_from_host(uint8_t ct, const void* host, size_t len, bool async, ec_buf** out);
ec_status sh_to_host(const ec_buf* b, void* host);
ec_status sh_view(const ec_buf* b, size_t off, size_t len, ec_buf** out);
ec_status shm_slice(const ec_mask* m, size_t off, size_t len, ec_mask** out);
ec_status sh_map1(const ec_buf* b, uint8_t out_ct, const Map1Fn& fn, ec_buf** out);
ec_status sh_map2(const ec_buf* l, const ec_buf* r, size_t n, uint8_t out_ct, const Map2Fn& fn, ec_buf** out);
ec_status sh_map_bm(const ec_buf* b, const ec_mask* m, uint8_t out_ct, const MapBmFn& fn, ec_buf** out);
ec_status sh_buf_to_mask(const ec_buf* b, const BufToMaskFn& fn, ec_mask** out);
ec_status shm_map2(const ec_mask* l, const ec_mask* r, size_t n, const MaskMap2Fn& fn, ec_mask** out);
ec_status sh_masked_binary(int op, const ec_buf* lb, const ec_mask* lm, const ec_buf* rb, const ec_mask* rm, size_t n,
                           ec_buf** out_buf, ec_mask** out_mask);
ec_status sh_min_max(const ec_buf* b, const ec_mask* m, uint64_t* k0, uint64_t* k1);
ec_status sh_first_diff(const ec_buf* l, const ec_buf* r, size_t n, uint64_t* idx);
ec_status shm_first_diff(const ec_mask* l, const ec_mask* r, size_t n, uint64_t* bit);
ec_status sh_moments(const ec_buf* b, const ec_mask* m, double pivot, int exp2, uint64_t* raw);
ec_status sh_statistics(const ec_buf* b, const ec_mask* m, ec_statistics* out);

struct Launch {  // launch context handed to every launcher
    cudaStream_t stream;
    int sm_count;
    int max_grid;  // cap for the persistent grids (<= 0: one tile per CTA)
    bool overlap;  // programmatic dependent launch: the grid may be scheduled while its predecessor drains
    bool graph = false;  // launch through a cached one-node CUDA graph (reductions: the host waits for their answer)
};
// One-node CUDA graphs, one per (kernel, device), re-parameterised before every launch. On B200 / driver 580 "update the node's
// parameters + cudaGraphLaunch" reaches the GPU 2 us sooner than a stream launch of the same kernel and with a tenth of the
// spread (tools/launch_floor.cu, profiles/r02_launch_floor.txt: 5.8 us median / 5.9 p90 against 7.8 / 9.6 until the host sees
// the kernel's answer) — which is most of what a small reduction costs, and the start skew of a sharded one.
struct GraphSlot {
    cudaGraph_t graph = nullptr;
    cudaGraphNode_t node = nullptr;
    cudaGraphExec_t exec = nullptr;
    std::atomic_flag busy = ATOMIC_FLAG_INIT;  // held from the parameter update to the launch
};
GraphSlot* graph_slot_acquire(const void* func);  // a slot nobody else is updating, for the calling thread's current device
inline void graph_slot_release(GraphSlot* s) { s->busy.clear(std::memory_order_release); }

void set_error(const char* fmt, ...);
ec_status cuda_fail(cudaError_t e, const char* what);
void note_launch(const char* family);
int env_int(const char* name, int dflt);

// grid for a tiled streaming kernel
inline int grid_for(size_t n, size_t tile, const Launch& L) {
    size_t full = n / tile;
    if (full == 0) full = 1;
    size_t cap = L.max_grid > 0 ? size_t(L.max_grid) : (size_t(1) << 30);
    return int(full < cap ? full : cap);
}

// Launch of a kernel whose first statement is overlap_prologue() (ec_common.cuh). With L.overlap the grid carries the
// programmatic-stream-serialization attribute; memory ordering against the previous grid is kept by the prologue.
#ifdef __CUDACC__
template <class... P, size_t... I>
inline void graph_arg_pointers(std::tuple<P...>& params, void** argv, std::index_sequence<I...>) {
    ((argv[I] = static_cast<void*>(&std::get<I>(params))), ...);
}
template <class... P, class... A>
inline cudaError_t launch_graph(const Launch& L, void (*kernel)(P...), int grid, int threads, A&&... args) {
    std::tuple<std::remove_cv_t<P>...> params(static_cast<A&&>(args)...);  // the kernel's own parameter types
    void* argv[sizeof...(P) ? sizeof...(P) : 1];
    graph_arg_pointers(params, argv, std::index_sequence_for<P...>{});
    cudaKernelNodeParams kp{};
    kp.func = reinterpret_cast<void*>(kernel);
    kp.gridDim = dim3(unsigned(grid));
    kp.blockDim = dim3(unsigned(threads));
    kp.sharedMemBytes = 0;
    kp.kernelParams = argv;
    GraphSlot* s = graph_slot_acquire(reinterpret_cast<const void*>(kernel));
    cudaError_t e = cudaSuccess;
    if (!s->exec) {
        e = cudaGraphCreate(&s->graph, 0);
        if (e == cudaSuccess) e = cudaGraphAddKernelNode(&s->node, s->graph, nullptr, 0, &kp);
        if (e == cudaSuccess) e = cudaGraphInstantiate(&s->exec, s->graph, 0);
        if (e != cudaSuccess) {
            if (s->graph) cudaGraphDestroy(s->graph);
            s->graph = nullptr; s->node = nullptr; s->exec = nullptr;
        }
    } else {
        e = cudaGraphExecKernelNodeSetParams(s->exec, s->node, &kp);
    }
    if (e == cudaSuccess) e = cudaGraphLaunch(s->exec, L.stream);
    graph_slot_release(s);
    return e;
}
template <class... P, class... A>
inline cudaError_t launch_k(const Launch& L, void (*kernel)(P...), int grid, int threads, A&&... args) {
    if (L.graph && !L.overlap) return launch_graph(L, kernel, grid, threads, static_cast<A&&>(args)...);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(unsigned(grid));
    cfg.blockDim = dim3(unsigned(threads));
    cfg.stream = L.stream;
    cudaLaunchAttribute attr{};
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = L.overlap ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<A&&>(args)...);
}
#endif

// ---- launchers (each TU instantiates its kernel family) ----------------------------------------
// binary: out[i] = (f64)l[i] op (f64)r[i]; optional fused mask AND
cudaError_t launch_binary(const Launch& L, int op, int lct, const void* l, int rct, const void* r, double* out,
                          size_t n, const uint32_t* lm, const uint32_t* rm, uint32_t* om, const MaskCount& mc);
cudaError_t launch_scalar(const Launch& L, int op, int lct, const void* l, double s, double* out, size_t n);
cudaError_t launch_neg(const Launch& L, int ct, const void* a, void* out, size_t n);
cudaError_t launch_convert(const Launch& L, int sct, const void* a, int dct, void* out, size_t n);
cudaError_t launch_checked_cast(const Launch& L, int sct, const void* a, int dct, void* out, size_t n, unsigned int* fail_flag);
cudaError_t launch_copy(const Launch& L, int cell_bytes, const void* a, void* out, size_t n);
cudaError_t launch_fill(const Launch& L, int ct, void* out, size_t n, uint64_t bits);
cudaError_t launch_fill_nodata(const Launch& L, int sct, const void* a, const uint32_t* m, int dct, void* out,
                               size_t n, uint64_t nodata_bits);
cudaError_t launch_normdiff(const Launch& L, int lct, const void* l, int rct, const void* r, double* out, size_t n);
cudaError_t launch_binary_scalar(const Launch& L, int op1, int lct, const void* l, int rct, const void* r, int op2,
                                 double s, double* out, size_t n);
cudaError_t launch_binary_scalar_static(const Launch& L, int op1, int lct, const void* l, int rct, const void* r, int op2,
                                        double s, double* out, size_t n);
cudaError_t launch_scalar_scalar(const Launch& L, int op1, int ct, const void* a, double s1, int op2, double s2, double* out, size_t n);
// run-time specialised kernel of one pending expression (ec_jit.cu): `expr` is straight-line C over v0.. (operands as
// f64) and c0.. (scalars) built from ecj_add/sub/mul/div calls. Returns 0 = launched (*err = launch status),
// 1 = not available (no NVRTC / build failed; ec_last_error says why) -> the caller evaluates op by op.
constexpr int kJitInputs = 8, kJitConsts = 8, kJitOps = 48;
struct JitProgram {
    int n_in = 0, n_const = 0;
    const void* in[kJitInputs] = {};
    uint8_t ct[kJitInputs] = {};
    double consts[kJitConsts] = {};
    std::string expr;
};
int launch_jit(const Launch& L, const JitProgram& p, double* out, size_t n, cudaError_t* err);
size_t jit_cached_kernels();
size_t jit_builds();
int jit_dry_build(const JitProgram& p, std::string* source, std::string* log);
// reductions: results land in scratch.result[0..1] (device); keys are unsigned order keys
cudaError_t launch_min_max(const Launch& L, int ct, const void* a, const uint32_t* mask, size_t n,
                           const ReduceScratch& s);
cudaError_t launch_popcount(const Launch& L, const uint32_t* words, size_t nwords, const ReduceScratch& s, uint64_t second_word);
cudaError_t launch_exchange_sums(const Launch& L, const unsigned long long* local, int words, int pairs, const PeerExchange& px,
                                 unsigned long long region_off, uint64_t* host_words, uint64_t host_seq);
// this strip's raw statistics sums, added exactly over all ranks inside one extra CTA behind the statistics kernel; total[9]
ec_status moments_exchange(const ec_buf* b, const ec_mask* m, double pivot, int exp2, const PeerExchange& px, unsigned long long region_off,
                           uint64_t* total);
ec_status reduce_min_max_peer(const ec_buf* b, const ec_mask* m, const PeerExchange& px, uint64_t* k0, uint64_t* k1);
ec_status reduce_popcount_peer(const ec_mask* m, const PeerExchange& px, uint64_t* ones, uint64_t* len_sum);
// statistics extension: exact fixed-point moment sums into acc[9] (zeroed by the caller), see ec_stats.cuh
cudaError_t launch_moments(const Launch& L, int ct, const void* a, const uint32_t* mask, size_t n, double pivot, double scale,
                           unsigned long long* acc);
cudaError_t launch_int_stats(const Launch& L, int ct, const void* a, const uint32_t* mask, size_t n, unsigned long long* acc);
// Float32 cells below 2^exp2 in magnitude, read as the int32 raster rint(x * 2^(26 - exp2)): same accumulators as launch_int_stats
cudaError_t launch_quant_stats(const Launch& L, const void* a, const uint32_t* mask, size_t n, unsigned long long* acc, int exp2);
cudaError_t launch_first_diff(const Launch& L, int cell_bytes, const void* a, const void* b, size_t n,
                              const ReduceScratch& s);
// masks
cudaError_t launch_mask_build(const Launch& L, int cell_bytes, const void* a, size_t n, uint64_t sentinel_bits,
                              bool pack_bools, uint32_t* out, const MaskCount& mc);
cudaError_t launch_mask_unpack(const Launch& L, const uint32_t* m, size_t n, uint8_t* out);
cudaError_t launch_mask_bitop(const Launch& L, int mop, const uint32_t* l, const uint32_t* r, size_t n, uint32_t* out, const MaskCount& mc);
cudaError_t launch_mask_fill(const Launch& L, uint32_t* out, size_t n, bool value);
cudaError_t launch_synth(const Launch& L, int ct, void* out, size_t n, uint64_t seed, uint64_t index_offset, int kind,
                         double lo, double hi, uint64_t period, uint64_t sentinel_bits);
// keys <-> values on the host (same functions the kernels use)
void key_seeds(int ct, uint64_t* seed_min, uint64_t* seed_max);
uint64_t key_to_bits(int ct, uint64_t key);
uint64_t key_from_bits(int ct, uint64_t bits);
int64_t key_to_signed(uint64_t key);    // order-preserving int64 for NCCL min/max
uint64_t key_from_signed(int64_t skey);

}  // namespace ec
