// C ABI of liberased_cells_b200.so: device context, handles, host-side CellType/CellValue logic and
// dispatch into the kernel launchers. See include/erased_cells_b200.h for the contract and the
// reference file:line each entry point replaces.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <sys/mman.h>

#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include "ec_hostcopy.hpp"
#include "ec_internal.hpp"
#include "ec_reduce.cuh"

namespace ec {

// ---- thread-local error state --------------------------------------------------------------------
static thread_local std::string t_error;
static thread_local uint8_t t_narrow_src = 0, t_narrow_dst = 0;
static thread_local const char* t_last_kernel = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    t_error = buf;
}
ec_status cuda_fail(cudaError_t e, const char* what) {
    set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    cudaGetLastError();  // the error has been reported: do not let a recoverable one (e.g. out of memory) taint the next launch check
    if (e == cudaErrorMemoryAllocation) return EC_OOM;
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorInvalidDevice) return EC_NO_DEVICE;
    return EC_CUDA;
}
void note_launch(const char* family) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    t_last_kernel = family;
}
int env_int(const char* name, int dflt) {
    const char* s = getenv(name);
    return (s && *s) ? atoi(s) : dflt;
}
static ec_status narrowing(uint8_t src, uint8_t dst) {
    t_narrow_src = src;
    t_narrow_dst = dst;
    set_error("Invalid narrowing from cell-type %s to %s", ec_ctype_name(src), ec_ctype_name(dst));  // src/error.rs:14
    return EC_NARROWING;
}
static ec_status invalid(const char* what) {
    set_error("invalid argument: %s", what);
    return EC_INVALID_ARG;
}

// ---- device context --------------------------------------------------------------------------------
// One process drives 1..kMaxDev LOGICAL devices (ec_init / ec_init_devices / $EC_DEVICES). A logical device is a CUDA
// device plus its own pair of streams and its own block cache; the same CUDA device may appear more than once
// (that is how the sharded paths are tested on a one-GPU box). Logical device 0 is the primary: plain buffers live
// there, and a caller-supplied stream (ec_set_stream) replaces its compute stream for the calling thread.
struct DevCache {
    std::mutex mu;
    std::map<cudaStream_t, std::multimap<size_t, void*>> free_by_stream;
    std::map<void*, size_t> live;  // every block we own -> its size
    size_t cached_bytes = 0;
};
struct DevCtx {
    int phys = -1;
    cudaStream_t own = nullptr;
    cudaStream_t upload = nullptr;  // H2D of ec_buf_from_host_async: overlaps the D2H traffic of the compute stream
    DevCache cache;
    unsigned long long* mailbox = nullptr;         // peer-exchange finish of sharded reductions: 2 epochs x n_dev slots x 4 words
    unsigned long long** peer_ptrs_dev = nullptr;  // device array [n_dev]: every logical device's mailbox
};
struct Ctx {
    std::mutex mu;
    std::atomic<bool> inited{false};
    int n_dev = 0;
    DevCtx dev[kMaxDev];
    cudaDeviceProp prop;  // of logical device 0 (the GPUs of one box are alike)
    int max_grid = 0;
    std::atomic<int> overlap{1};  // programmatic dependent launch of the streaming kernels (EC_LAUNCH_OVERLAP, ec_set_launch_overlap)
    bool distinct = true;         // no CUDA device appears twice
    bool peer_ok = false;         // every device can address every other device's memory
    std::atomic<size_t> shard_min_cells{size_t(1) << 24};
    std::atomic<int> finish{FINISH_HOST};
    std::mutex px_mu;             // one peer-exchange / NCCL collective at a time
    unsigned long long epoch = 0;
    unsigned long long spin_limit = 0;
};
static Ctx g_ctx;
static thread_local cudaStream_t t_stream = nullptr;
static thread_local bool t_stream_set = false;
static thread_local int t_dev = 0;       // logical device the calling thread is working on (DevScope)
static thread_local int t_in_shard = 0;  // > 0 inside a per-strip call: the sharding policy is off

constexpr size_t kMaxReduceBlocks = 8192;
struct StreamScratch {
    cudaStream_t stream;
    ReduceScratch rs;
    unsigned long long* count_acc;  // accumulator of the mask-producing kernels on this stream (zero between launches)
};
struct ThreadDev {
    std::vector<StreamScratch> per_stream;
    uint64_t* pinned = nullptr;      // 64 words of pinned, device-mapped host memory: [0..4] single cells / keys, [8..12] reduction results, [16..24] statistics sums, [32..39] trace stamps
    uint64_t* pinned_dev = nullptr;  // its device alias: reduction kernels write results [8..11] straight into it
    uint64_t seq = 0;                // tag of the last result asked for
};
static thread_local ThreadDev t_td[kMaxDev];

static ec_status bind_device(int dev) {
    int cur = -1;
    const int phys = g_ctx.dev[dev].phys;
    if (cudaGetDevice(&cur) != cudaSuccess || cur != phys)
        if (cudaError_t e = cudaSetDevice(phys)) return cuda_fail(e, "cudaSetDevice");
    return EC_OK;
}
static ec_status init_from_env();
static ec_status ensure() {
    if (!g_ctx.inited.load(std::memory_order_acquire))
        if (ec_status s = init_from_env()) return s;
    return bind_device(t_dev);
}
static cudaStream_t cur_stream() { return (t_dev == 0 && t_stream_set) ? t_stream : g_ctx.dev[t_dev].own; }
static Launch launch_ctx() {
    return Launch{cur_stream(), g_ctx.prop.multiProcessorCount, g_ctx.max_grid, g_ctx.overlap.load(std::memory_order_relaxed) != 0};
}
// Reductions are NOT launched as programmatic dependents: their persistent grid (4-8 CTAs per SM, all resident at
// once) is scheduled while the grid before it still runs and parks on griddepcontrol.wait holding half of every SM —
// measured on the NDVI -> min_max chain at 32768^2: 30.3 ms per 8 tiles with the attribute, 25.2 ms without
// (profiles/r02_pdl_reductions.txt). $EC_PDL_REDUCE_ATTR=1 brings the attribute back for the A/B.
static Launch launch_ctx_reduce() {
    static const int attr = env_int("EC_PDL_REDUCE_ATTR", 0);
    static const int graph = env_int("EC_GRAPH_REDUCE", 1);  // 0: ordinary stream launches (A/B)
    Launch l = launch_ctx();
    l.overlap = l.overlap && attr;
    l.graph = graph != 0;
    return l;
}
// Slots of the graph-launched reductions: per (kernel, CUDA device) a short list; a launch takes the first slot nobody is
// updating (two threads reducing at once each get their own graph). Never destroyed: they live as long as the context.
struct GraphSlots {
    std::mutex mu;
    std::map<std::pair<const void*, int>, std::vector<std::unique_ptr<GraphSlot>>> slots;
};
static GraphSlots& graph_slots() {
    static GraphSlots* g = new GraphSlots;
    return *g;
}
GraphSlot* graph_slot_acquire(const void* func) {
    int dev = 0;
    cudaGetDevice(&dev);
    struct Cached { const void* func; int dev; GraphSlot* slot; };
    static thread_local Cached last[4] = {};  // the reductions a thread keeps calling: no lock, no map lookup
    for (Cached& c : last)
        if (c.func == func && c.dev == dev && c.slot && !c.slot->busy.test_and_set(std::memory_order_acquire)) return c.slot;
    GraphSlots& g = graph_slots();
    std::lock_guard<std::mutex> lk(g.mu);
    auto& list = g.slots[{func, dev}];
    GraphSlot* got = nullptr;
    for (auto& s : list)
        if (!s->busy.test_and_set(std::memory_order_acquire)) { got = s.get(); break; }
    if (!got) {
        list.emplace_back(new GraphSlot);
        got = list.back().get();
        got->busy.test_and_set(std::memory_order_acquire);
    }
    static thread_local unsigned turn = 0;
    last[turn++ % 4] = Cached{func, dev, got};
    return got;
}
int n_devices() { return g_ctx.n_dev; }
int device_phys(int dev) { return g_ctx.dev[dev].phys; }
bool shard_policy(size_t n) {
    return g_ctx.n_dev > 1 && t_in_shard == 0 && n >= g_ctx.shard_min_cells.load(std::memory_order_relaxed) && n > 0;
}
DevScope::DevScope(int dev) : prev_dev(t_dev) {
    cudaGetDevice(&prev_phys);
    t_dev = dev;
    ++t_in_shard;
    bind_device(dev);
}
DevScope::~DevScope() {
    --t_in_shard;
    t_dev = prev_dev;
    int cur = -1;
    if (prev_phys >= 0 && (cudaGetDevice(&cur) != cudaSuccess || cur != prev_phys)) cudaSetDevice(prev_phys);
}
cudaStream_t device_stream(int dev) { return (dev == 0 && t_stream_set) ? t_stream : g_ctx.dev[dev].own; }
struct PhysGuard {  // current CUDA device := the one of logical device `dev`, restored on exit (events are created on it)
    int prev = -1;
    explicit PhysGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != g_ctx.dev[dev].phys) cudaSetDevice(g_ctx.dev[dev].phys); else prev = -1;
    }
    ~PhysGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// ---- device memory: a stream-keyed caching allocator per logical device ---------------------------------
// Every op returns a fresh buffer, so allocation sits on the hot path. cudaMallocAsync's pool re-maps
// physical memory when sizes alternate (measured on B200: 0.2-2.8 s per convert sweep spent in the
// allocator), so freed blocks are kept in per-stream free lists instead, keyed by size, and reused
// exactly when a request fits within 25 % waste.
//
// Ordering. A block has a HOME stream: the stream it was allocated for, on which its producer kernel runs and to
// whose free list it returns, so reuse is ordered after every earlier use on that stream by stream order. A
// handle may also be read, written or dropped on another stream (another thread with its own ec_set_stream, the
// strip of another GPU copying from it): use_block() then makes that stream wait for everything queued on the
// home stream so far (the producer is among it) and remembers the stream as FOREIGN; when the block is released
// the home stream is made to wait for every foreign stream before the block can be handed out again.
static bool g_guard = false;
static std::atomic<uint64_t> g_guard_violations{0};
static std::mutex g_guard_mu;
static std::map<void*, std::pair<void*, size_t>> g_guard_live;  // user ptr -> (real ptr, requested bytes)
constexpr size_t kGuard = 256;

static size_t round_block(size_t bytes) {
    const size_t g = bytes < (size_t(1) << 20) ? 512 : (size_t(2) << 20);
    return (bytes + g - 1) / g * g;
}
static void cache_release_all_locked(DevCache& c) {
    for (auto& kv : c.free_by_stream)
        for (auto& b : kv.second) { cudaFree(b.second); c.live.erase(b.second); }
    c.free_by_stream.clear();
    c.cached_bytes = 0;
}
// EC_DEBUG_GUARD=1 (compute-sanitizer is not available on every pool): every block gets 256-byte red zones
// directly before its first and after its last requested byte, filled with 0xA5 and verified when the block is
// freed; no caching in this mode. ec_guard_violations() reports how many zones were found overwritten.
static ec_status guard_alloc(void** p, size_t bytes) {
    void* real = nullptr;
    if (cudaError_t e = cudaMalloc(&real, bytes + 2 * kGuard)) return cuda_fail(e, "cudaMalloc");
    char* user = static_cast<char*>(real) + kGuard;
    cudaMemsetAsync(real, 0xA5, kGuard, cur_stream());
    cudaMemsetAsync(user + bytes, 0xA5, kGuard, cur_stream());
    std::lock_guard<std::mutex> lk(g_guard_mu);
    g_guard_live[user] = {real, bytes};
    *p = user;
    return EC_OK;
}
static bool guard_free(void* p, cudaStream_t home) {
    std::pair<void*, size_t> rec;
    {
        std::lock_guard<std::mutex> lk(g_guard_mu);
        auto it = g_guard_live.find(p);
        if (it == g_guard_live.end()) return false;
        rec = it->second;
        g_guard_live.erase(it);
    }
    unsigned char zones[2 * kGuard];
    cudaStreamSynchronize(home);
    cudaMemcpy(zones, rec.first, kGuard, cudaMemcpyDeviceToHost);
    cudaMemcpy(zones + kGuard, static_cast<char*>(p) + rec.second, kGuard, cudaMemcpyDeviceToHost);
    for (size_t i = 0; i < 2 * kGuard; ++i)
        if (zones[i] != 0xA5) { g_guard_violations.fetch_add(1); fprintf(stderr, "erased_cells_b200: red zone of a %zu-byte block overwritten at %s%zu\n", rec.second, i < kGuard ? "-" : "+", i < kGuard ? kGuard - i : i - kGuard); break; }
    cudaFree(rec.first);
    return true;
}
// a block of `bytes` on the calling thread's current logical device, homed on its current stream
static ec_status dev_alloc(void** p, size_t bytes) {
    *p = nullptr;
    if (bytes == 0) return EC_OK;
    if (g_guard) return guard_alloc(p, bytes);
    bytes = round_block(bytes);
    DevCache& c = g_ctx.dev[t_dev].cache;
    std::lock_guard<std::mutex> lk(c.mu);
    auto& fl = c.free_by_stream[cur_stream()];
    auto it = fl.lower_bound(bytes);
    if (it != fl.end() && it->first <= bytes + bytes / 4) {
        *p = it->second;
        c.cached_bytes -= it->first;
        fl.erase(it);
        return EC_OK;
    }
    cudaError_t e = cudaMalloc(p, bytes);
    if (e == cudaErrorMemoryAllocation) {  // give cached blocks back and retry once
        cudaGetLastError();
        cudaDeviceSynchronize();
        cache_release_all_locked(c);
        e = cudaMalloc(p, bytes);
    }
    if (e) return cuda_fail(e, "cudaMalloc");
    c.live[*p] = bytes;
    return EC_OK;
}
// give a block back to the free list of `home` on logical device `dev`; callable from any thread
static void dev_release(void* p, int dev, cudaStream_t home, const std::vector<std::pair<cudaStream_t, int>>* foreign) {
    if (!p) return;
    if (g_guard) {
        PhysGuard g(dev);
        if (foreign) for (auto& f : *foreign) cudaStreamSynchronize(f.first);
        if (guard_free(p, home)) return;
    }
    if (foreign && !foreign->empty()) {  // the next user on the home stream must not overtake readers on other streams
        for (auto& f : *foreign) {
            PhysGuard g(f.second);
            cudaEvent_t ev;
            if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); cudaStreamSynchronize(f.first); continue; }
            cudaEventRecord(ev, f.first);
            cudaStreamWaitEvent(home, ev, 0);
            cudaEventDestroy(ev);
        }
    }
    DevCache& c = g_ctx.dev[dev].cache;
    std::lock_guard<std::mutex> lk(c.mu);
    auto it = c.live.find(p);
    if (it == c.live.end()) return;
    c.free_by_stream[home].emplace(it->second, p);
    c.cached_bytes += it->second;
    // keep the cache from squatting on the device: past half of HBM, hand everything idle back to the driver
    if (c.cached_bytes > g_ctx.prop.totalGlobalMem / 2) {
        PhysGuard g(dev);
        cudaDeviceSynchronize();
        cache_release_all_locked(c);
    }
}
DevBlock::~DevBlock() { dev_release(p, dev, home, &foreign); }
static std::shared_ptr<DevBlock> own_block(void* p) { return std::make_shared<DevBlock>(p, t_dev, cur_stream()); }
// The calling thread is about to enqueue work that touches `b` on its current stream. Blocks are written once (by
// their producer, on the home stream) and mutated only synchronously (put / extend wait for their copy), so ONE wait
// per (block, foreign stream) is enough: later uses from the same stream find it registered and cost nothing.
static inline void use_block(DevBlock* b) {
    if (!b) return;
    const cudaStream_t cur = cur_stream();
    if (b->home == cur) return;
    {
        std::lock_guard<std::mutex> lk(g_ctx.dev[b->dev].cache.mu);
        for (auto& f : b->foreign)
            if (f.first == cur) return;
        b->foreign.emplace_back(cur, t_dev);
    }
    PhysGuard g(b->dev);
    cudaEvent_t ev;
    if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); cudaStreamSynchronize(b->home); return; }
    cudaEventRecord(ev, b->home);
    cudaStreamWaitEvent(cur, ev, 0);
    cudaEventDestroy(ev);
}
// device pointer of a buffer about to be read or written on the current stream: orders the access after an
// asynchronous upload that may still be in flight on the upload stream, and after its producer on another stream
static inline void* rd(const ec_buf* b) {
    if (b->ready) cudaStreamWaitEvent(cur_stream(), b->ready, 0);
    use_block(b->blk.get());
    return b->dptr;
}
static inline uint32_t* rdm(const ec_mask* m) {
    use_block(m->blk.get());
    return m->words;
}
// scoped owners for temporaries and half-built results: an early error return must not leak device memory
struct Scratch {
    void* p = nullptr;
    int dev = 0;
    cudaStream_t home = nullptr;
    Scratch() = default;
    Scratch(const Scratch&) = delete;
    Scratch& operator=(const Scratch&) = delete;
    ~Scratch() { dev_release(p, dev, home, nullptr); }
    ec_status alloc(size_t bytes) { dev = t_dev; home = cur_stream(); return dev_alloc(&p, bytes); }
    void* release() { void* q = p; p = nullptr; return q; }
};
static ec_status sync_stream() {
    if (cudaError_t e = cudaStreamSynchronize(cur_stream())) return cuda_fail(e, "cudaStreamSynchronize");
    return EC_OK;
}
static ec_status pinned_words(uint64_t** out) {
    ThreadDev& td = t_td[t_dev];
    if (!td.pinned) {
        if (cudaError_t e = cudaHostAlloc(reinterpret_cast<void**>(&td.pinned), 512, cudaHostAllocMapped | cudaHostAllocPortable)) return cuda_fail(e, "cudaHostAlloc");
        memset(td.pinned, 0, 512);
        if (cudaError_t e = cudaHostGetDevicePointer(reinterpret_cast<void**>(&td.pinned_dev), td.pinned, 0)) return cuda_fail(e, "cudaHostGetDevicePointer");
    }
    *out = td.pinned;
    return EC_OK;
}
static ec_status stream_scratch(StreamScratch** out) {
    ThreadDev& td = t_td[t_dev];
    const cudaStream_t s = cur_stream();
    for (auto& sc : td.per_stream)
        if (sc.stream == s) { *out = &sc; return EC_OK; }
    StreamScratch sc{};
    sc.stream = s;
    uint64_t* pin;
    if (ec_status st = pinned_words(&pin)) return st;
    sc.rs.host_result = td.pinned_dev + 8;  // results land in pinned[8..11] without a D2H copy
    void* base = nullptr;
    const size_t bytes = (2 * kMaxReduceBlocks + 4 + 2 + 2) * sizeof(uint64_t);
    if (cudaError_t e = cudaMalloc(&base, bytes)) return cuda_fail(e, "cudaMalloc(reduce scratch)");
    if (cudaError_t e = cudaMemset(base, 0, bytes)) return cuda_fail(e, "cudaMemset(reduce scratch)");
    sc.rs.partials = static_cast<uint64_t*>(base);
    sc.rs.result = sc.rs.partials + 2 * kMaxReduceBlocks;
    sc.rs.ticket = reinterpret_cast<unsigned int*>(sc.rs.result + 4);
    sc.count_acc = reinterpret_cast<unsigned long long*>(sc.rs.result + 6);
    td.per_stream.push_back(sc);
    *out = &td.per_stream.back();
    return EC_OK;
}
// ---- where the time of one reduction call goes (ec_set_reduce_trace) ---------------------------------------------------
static std::atomic<int> g_reduce_trace{0};
static thread_local uint64_t t_trace_host[3];  // call entered, launch returned, result seen (CLOCK_REALTIME ns)
static thread_local const uint64_t* t_trace_gpu = nullptr;
static uint64_t host_ns() {
    timespec t;
    clock_gettime(CLOCK_REALTIME, &t);
    return uint64_t(t.tv_sec) * 1000000000ull + uint64_t(t.tv_nsec);
}
// scratch of a reduction launched now on the current stream, its result tagged with a fresh sequence number
static ec_status reduce_scratch(ReduceScratch* out, PendingReduce* pend) {
    StreamScratch* sc;
    if (ec_status st = stream_scratch(&sc)) return st;
    ThreadDev& td = t_td[t_dev];
    *out = sc->rs;
    out->px = PeerExchange{nullptr, 0, 0, 0, 0};
    out->host_seq = 0x80000000ull | (++td.seq & 0x7FFFFFFFull);
    static const int trig = env_int("EC_PDL_REDUCE_TRIGGER", 1);
    out->early_trigger = trig;
    out->trace = nullptr;
    if (g_reduce_trace) {
        for (int i = 32; i < 40; ++i) td.pinned[i] = 0;
        out->trace = td.pinned_dev + 32;
        t_trace_host[0] = host_ns();
    }
    if (pend) *pend = PendingReduce{td.pinned + 8, out->host_seq, cur_stream(), t_dev};
    return EC_OK;
}
// Wait for a tagged result in mapped pinned memory: spin on the tag word (the finishing CTA writes it last, after a
// system-scope fence) instead of synchronising the stream — the result is visible ~1 us after the kernel wrote it,
// where cudaStreamSynchronize costs 10-15 us. The stream is queried now and then so that a failed launch or a kernel
// that died surfaces as an error instead of an endless spin.
static ec_status poll_tag(volatile uint64_t* tag_word, uint64_t tag, cudaStream_t stream, int dev) {
    for (unsigned long spins = 0;; ++spins) {
        if (*tag_word == tag) { std::atomic_thread_fence(std::memory_order_acquire); return EC_OK; }
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
        if ((spins & 0x3FFF) == 0x3FFF) {
            PhysGuard g(dev);
            const cudaError_t q = cudaStreamQuery(stream);
            if (q == cudaSuccess) {
                if (*tag_word == tag) { std::atomic_thread_fence(std::memory_order_acquire); return EC_OK; }
                set_error("a kernel finished without publishing its result");
                return EC_CUDA;
            }
            if (q != cudaErrorNotReady) return cuda_fail(q, "cudaStreamQuery");
        }
    }
}
// wait until all five words of a published result carry the call's tag (see block_finish in ec_reduce.cuh)
static ec_status poll_result(const PendingReduce& p, int words = 5) {
    const uint64_t tag = p.seq << 32;
    for (unsigned long spins = 0;; ++spins) {
        bool all = true;
        for (int i = 0; i < words; ++i) all = all && ((p.pin[i] ^ tag) >> 32) == 0;
        if (all) { std::atomic_thread_fence(std::memory_order_acquire); return EC_OK; }
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
        if ((spins & 0x3FFF) == 0x3FFF) {
            PhysGuard g(p.dev);
            const cudaError_t q = cudaStreamQuery(p.stream);
            if (q == cudaSuccess) {
                all = true;
                for (int i = 0; i < words; ++i) all = all && ((p.pin[i] ^ tag) >> 32) == 0;
                if (all) return EC_OK;
                set_error("a reduction kernel finished without publishing its result");
                return EC_CUDA;
            }
            if (q != cudaErrorNotReady) return cuda_fail(q, "cudaStreamQuery");
        }
    }
}
ec_status reduce_end(const PendingReduce& p, uint64_t* r0, uint64_t* r1) {
    if (g_reduce_trace) t_trace_host[1] = host_ns();
    if (ec_status s = poll_result(p)) return s;
    if (g_reduce_trace) { t_trace_host[2] = host_ns(); t_trace_gpu = const_cast<const uint64_t*>(p.pin) + 24; }
    *r0 = (p.pin[0] & 0xFFFFFFFFull) | (p.pin[1] << 32);
    *r1 = (p.pin[2] & 0xFFFFFFFFull) | (p.pin[3] << 32);
    if ((p.pin[4] & 0xFFFFFFFFull) != 0) { set_error("a peer GPU did not deliver its partial result within the spin limit"); return EC_NCCL; }
    return EC_OK;
}

// ---- pinned slots for the counts of masks (see MaskCount in ec_common.cuh) ------------------------------------
struct CountSlot {
    volatile uint64_t* host;  // {ones, seq}; also valid as a device pointer (portable mapped allocation, unified addressing)
};
struct SlotPool {
    std::mutex mu;
    std::vector<volatile uint64_t*> free;
};
static SlotPool g_slots;
static std::atomic<uint64_t> g_count_seq{0};
static std::shared_ptr<CountSlot> take_slot() {
    std::lock_guard<std::mutex> lk(g_slots.mu);
    if (g_slots.free.empty()) {
        constexpr size_t kChunk = 1024;
        void* base = nullptr;
        if (cudaHostAlloc(&base, kChunk * 16, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        memset(base, 0, kChunk * 16);
        for (size_t i = 0; i < kChunk; ++i) g_slots.free.push_back(static_cast<volatile uint64_t*>(base) + 2 * i);
    }
    volatile uint64_t* h = g_slots.free.back();
    g_slots.free.pop_back();
    return std::shared_ptr<CountSlot>(new CountSlot{h}, [](CountSlot* s) {
        std::lock_guard<std::mutex> lk(g_slots.mu);
        g_slots.free.push_back(s->host);
        delete s;
    });
}
// arm mask `m` (about to be written by a kernel with `grid_hint` CTAs at most) for counting; returns what the kernel gets
static MaskCount arm_count(ec_mask* m) {
    MaskCount mc{nullptr, nullptr, 0};
    m->cnt_known = false;
    m->cnt_seq = 0;
    if (m->len >= (size_t(1) << 35)) return mc;  // the arrival field holds 2^24 CTAs, the smallest tile of a producer is 4096 cells
    StreamScratch* sc;
    if (stream_scratch(&sc) != EC_OK) return mc;
    if (!m->cnt) m->cnt = take_slot();
    if (!m->cnt) return mc;
    m->cnt_seq = g_count_seq.fetch_add(1) + 1;
    mc.acc = sc->count_acc;
    mc.slot = reinterpret_cast<unsigned long long*>(const_cast<uint64_t*>(m->cnt->host));
    mc.seq = m->cnt_seq;
    return mc;
}
static const MaskCount kNoCount{nullptr, nullptr, 0};

// ---- host-side CellType lattice — src/ctype.rs ------------------------------------------------------
static const size_t kSize[10] = {1, 2, 4, 8, 1, 2, 4, 8, 4, 8};
static const char* const kName[10] = {"UInt8", "UInt16", "UInt32", "UInt64", "Int8", "Int16", "Int32", "Int64", "Float32", "Float64"};
static inline bool ct_ok(unsigned ct) { return ct < 10; }
static inline bool ct_integral(uint8_t ct) { return ct < EC_FLOAT32; }
static inline bool ct_signed(uint8_t ct) { return ct >= EC_INT8; }

// src/ctype.rs:99-126: smallest type holding both; anything not representable is Float64
static uint8_t ct_union(uint8_t a, uint8_t b) {
    const bool ai = ct_integral(a), bi = ct_integral(b), as = ct_signed(a), bs = ct_signed(b);
    const size_t sa = kSize[a], sb = kSize[b];
    size_t need;
    if (ai != bi) need = ai ? std::max(sb, 2 * sa) : std::max(sa, 2 * sb);          // int with float
    else if (as != bs) need = as ? std::max(sa, 2 * sb) : std::max(sb, 2 * sa);      // signed with unsigned
    else need = std::max(sa, sb);
    const bool sg = as || bs, in = ai && bi;
    if (in) {
        switch (need) {
            case 1: return sg ? EC_INT8 : EC_UINT8;
            case 2: return sg ? EC_INT16 : EC_UINT16;
            case 4: return sg ? EC_INT32 : EC_UINT32;
            case 8: return sg ? EC_INT64 : EC_UINT64;
        }
        return EC_FLOAT64;
    }
    return need == 4 ? EC_FLOAT32 : EC_FLOAT64;
}
static inline bool ct_fits(uint8_t s, uint8_t d) { return ct_union(s, d) == d; }

// ---- host-side CellValue — src/value.rs --------------------------------------------------------------
template <class T> static inline T payload(const ec_value& v) {
    T x;
    memcpy(&x, &v.bits, sizeof(T));
    return x;
}
template <class T> static inline ec_value tagged(uint8_t ct, T x) {
    ec_value v;
    memset(&v, 0, sizeof v);
    v.ct = ct;
    memcpy(&v.bits, &x, sizeof(T));
    return v;
}
struct Widened {  // a cell widened the way ToPrimitive sees it
    bool is_float, is_neg_int;
    double f;
    uint64_t u;  // magnitude-preserving two's complement for ints
    int64_t i;
};
static Widened widen(const ec_value& v) {
    Widened w{};
    switch (v.ct) {
        case EC_UINT8: w.u = payload<uint8_t>(v); w.i = (int64_t)w.u; break;
        case EC_UINT16: w.u = payload<uint16_t>(v); w.i = (int64_t)w.u; break;
        case EC_UINT32: w.u = payload<uint32_t>(v); w.i = (int64_t)w.u; break;
        case EC_UINT64: w.u = payload<uint64_t>(v); w.i = (int64_t)w.u; break;
        case EC_INT8: w.i = payload<int8_t>(v); w.u = (uint64_t)w.i; break;
        case EC_INT16: w.i = payload<int16_t>(v); w.u = (uint64_t)w.i; break;
        case EC_INT32: w.i = payload<int32_t>(v); w.u = (uint64_t)w.i; break;
        case EC_INT64: w.i = payload<int64_t>(v); w.u = (uint64_t)w.i; break;
        case EC_FLOAT32: w.is_float = true; w.f = (double)payload<float>(v); break;  // cvtss2sd
        default: w.is_float = true; w.f = payload<double>(v); break;
    }
    w.is_neg_int = !w.is_float && ct_signed(v.ct) && w.i < 0;
    return w;
}
// `as f64` (src/value.rs:144-156)
static double value_as_f64(const ec_value& v) {
    const Widened w = widen(v);
    if (w.is_float) return w.f;
    return ct_signed(v.ct) ? (double)w.i : (double)w.u;
}
// the four ops as the SSE2 instructions of the reference's platform, operand order pinned
static double host_f64_op(int op, double a, double b) {
#if defined(__x86_64__)
    switch (op) {
        case EC_ADD: __asm__("addsd %1, %0" : "+x"(a) : "x"(b)); break;
        case EC_SUB: __asm__("subsd %1, %0" : "+x"(a) : "x"(b)); break;
        case EC_MUL: __asm__("mulsd %1, %0" : "+x"(a) : "x"(b)); break;
        default: __asm__("divsd %1, %0" : "+x"(a) : "x"(b)); break;
    }
    return a;
#else
    double r = op == EC_ADD ? a + b : op == EC_SUB ? a - b : op == EC_MUL ? a * b : a / b;
    if (r != r) {  // same rule the kernels apply
        uint64_t ab, bb, o = 0xFFF8000000000000ull;
        memcpy(&ab, &a, 8); memcpy(&bb, &b, 8);
        if (b != b) o = bb | 0x0008000000000000ull;
        if (a != a) o = ab | 0x0008000000000000ull;
        memcpy(&r, &o, 8);
    }
    return r;
#endif
}
// legal widening of a scalar (src/value.rs:74-98); caller has checked ct_fits
static ec_value value_widen(const ec_value& v, uint8_t dst) {
    if (dst == v.ct) return v;
    const Widened w = widen(v);
    switch (dst) {
        case EC_FLOAT64: return tagged<double>(dst, value_as_f64(v));
        case EC_FLOAT32: return tagged<float>(dst, w.is_float ? (float)w.f : (float)w.i);  // 8/16-bit ints: exact
        case EC_UINT16: return tagged<uint16_t>(dst, (uint16_t)w.u);
        case EC_UINT32: return tagged<uint32_t>(dst, (uint32_t)w.u);
        case EC_UINT64: return tagged<uint64_t>(dst, w.u);
        case EC_INT16: return tagged<int16_t>(dst, (int16_t)w.i);
        case EC_INT32: return tagged<int32_t>(dst, (int32_t)w.i);
        case EC_INT64: return tagged<int64_t>(dst, w.i);
    }
    return v;
}
static int cmp3u(uint64_t a, uint64_t b) { return a < b ? -1 : (a > b ? 1 : 0); }
static int cmp3i(int64_t a, int64_t b) { return a < b ? -1 : (a > b ? 1 : 0); }
// total-order key of a float payload
static int64_t total_key64(double d) {
    int64_t b;
    memcpy(&b, &d, 8);
    return b ^ (int64_t)((uint64_t)(b >> 63) >> 1);
}
static int32_t total_key32(float f) {
    int32_t b;
    memcpy(&b, &f, 4);
    return b ^ (int32_t)((uint32_t)(b >> 31) >> 1);
}
// src/value.rs:248-265
static int value_cmp(const ec_value& l, const ec_value& r) {
    const uint8_t u = ct_union(l.ct, r.ct);
    const ec_value a = value_widen(l, u), b = value_widen(r, u);
    switch (u) {
        case EC_FLOAT32: return cmp3i(total_key32(payload<float>(a)), total_key32(payload<float>(b)));
        case EC_FLOAT64: return cmp3i(total_key64(payload<double>(a)), total_key64(payload<double>(b)));
        default: return ct_signed(u) ? cmp3i(widen(a).i, widen(b).i) : cmp3u(widen(a).u, widen(b).u);
    }
}
static ec_value value_min(uint8_t ct) {
    switch (ct) {
        case EC_INT8: return tagged<int8_t>(ct, std::numeric_limits<int8_t>::min());
        case EC_INT16: return tagged<int16_t>(ct, std::numeric_limits<int16_t>::min());
        case EC_INT32: return tagged<int32_t>(ct, std::numeric_limits<int32_t>::min());
        case EC_INT64: return tagged<int64_t>(ct, std::numeric_limits<int64_t>::min());
        case EC_FLOAT32: return tagged<float>(ct, std::numeric_limits<float>::lowest());
        case EC_FLOAT64: return tagged<double>(ct, std::numeric_limits<double>::lowest());
        default: return tagged<uint64_t>(ct, 0);
    }
}
static ec_value value_max(uint8_t ct) {
    switch (ct) {
        case EC_UINT8: return tagged<uint8_t>(ct, 0xFF);
        case EC_UINT16: return tagged<uint16_t>(ct, 0xFFFF);
        case EC_UINT32: return tagged<uint32_t>(ct, 0xFFFFFFFFu);
        case EC_UINT64: return tagged<uint64_t>(ct, ~0ull);
        case EC_INT8: return tagged<int8_t>(ct, std::numeric_limits<int8_t>::max());
        case EC_INT16: return tagged<int16_t>(ct, std::numeric_limits<int16_t>::max());
        case EC_INT32: return tagged<int32_t>(ct, std::numeric_limits<int32_t>::max());
        case EC_INT64: return tagged<int64_t>(ct, std::numeric_limits<int64_t>::max());
        case EC_FLOAT32: return tagged<float>(ct, std::numeric_limits<float>::max());
        default: return tagged<double>(ct, std::numeric_limits<double>::max());
    }
}
static ec_value value_of_int(uint8_t ct, int x) {  // zero()/one()
    switch (ct) {
        case EC_FLOAT32: return tagged<float>(ct, (float)x);
        case EC_FLOAT64: return tagged<double>(ct, (double)x);
        default: return tagged<uint64_t>(ct, (uint64_t)x);
    }
}
// NoData::value — src/masked/nodata.rs:23-40
static bool nodata_sentinel(int kind, uint8_t ct, const ec_value* v, ec_value* out) {
    if (kind == EC_NODATA_NONE) return false;
    if (kind == EC_NODATA_VALUE) { *out = *v; return true; }
    if (ct == EC_FLOAT32) *out = tagged<uint32_t>(ct, 0x7FC00000u);                 // f32::NAN
    else if (ct == EC_FLOAT64) *out = tagged<uint64_t>(ct, 0x7FF8000000000000ull);  // f64::NAN
    else *out = value_min(ct);
    return true;
}

// ---- handles -----------------------------------------------------------------------------------------
static ec_status new_buf(uint8_t ct, size_t len, ec_buf** out) {
    ec_buf* b = new ec_buf;
    b->ct = ct; b->len = len; b->capacity_bytes = len * kSize[ct]; b->dev = t_dev;
    if (ec_status s = dev_alloc(&b->dptr, b->capacity_bytes)) { delete b; return s; }
    if (b->dptr) b->blk = own_block(b->dptr);
    *out = b;
    return EC_OK;
}
static size_t mask_bytes(size_t len) { return (((len + 31) / 32) * 4 + 15) & ~size_t(15); }
static ec_status new_mask(size_t len, ec_mask** out) {
    ec_mask* m = new ec_mask;
    m->len = len; m->capacity_bytes = mask_bytes(len); m->dev = t_dev;
    void* p = nullptr;
    if (ec_status s = dev_alloc(&p, m->capacity_bytes)) { delete m; return s; }
    m->words = static_cast<uint32_t*>(p);
    if (p) m->blk = own_block(p);
    if (len == 0) m->cnt_known = true;
    *out = m;
    return EC_OK;
}
#define EC_TRY(expr) do { if (ec_status _s = (expr)) return _s; } while (0)
#define EC_CUDA_TRY(expr, what) do { if (cudaError_t _e = (expr)) return cuda_fail(_e, what); } while (0)
#define EC_LAUNCH(expr, family) do { if (cudaError_t _e = (expr)) return cuda_fail(_e, family); note_launch(family); } while (0)

// per-left-type launchers generated by ec_tu_binary.cu
#define DECL(n)                                                                                                          \
    cudaError_t launch_binary_l##n(const Launch&, int, const void*, int, const void*, double*, size_t, const uint32_t*, \
                                   const uint32_t*, uint32_t*, const MaskCount&);                                       \
    cudaError_t launch_normdiff_l##n(const Launch&, const void*, int, const void*, double*, size_t);                    \
    cudaError_t launch_binary_scalar_l##n(const Launch&, int, const void*, int, const void*, int, double, double*, size_t);
DECL(0) DECL(1) DECL(2) DECL(3) DECL(4) DECL(5) DECL(6) DECL(7) DECL(8) DECL(9)
#undef DECL
using BinFn = cudaError_t (*)(const Launch&, int, const void*, int, const void*, double*, size_t, const uint32_t*, const uint32_t*, uint32_t*, const MaskCount&);
using NdFn = cudaError_t (*)(const Launch&, const void*, int, const void*, double*, size_t);
using BsFn = cudaError_t (*)(const Launch&, int, const void*, int, const void*, int, double, double*, size_t);
static const BinFn kBinary[10] = {launch_binary_l0, launch_binary_l1, launch_binary_l2, launch_binary_l3, launch_binary_l4,
                                  launch_binary_l5, launch_binary_l6, launch_binary_l7, launch_binary_l8, launch_binary_l9};
static const NdFn kNormDiff[10] = {launch_normdiff_l0, launch_normdiff_l1, launch_normdiff_l2, launch_normdiff_l3, launch_normdiff_l4,
                                   launch_normdiff_l5, launch_normdiff_l6, launch_normdiff_l7, launch_normdiff_l8, launch_normdiff_l9};
static const BsFn kBinScalar[10] = {launch_binary_scalar_l0, launch_binary_scalar_l1, launch_binary_scalar_l2, launch_binary_scalar_l3,
                                    launch_binary_scalar_l4, launch_binary_scalar_l5, launch_binary_scalar_l6, launch_binary_scalar_l7,
                                    launch_binary_scalar_l8, launch_binary_scalar_l9};

cudaError_t launch_binary(const Launch& L, int op, int lct, const void* l, int rct, const void* r, double* out, size_t n,
                          const uint32_t* lm, const uint32_t* rm, uint32_t* om, const MaskCount& mc) {
    return kBinary[lct](L, op, l, rct, r, out, n, lm, rm, om, mc);
}
cudaError_t launch_normdiff(const Launch& L, int lct, const void* l, int rct, const void* r, double* out, size_t n) {
    return kNormDiff[lct](L, l, rct, r, out, n);
}
cudaError_t launch_binary_scalar(const Launch& L, int op1, int lct, const void* l, int rct, const void* r, int op2, double s,
                                 double* out, size_t n) {
    const cudaError_t e = launch_binary_scalar_static(L, op1, lct, l, rct, r, op2, s, out, n);  // compile-time ops where instantiated
    if (e != cudaErrorNotSupported) return e;
    return kBinScalar[lct](L, op1, l, rct, r, op2, s, out, n);
}


// ---- lazy op chains ---------------------------------------------------------
// With ec_set_lazy(1) the arithmetic operators return a buffer whose value is a pending Expr over
// refcounted snapshots of its operands. The first access evaluates it and recognises three shapes that
// would otherwise cost extra passes over HBM:
//     (X - Y) / (X + Y)      -> one normalized-difference kernel  (48 -> 12 B/cell for u16 NDVI)
//     (X op1 Y) op2 scalar   -> one binary-then-scalar kernel      (27 -> 11 B/cell for u8/u16*0.5)
//     (X op1 s1) op2 s2      -> one scale-and-offset kernel        (26 -> 10 B/cell for u16 * gain + offset)
// Every op keeps its own IEEE rounding, so results are bit-identical to eager evaluation. Operands are
// immutable snapshots: put/extend on a buffer that a pending Expr still references copy it first.
static thread_local int t_lazy = 0;  // 0 eager, 1 deferred with the dedicated fused shapes, 3 also run-time specialised kernels
static std::recursive_mutex g_lazy_mu;  // pending values are shared between handles: evaluation and the hand-over of the result are serialised
constexpr int kMaxPendingDepth = 128;   // a chain deeper than this evaluates its operand first (bounds the recursion of eval and of ~Expr)
enum : int { EX_BIN = 0, EX_SCALAR = 1 };
struct Operand {  // an immutable snapshot of a buffer: while it lives, in-place mutation of the block copies first
    uint8_t ct = 0;
    size_t len = 0;
    const void* ptr = nullptr;
    std::shared_ptr<DevBlock> blk;
    std::shared_ptr<Expr> expr;
    Operand() = default;
    Operand(const Operand& o) : ct(o.ct), len(o.len), ptr(o.ptr), blk(o.blk), expr(o.expr) { if (blk) blk->snapshots.fetch_add(1); }
    Operand& operator=(const Operand& o) {
        if (this != &o) {
            if (o.blk) o.blk->snapshots.fetch_add(1);
            if (blk) blk->snapshots.fetch_sub(1);
            ct = o.ct; len = o.len; ptr = o.ptr; blk = o.blk; expr = o.expr;
        }
        return *this;
    }
    ~Operand() { if (blk) blk->snapshots.fetch_sub(1); }
    void set_block(std::shared_ptr<DevBlock> b) {
        if (b) b->snapshots.fetch_add(1);
        if (blk) blk->snapshots.fetch_sub(1);
        blk = std::move(b);
    }
};
struct Expr {
    int kind, op;
    double s;
    Operand l, r;
    size_t n;
    int depth = 1;
    bool done = false;
    std::shared_ptr<DevBlock> out;
    void* out_ptr = nullptr;
};
static bool lazy_capable(const ec_buf* b) { return b->expr || b->blk; }  // wrapped memory is not ours to keep alive
static ec_status eval(Expr& e);
static ec_status eval_operand(Operand& o);
static ec_status snapshot(const ec_buf* b, Operand* o) {
    std::lock_guard<std::recursive_mutex> lk(g_lazy_mu);
    if (b->ready) cudaStreamWaitEvent(cur_stream(), b->ready, 0);
    o->ct = b->ct; o->len = b->len; o->ptr = b->dptr; o->expr = b->expr;
    o->set_block(b->blk);
    if (o->expr && o->expr->depth >= kMaxPendingDepth) return eval_operand(*o);
    return EC_OK;
}
static int depth_of(const Operand& o) { return o.expr ? o.expr->depth : 0; }
static bool same_operand(const Operand& a, const Operand& b) {
    if (a.expr || b.expr) return a.expr == b.expr;
    return a.ptr == b.ptr && a.ct == b.ct && a.len == b.len;
}

// ---- a pending tree as the source of ONE run-time specialised kernel (ec_jit.cu), ec_set_lazy(3) -------------------
// A child is inlined when nothing else can ask for its value: it is pending and this parent holds the only
// reference. Anything shared (the user kept the handle, or two parents use it) is evaluated once and becomes an
// input. Running out of inputs / scalars / ops is not an error: the caller falls back to evaluating the children
// separately.
static bool inlineable(const Operand& o) { return o.expr && !o.expr->done && o.expr.use_count() == 1; }
// materialise every operand in the tree that cannot be inlined (shared or already-started work)
static ec_status jit_prepare(Expr& e) {
    Operand* ops[2] = {&e.l, e.kind == EX_BIN ? &e.r : nullptr};
    for (Operand* o : ops) {
        if (!o) continue;
        if (inlineable(*o)) { if (ec_status s = jit_prepare(*o->expr)) return s; }
        else if (o->expr) { if (ec_status s = eval_operand(*o)) return s; }
    }
    return EC_OK;
}
static int fusable_ops(const Expr& e) {  // number of ops the tree would fuse
    int n = 1;
    if (inlineable(e.l)) n += fusable_ops(*e.l.expr);
    if (e.kind == EX_BIN && inlineable(e.r)) n += fusable_ops(*e.r.expr);
    return n;
}
struct JitBuild {
    JitProgram p;
    int ops = 0;
    bool ok = true;
};
// A generated sub-expression and, when it is integer-typed, a bound on its magnitude in bits: integer cells (8..64) and
// sums / differences of integer-typed values (one more bit each). Such values are 0 or lie in [1, 2^bits], never -0
// (a product could be: 0 * -3), never NaN / infinite — the domain of the guard-free quotient (ecj_divi). -1: anything else.
struct JitTerm {
    std::string s;
    int int_bits;
};
static JitTerm jit_gen(JitBuild& b, Expr& e);
static JitTerm jit_value(JitBuild& b, Operand& o) {
    if (inlineable(o)) return jit_gen(b, *o.expr);
    int k = 0;
    for (; k < b.p.n_in; ++k)
        if (b.p.in[k] == o.ptr && b.p.ct[k] == o.ct) break;
    if (k == b.p.n_in) {
        if (k == kJitInputs) { b.ok = false; return {"v0", -1}; }
        use_block(o.blk.get());
        b.p.in[k] = o.ptr;
        b.p.ct[k] = o.ct;
        ++b.p.n_in;
    }
    return {"v" + std::to_string(k), ct_integral(o.ct) ? int(kSize[o.ct]) * 8 : -1};
}
static JitTerm jit_gen(JitBuild& b, Expr& e) {
    static const char* const fn[4] = {"ecj_add", "ecj_sub", "ecj_mul", "ecj_div"};
    if (!b.ok || ++b.ops > kJitOps) { b.ok = false; return {"v0", -1}; }
    const JitTerm l = jit_value(b, e.l);
    JitTerm r{"", -1};
    if (e.kind == EX_SCALAR) {  // scalars are kernel parameters: one binary serves every value
        if (b.p.n_const == kJitConsts) { b.ok = false; return {"v0", -1}; }
        b.p.consts[b.p.n_const] = e.s;
        r.s = "c" + std::to_string(b.p.n_const++);
    } else {
        r = jit_value(b, e.r);
    }
    const int op = e.op & 3;
    const bool ints = l.int_bits > 0 && r.int_bits > 0;
    const char* f = (op == EC_DIV && ints && std::max(l.int_bits, r.int_bits) <= 65) ? "ecj_divi" : fn[op];
    const int bits = (ints && (op == EC_ADD || op == EC_SUB) && std::max(l.int_bits, r.int_bits) < 65) ? std::max(l.int_bits, r.int_bits) + 1 : -1;
    return {std::string(f) + "(" + l.s + ", " + r.s + ")", bits};
}

static ec_status eval_operand(Operand& o) {
    if (!o.expr) return EC_OK;
    if (ec_status s = eval(*o.expr)) return s;
    o.ptr = o.expr->out_ptr; o.set_block(o.expr->out); o.ct = EC_FLOAT64; o.len = o.expr->n;
    o.expr.reset();
    return EC_OK;
}
static const void* op_ptr(const Operand& o) { use_block(o.blk.get()); return o.ptr; }
static ec_status eval(Expr& e) {
    if (e.done) return EC_OK;
    void* out = nullptr;
    if (ec_status s = dev_alloc(&out, e.n * sizeof(double))) return s;
    std::shared_ptr<DevBlock> blk = own_block(out);
    cudaError_t err;
    const char* family;
    Expr* cl = e.l.expr && !e.l.expr->done ? e.l.expr.get() : nullptr;
    Expr* cr = e.kind == EX_BIN && e.r.expr && !e.r.expr->done ? e.r.expr.get() : nullptr;
    if (e.kind == EX_SCALAR && cl && cl->kind == EX_SCALAR) {  // (X op1 s1) op2 s2: scale-and-offset in one pass
        if (ec_status s = eval_operand(cl->l)) return s;
        err = launch_scalar_scalar(launch_ctx(), cl->op, cl->l.ct, op_ptr(cl->l), cl->s, e.op, e.s, static_cast<double*>(out), e.n);
        family = "scalar_scalar(lazy)";
    } else if (e.kind == EX_SCALAR && cl && cl->kind == EX_BIN) {
        if (ec_status s = eval_operand(cl->l)) return s;
        if (ec_status s = eval_operand(cl->r)) return s;
        err = launch_binary_scalar(launch_ctx(), cl->op, cl->l.ct, op_ptr(cl->l), cl->r.ct, op_ptr(cl->r), e.op, e.s, static_cast<double*>(out), e.n);
        family = "binary_scalar(lazy)";
    } else if (e.kind == EX_BIN && e.op == EC_DIV && cl && cr && cl->kind == EX_BIN && cr->kind == EX_BIN && cl->op == EC_SUB &&
               cr->op == EC_ADD && same_operand(cl->l, cr->l) && same_operand(cl->r, cr->r)) {
        if (ec_status s = eval_operand(cl->l)) return s;
        if (ec_status s = eval_operand(cl->r)) return s;
        err = launch_normdiff(launch_ctx(), cl->l.ct, op_ptr(cl->l), cl->r.ct, op_ptr(cl->r), static_cast<double*>(out), e.n);
        family = "normalized_difference(lazy)";
    } else if (t_lazy == 3 && fusable_ops(e) >= 2 && [&] {  // opt-in: a longer chain as ONE kernel specialised at run time (ec_jit.cu)
                   if (jit_prepare(e) != EC_OK) return false;
                   JitBuild b;
                   b.p.expr = jit_gen(b, e).s;
                   if (!b.ok) return false;
                   if (launch_jit(launch_ctx(), b.p, static_cast<double*>(out), e.n, &err) != 0) return false;  // no NVRTC here: op by op
                   family = "expression_jit(lazy)";
                   return true;
               }()) {
    } else {
        if (ec_status s = eval_operand(e.l)) return s;
        if (e.kind == EX_BIN) {
            if (ec_status s = eval_operand(e.r)) return s;
            err = launch_binary(launch_ctx(), e.op, e.l.ct, op_ptr(e.l), e.r.ct, op_ptr(e.r), static_cast<double*>(out), e.n, nullptr, nullptr, nullptr, kNoCount);
            family = "binary";
        } else {
            err = launch_scalar(launch_ctx(), e.op, e.l.ct, op_ptr(e.l), e.s, static_cast<double*>(out), e.n);
            family = "scalar";
        }
    }
    if (err) return cuda_fail(err, family);
    note_launch(family);
    e.out = std::move(blk);
    e.out_ptr = out;
    e.done = true;
    e.l = Operand();  // let the inputs go as early as possible
    e.r = Operand();
    return EC_OK;
}
// make a (possibly pending) buffer usable: after this b->dptr is valid. Safe against concurrent readers of one handle.
static ec_status resolve(const ec_buf* b) {
    if (!b->expr) return EC_OK;
    std::lock_guard<std::recursive_mutex> lk(g_lazy_mu);
    if (!b->expr) return EC_OK;
    ec_buf* m = const_cast<ec_buf*>(b);
    if (ec_status s = eval(*m->expr)) return s;
    m->dptr = m->expr->out_ptr;
    m->blk = m->expr->out;
    m->expr.reset();
    return EC_OK;
}
static ec_buf* pending_buf(std::shared_ptr<Expr> e) {
    ec_buf* b = new ec_buf;
    b->ct = EC_FLOAT64; b->len = e->n; b->capacity_bytes = e->n * sizeof(double); b->dev = t_dev;
    b->expr = std::move(e);
    return b;
}
// In-place mutation of a block that pending expressions still read: copy first (their operands are snapshots).
// Views (ec_buf_view) are aliases, not snapshots: a put through a view or its parent is seen by both — unless a
// pending expression forced the copy, after which the mutated handle has left the shared allocation.
static ec_status ensure_unique(ec_buf* b) {
    if (!b->blk || b->blk->snapshots.load() == 0) return EC_OK;
    void* p = nullptr;
    if (ec_status s = dev_alloc(&p, b->capacity_bytes)) return s;
    std::shared_ptr<DevBlock> nb = own_block(p);
    if (cudaError_t e = launch_copy(launch_ctx(), (int)kSize[b->ct], rd(b), p, b->len)) return cuda_fail(e, "clone");
    note_launch("clone");
    b->blk = std::move(nb);
    b->dptr = p;
    return EC_OK;
}

// reduce a buffer to {min_key, max_key}: the kernel is in flight when this returns, reduce_end() waits for its result
ec_status min_max_begin(const ec_buf* b, const ec_mask* m, const PeerExchange* px, PendingReduce* pend, ReduceScratch* sc_out) {
    EC_TRY(ensure());
    if (m && m->len != b->len) { set_error("Mask and buffer must have the same length."); return EC_LEN_MISMATCH; }
    EC_TRY(resolve(b));
    ReduceScratch sc;
    EC_TRY(reduce_scratch(&sc, pend));
    if (px) sc.px = *px;
    EC_LAUNCH(launch_min_max(launch_ctx_reduce(), b->ct, rd(b), m ? rdm(m) : nullptr, b->len, sc), px ? "min_max(peer exchange)" : "min_max");
    if (sc_out) *sc_out = sc;
    return EC_OK;
}
ec_status popcount_begin(const ec_mask* m, const PeerExchange* px, uint64_t second_word, PendingReduce* pend) {
    EC_TRY(ensure());
    ReduceScratch sc;
    EC_TRY(reduce_scratch(&sc, pend));
    if (px) sc.px = *px;
    EC_LAUNCH(launch_popcount(launch_ctx_reduce(), rdm(m), (m->len + 31) / 32, sc, second_word), px ? "mask_counts(peer exchange)" : "mask_counts");
    return EC_OK;
}
ec_status first_diff_begin(const ec_buf* l, const ec_buf* r, size_t n, PendingReduce* pend) {
    EC_TRY(ensure());
    EC_TRY(resolve(l));
    EC_TRY(resolve(r));
    ReduceScratch sc;
    EC_TRY(reduce_scratch(&sc, pend));
    EC_LAUNCH(launch_first_diff(launch_ctx_reduce(), (int)kSize[l->ct], rd(l), rd(r), n, sc), "first_diff");
    return EC_OK;
}
// first differing bit of two plain masks on the current device, ~0 if none below n
ec_status mask_first_diff(const ec_mask* l, const ec_mask* r, size_t n, uint64_t* bit_out) {
    *bit_out = ~0ull;
    ReduceScratch sc;
    PendingReduce pend;
    EC_TRY(reduce_scratch(&sc, &pend));
    EC_LAUNCH(launch_first_diff(launch_ctx_reduce(), 4, rdm(l), rdm(r), (n + 31) / 32, sc), "first_diff");
    uint64_t w, unused;
    EC_TRY(reduce_end(pend, &w, &unused));
    if (w == ~0ull) return EC_OK;
    uint64_t* pin;
    EC_TRY(pinned_words(&pin));
    EC_CUDA_TRY(cudaMemcpyAsync(pin, l->words + w, 4, cudaMemcpyDeviceToHost, cur_stream()), "cudaMemcpyAsync(D2H)");
    EC_CUDA_TRY(cudaMemcpyAsync(pin + 1, r->words + w, 4, cudaMemcpyDeviceToHost, cur_stream()), "cudaMemcpyAsync(D2H)");
    EC_TRY(sync_stream());
    const uint32_t a = static_cast<uint32_t>(pin[0]), b = static_cast<uint32_t>(pin[1]);
    const uint64_t bit = w * 32 + __builtin_ctz(a ^ b);
    if (bit < n) *bit_out = bit;
    return EC_OK;
}
// Sharded reductions of the one-rank-per-process path (ec_comm.cu), finished inside the kernel over NVLink peer memory
// (see PeerExchange in ec_reduce.cuh). An empty strip still takes part in the exchange: its kernel runs over zero
// cells and contributes the seeds / zero.
ec_status reduce_min_max_peer(const ec_buf* b, const ec_mask* m, const PeerExchange& px, uint64_t* k0, uint64_t* k1) {
    PendingReduce pend;
    EC_TRY(min_max_begin(b, m, &px, &pend, nullptr));
    return reduce_end(pend, k0, k1);
}
ec_status reduce_popcount_peer(const ec_mask* m, const PeerExchange& px, uint64_t* ones, uint64_t* len_sum) {
    PendingReduce pend;
    // second word of the pair carries this strip's length so every rank also learns the total
    EC_TRY(popcount_begin(m, &px, m->len, &pend));
    return reduce_end(pend, ones, len_sum);
}

static const uint64_t kIntStatsInit[8] = {0, 0, 0, 0, 0, 0xFFFFFFFFull, 0, 0};
// The statistics kernels leave their raw accumulators in pinned[16..24] of the (thread, device) scratch: begin enqueues
// kernel + copy, end waits for that device's stream — so the strips of a sharded raster run side by side.
static ec_status stats_pending(StatsPending* p, int words) {
    uint64_t* pin;
    EC_TRY(pinned_words(&pin));
    *p = StatsPending{pin + 16, cur_stream(), t_dev, words};
    return EC_OK;
}
ec_status int_stats_begin(const ec_buf* b, const ec_mask* m, StatsPending* p) {
    EC_TRY(ensure());
    EC_TRY(resolve(b));
    EC_TRY(stats_pending(p, 8));
    Scratch acc;  // released at return: reuse of the block is ordered after the copy below by stream order
    EC_TRY(acc.alloc(sizeof kIntStatsInit));
    EC_CUDA_TRY(cudaMemcpyAsync(acc.p, kIntStatsInit, sizeof kIntStatsInit, cudaMemcpyHostToDevice, cur_stream()), "cudaMemcpyAsync(H2D)");
    EC_LAUNCH(launch_int_stats(launch_ctx(), b->ct, rd(b), m ? rdm(m) : nullptr, b->len, static_cast<unsigned long long*>(acc.p)), "int_stats");
    EC_CUDA_TRY(cudaMemcpyAsync(p->pin, acc.p, sizeof kIntStatsInit, cudaMemcpyDeviceToHost, cur_stream()), "cudaMemcpyAsync(D2H)");
    return EC_OK;
}
ec_status quant_stats_begin(const ec_buf* b, const ec_mask* m, int exp2, StatsPending* p) {
    EC_TRY(ensure());
    EC_TRY(resolve(b));
    EC_TRY(stats_pending(p, 8));
    Scratch acc;
    EC_TRY(acc.alloc(sizeof kIntStatsInit));
    EC_CUDA_TRY(cudaMemcpyAsync(acc.p, kIntStatsInit, sizeof kIntStatsInit, cudaMemcpyHostToDevice, cur_stream()), "cudaMemcpyAsync(H2D)");
    EC_LAUNCH(launch_quant_stats(launch_ctx(), rd(b), m ? rdm(m) : nullptr, b->len, static_cast<unsigned long long*>(acc.p), exp2), "quant_stats");
    EC_CUDA_TRY(cudaMemcpyAsync(p->pin, acc.p, sizeof kIntStatsInit, cudaMemcpyDeviceToHost, cur_stream()), "cudaMemcpyAsync(D2H)");
    return EC_OK;
}
ec_status moments_begin(const ec_buf* b, const ec_mask* m, double pivot, int exp2, StatsPending* p) {
    EC_TRY(ensure());
    EC_TRY(resolve(b));
    EC_TRY(stats_pending(p, EC_MOMENT_WORDS));
    Scratch acc;
    EC_TRY(acc.alloc(EC_MOMENT_WORDS * sizeof(uint64_t)));
    EC_CUDA_TRY(cudaMemsetAsync(acc.p, 0, EC_MOMENT_WORDS * sizeof(uint64_t), cur_stream()), "cudaMemsetAsync");
    EC_LAUNCH(launch_moments(launch_ctx(), b->ct, rd(b), m ? rdm(m) : nullptr, b->len, pivot, std::ldexp(1.0, -exp2),
                             static_cast<unsigned long long*>(acc.p)), "moments");
    EC_CUDA_TRY(cudaMemcpyAsync(p->pin, acc.p, EC_MOMENT_WORDS * sizeof(uint64_t), cudaMemcpyDeviceToHost, cur_stream()), "cudaMemcpyAsync(D2H)");
    return EC_OK;
}
ec_status moments_exchange(const ec_buf* b, const ec_mask* m, double pivot, int exp2, const PeerExchange& px, unsigned long long region_off,
                           uint64_t* total) {
    EC_TRY(ensure());
    EC_TRY(resolve(b));
    const bool quant = b->ct == EC_FLOAT32, ints = stats_integer_route(b->ct) || quant;
    const int K = ints ? 5 : EC_MOMENT_WORDS, pairs = ints ? 2 : 4;
    Scratch acc;  // released at return: reuse of the block is ordered after the exchange kernel by stream order
    EC_TRY(acc.alloc(EC_MOMENT_WORDS * sizeof(uint64_t)));
    if (ints) EC_CUDA_TRY(cudaMemcpyAsync(acc.p, kIntStatsInit, sizeof kIntStatsInit, cudaMemcpyHostToDevice, cur_stream()), "cudaMemcpyAsync(H2D)");
    else EC_CUDA_TRY(cudaMemsetAsync(acc.p, 0, EC_MOMENT_WORDS * sizeof(uint64_t), cur_stream()), "cudaMemsetAsync");
    unsigned long long* dacc = static_cast<unsigned long long*>(acc.p);
    if (b->len) {  // an empty strip still takes part in the exchange, with all-zero sums
        if (quant) EC_LAUNCH(launch_quant_stats(launch_ctx(), rd(b), m ? rdm(m) : nullptr, b->len, dacc, exp2), "quant_stats");
        else if (ints) EC_LAUNCH(launch_int_stats(launch_ctx(), b->ct, rd(b), m ? rdm(m) : nullptr, b->len, dacc), "int_stats");
        else EC_LAUNCH(launch_moments(launch_ctx(), b->ct, rd(b), m ? rdm(m) : nullptr, b->len, pivot, std::ldexp(1.0, -exp2), dacc), "moments");
    }
    uint64_t* pin;
    EC_TRY(pinned_words(&pin));
    ThreadDev& td = t_td[t_dev];
    PendingReduce pend{td.pinned + 40, 0x80000000ull | (++td.seq & 0x7FFFFFFFull), cur_stream(), t_dev};
    EC_LAUNCH(launch_exchange_sums(launch_ctx(), dacc, K, pairs, px, region_off, td.pinned_dev + 40, pend.seq), "statistics_sums(peer exchange)");
    EC_TRY(poll_result(pend, 2 * K + 1));
    memset(total, 0, EC_MOMENT_WORDS * sizeof(uint64_t));
    for (int k = 0; k < K; ++k) total[k] = (pend.pin[2 * k] & 0xFFFFFFFFull) | (pend.pin[2 * k + 1] << 32);
    if ((pend.pin[2 * K] & 0xFFFFFFFFull) != 0) { set_error("a peer GPU did not deliver its statistics sums within the spin limit"); return EC_NCCL; }
    return EC_OK;
}
ec_status stats_end(const StatsPending& p, uint64_t* w) {
    PhysGuard g(p.dev);
    EC_CUDA_TRY(cudaStreamSynchronize(p.stream), "cudaStreamSynchronize");
    memcpy(w, p.pin, p.words * sizeof(uint64_t));
    return EC_OK;
}
bool stats_integer_route(uint8_t ct) { return ct_integral(ct) && kSize[ct] <= 4; }
// the int-route words {count, A.lo, A.hi, B.lo, B.hi, min, max, -} of one strip as min/max values of cell type ct
void int_stats_min_max(uint8_t ct, const uint64_t* w, ec_value* mn, ec_value* mx) {
    if (w[0] == 0) {  // no valid cell: the seeds, as min_max reports them
        uint64_t k[2];
        key_seeds(ct, &k[0], &k[1]);
        *mn = tagged<uint64_t>(ct, key_to_bits(ct, k[0]));
        *mx = tagged<uint64_t>(ct, key_to_bits(ct, k[1]));
    } else {
        const uint64_t bias = ct_signed(ct) ? 1ull << (8 * kSize[ct] - 1) : 0ull;
        *mn = tagged<uint64_t>(ct, w[5] ^ bias);
        *mx = tagged<uint64_t>(ct, w[6] ^ bias);
    }
}

// In-place mutation of words that a clone still shares: copy first.
static ec_status mask_unique(ec_mask* m) {
    if (!m->blk || m->blk.use_count() == 1) return EC_OK;
    void* p = nullptr;
    EC_TRY(dev_alloc(&p, m->capacity_bytes));
    std::shared_ptr<DevBlock> nb = own_block(p);
    EC_CUDA_TRY(cudaMemcpyAsync(p, rdm(m), m->capacity_bytes, cudaMemcpyDeviceToDevice, cur_stream()), "cudaMemcpyAsync(D2D)");
    m->blk = std::move(nb);
    m->words = static_cast<uint32_t*>(p);
    return EC_OK;
}
// Number of set bits of a plain mask. Known at once when the kernel that produced the mask counted for it (it has
// published into the mask's pinned slot by the time it completes), else one popcount pass; cached either way.
static ec_status mask_ones_begin(const ec_mask* m, PendingReduce* pend, bool* launched) {
    *launched = false;
    if (m->cnt_known || m->cnt_seq != 0 || m->len == 0) return EC_OK;
    EC_TRY(popcount_begin(m, nullptr, 0, pend));
    *launched = true;
    return EC_OK;
}
static ec_status mask_ones_end(const ec_mask* m, const PendingReduce& pend, bool launched, uint64_t* ones) {
    ec_mask* mm = const_cast<ec_mask*>(m);  // a cache: concurrent readers store the same value
    if (!m->cnt_known) {
        uint64_t v = 0;
        if (launched) {
            uint64_t unused;
            EC_TRY(reduce_end(pend, &v, &unused));
        } else if (m->cnt_seq != 0) {
            EC_TRY(poll_tag(m->cnt->host + 1, m->cnt_seq, m->blk ? m->blk->home : cur_stream(), m->dev));
            v = m->cnt->host[0];
        }
        mm->ones = v;
        __atomic_store_n(&mm->cnt_known, true, __ATOMIC_RELEASE);
    }
    *ones = m->ones;
    return EC_OK;
}
ec_status mask_ones(const ec_mask* m, uint64_t* ones) {
    DevScope on(m->dev);
    PendingReduce pend;
    bool launched;
    EC_TRY(mask_ones_begin(m, &pend, &launched));
    return mask_ones_end(m, pend, launched, ones);
}

// first use without an explicit ec_init: $EC_DEVICES ("0,1,2,3": one process drives them all), else $EC_DEVICE, else $LOCAL_RANK, else 0
static ec_status init_from_env() {
    int devs[kMaxDev], n = 0;
    if (const char* list = getenv("EC_DEVICES")) {
        for (const char* p = list; *p && n < kMaxDev;) {
            char* end = nullptr;
            const long v = strtol(p, &end, 10);
            if (end == p) break;
            devs[n++] = static_cast<int>(v);
            p = *end == ',' ? end + 1 : end;
        }
    }
    if (n == 0) devs[n++] = env_int("EC_DEVICE", env_int("LOCAL_RANK", 0));
    return ec_init_devices(devs, n);
}

}  // namespace ec

using namespace ec;

extern "C" {

// ---- library / device context ------------------------------------------------------------------------
int ec_abi_version(void) { return EC_ABI_VERSION; }
const char* ec_last_error(void) { return t_error.c_str(); }
void ec_last_narrowing(uint8_t* src, uint8_t* dst) {
    if (src) *src = t_narrow_src;
    if (dst) *dst = t_narrow_dst;
}
ec_status ec_init_devices(const int* devices, int n) {
    std::lock_guard<std::mutex> lk(g_ctx.mu);
    if (n < 1 || n > kMaxDev || !devices) return invalid("ec_init_devices: 1..16 devices");
    if (g_ctx.inited) {
        bool same = n == g_ctx.n_dev;
        for (int g = 0; same && g < n; ++g) same = devices[g] == g_ctx.dev[g].phys;
        if (!same) return invalid("ec_init: this process is already bound to another set of devices");
        return EC_OK;
    }
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        set_error("no usable CUDA device (%s); erased_cells_b200 has no CPU path", e ? cudaGetErrorString(e) : "0 devices");
        return EC_NO_DEVICE;
    }
    for (int g = 0; g < n; ++g)
        if (devices[g] < 0 || devices[g] >= count) return invalid("ec_init: device index out of range");
    int prev = -1;
    cudaGetDevice(&prev);
    bool distinct = true, peer_ok = true;
    for (int g = 0; g < n; ++g) {
        DevCtx& d = g_ctx.dev[g];
        EC_CUDA_TRY(cudaSetDevice(devices[g]), "cudaSetDevice");
        EC_CUDA_TRY(cudaStreamCreateWithFlags(&d.own, cudaStreamNonBlocking), "cudaStreamCreate");
        EC_CUDA_TRY(cudaStreamCreateWithFlags(&d.upload, cudaStreamNonBlocking), "cudaStreamCreate");
        d.phys = devices[g];
        for (int h = 0; h < g; ++h) {
            if (devices[h] == devices[g]) { distinct = false; continue; }
            // both directions: kernels of either GPU store into the other's mailbox, strips are copied GPU to GPU
            for (int dir = 0; dir < 2; ++dir) {
                const int from = dir ? devices[h] : devices[g], to = dir ? devices[g] : devices[h];
                int can = 0;
                if (cudaDeviceCanAccessPeer(&can, from, to) != cudaSuccess || !can) { cudaGetLastError(); peer_ok = false; continue; }
                cudaSetDevice(from);
                const cudaError_t pe = cudaDeviceEnablePeerAccess(to, 0);
                if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) peer_ok = false;
                cudaGetLastError();
            }
            cudaSetDevice(devices[g]);
        }
    }
    EC_CUDA_TRY(cudaSetDevice(devices[0]), "cudaSetDevice");
    EC_CUDA_TRY(cudaGetDeviceProperties(&g_ctx.prop, devices[0]), "cudaGetDeviceProperties");
    g_ctx.max_grid = env_int("EC_MAX_GRID", 0);
    g_ctx.overlap = env_int("EC_LAUNCH_OVERLAP", 1) != 0 && g_ctx.prop.major >= 9;
    g_guard = env_int("EC_DEBUG_GUARD", 0) != 0;
    g_ctx.n_dev = n;
    g_ctx.distinct = distinct;
    g_ctx.peer_ok = peer_ok && n > 1;
    if (const char* t = getenv("EC_SHARD_MIN_CELLS")) if (*t) g_ctx.shard_min_cells = strtoull(t, nullptr, 0);
    const char* fin = getenv("EC_SHARD_FINISH");
    g_ctx.finish = (fin && !strcmp(fin, "peer")) ? FINISH_PEER : (fin && !strcmp(fin, "nccl")) ? FINISH_NCCL : FINISH_HOST;
    // a peer that never arrives must become an error, not a hung GPU: give up after EC_PEER_WAIT_SECONDS of SM clock ticks
    int khz = 1900000;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, devices[0]);
    g_ctx.spin_limit = static_cast<unsigned long long>(env_int("EC_PEER_WAIT_SECONDS", 20)) * 1000ull * static_cast<unsigned long long>(khz);
    g_ctx.inited.store(true, std::memory_order_release);
    if (prev >= 0 && n > 1 && prev != devices[0]) cudaSetDevice(devices[0]);
    return EC_OK;
}
ec_status ec_init(int device) {
    if (g_ctx.inited.load(std::memory_order_acquire)) {  // idempotent for the primary device of an initialised library
        if (device != g_ctx.dev[0].phys) return invalid("ec_init: this process is already bound to another device");
        return EC_OK;
    }
    return ec_init_devices(&device, 1);
}
int ec_device_count(void) { return g_ctx.inited.load(std::memory_order_acquire) ? g_ctx.n_dev : 0; }
size_t ec_set_shard_min_cells(size_t cells) { return g_ctx.shard_min_cells.exchange(cells ? cells : 1); }
int ec_set_shard_finish(int mode) {
    if (mode < FINISH_HOST || mode > FINISH_NCCL) return -1;
    return g_ctx.finish.exchange(mode);
}
int ec_set_reduce_trace(int on) { return g_reduce_trace.exchange(on != 0); }
__global__ void globaltimer_probe(volatile uint64_t* out) {
    uint64_t t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    out[0] = t;
}
ec_status ec_reduce_trace_get(uint64_t* out12) {
    EC_TRY(ensure());
    if (!t_trace_gpu) return invalid("ec_reduce_trace_get: no traced reduction on this thread yet");
    for (int i = 0; i < 3; ++i) out12[i] = t_trace_host[i];
    for (int i = 0; i < 6; ++i) out12[3 + i] = t_trace_gpu[i];
    // offset of %globaltimer against CLOCK_REALTIME: a one-thread kernel stamps mapped memory, the host notes when it sees
    // the stamp; the smallest (host - gpu) over a few tries is the offset plus the ~1 us it takes the stamp to arrive
    uint64_t* pin;
    EC_TRY(pinned_words(&pin));
    ThreadDev& td = t_td[t_dev];
    int64_t best = INT64_MAX;
    for (int rep = 0; rep < 20; ++rep) {
        pin[4] = 0;
        globaltimer_probe<<<1, 1, 0, cur_stream()>>>(td.pinned_dev + 4);
        volatile uint64_t* w = pin + 4;
        while (*w == 0) {}
        const int64_t d = static_cast<int64_t>(host_ns()) - static_cast<int64_t>(*w);
        if (d < best) best = d;
    }
    EC_TRY(sync_stream());
    out12[9] = static_cast<uint64_t>(best);
    out12[10] = out12[11] = 0;
    return EC_OK;
}
ec_status ec_device_info_get(ec_device_info* out) {
    EC_TRY(ensure());
    memset(out, 0, sizeof *out);
    out->device = g_ctx.dev[0].phys;
    out->sm_count = g_ctx.prop.multiProcessorCount;
    out->cc_major = g_ctx.prop.major;
    out->cc_minor = g_ctx.prop.minor;
    out->l2_bytes = g_ctx.prop.l2CacheSize;
    out->total_mem_bytes = g_ctx.prop.totalGlobalMem;
    strncpy(out->name, g_ctx.prop.name, sizeof(out->name) - 1);
    return EC_OK;
}
ec_status ec_set_stream(void* cuda_stream) {
    EC_TRY(ensure());
    t_stream = static_cast<cudaStream_t>(cuda_stream);
    t_stream_set = cuda_stream != nullptr;
    return EC_OK;
}
void* ec_get_stream(void) { return ensure() == EC_OK ? cur_stream() : nullptr; }
ec_status ec_synchronize(void) {
    EC_TRY(ensure());
    for (int g = 1; g < g_ctx.n_dev; ++g) {  // strips of sharded buffers live on the other devices' streams
        PhysGuard pg(g);
        EC_CUDA_TRY(cudaStreamSynchronize(g_ctx.dev[g].own), "cudaStreamSynchronize");
    }
    return sync_stream();
}
namespace ec { void stage_pool_trim(); }
ec_status ec_trim(void) {
    EC_TRY(ec_synchronize());
    stage_pool_trim();
    for (int g = 0; g < g_ctx.n_dev; ++g) {
        PhysGuard pg(g);
        std::lock_guard<std::mutex> lk(g_ctx.dev[g].cache.mu);
        cache_release_all_locked(g_ctx.dev[g].cache);
    }
    return EC_OK;
}
size_t ec_cached_bytes(void) {
    size_t total = 0;
    for (int g = 0; g < g_ctx.n_dev; ++g) {
        std::lock_guard<std::mutex> lk(g_ctx.dev[g].cache.mu);
        total += g_ctx.dev[g].cache.cached_bytes;
    }
    return total;
}
uint64_t ec_guard_violations(void) { return g_guard_violations.load(); }
ec_status ec_set_lazy(int mode) {
    if (mode < 0 || mode > 3 || mode == 2) return invalid("lazy mode (0 eager, 1 fused shapes, 3 run-time specialised kernels; the expression VM of ABI 1 is gone)");
    t_lazy = mode;
    return EC_OK;
}
int ec_get_lazy(void) { return t_lazy; }
int ec_set_launch_overlap(int on) {
    if (ensure() != EC_OK) return -1;
    return g_ctx.overlap.exchange(on != 0 && g_ctx.prop.major >= 9);
}
size_t ec_jit_cached_kernels(void) { return jit_cached_kernels(); }
size_t ec_jit_builds(void) { return jit_builds(); }
ec_status ec_jit_dry_build(const uint8_t* cell_types, int n_in, int n_const, const char* expr, char* log, size_t log_capacity) {
    if (n_in < 1 || n_in > kJitInputs || n_const < 0 || n_const > kJitConsts || !expr) return invalid("ec_jit_dry_build: operand / scalar count");
    JitProgram p;
    p.n_in = n_in;
    p.n_const = n_const;
    for (int k = 0; k < n_in; ++k) {
        if (!ct_ok(cell_types[k])) return invalid("cell type");
        p.ct[k] = cell_types[k];
    }
    p.expr = expr;
    std::string source, text;
    const int rc = jit_dry_build(p, &source, &text);
    if (log && log_capacity) { strncpy(log, text.c_str(), log_capacity - 1); log[log_capacity - 1] = 0; }
    if (rc == 1) { set_error("expression JIT unavailable: %s", text.c_str()); return EC_NO_DEVICE; }
    if (rc != 0) { set_error("expression JIT: NVRTC build failed: %.900s", text.c_str()); return EC_INVALID_ARG; }
    return EC_OK;
}
uint64_t ec_kernel_launches(void) { return g_launches.load(); }
const char* ec_last_kernel(void) { return t_last_kernel; }
ec_status ec_event_create(ec_event** out) {
    EC_TRY(ensure());
    ec_event* e = new ec_event{};
    if (cudaError_t err = cudaEventCreate(&e->ev)) { delete e; return cuda_fail(err, "cudaEventCreate"); }
    *out = e;
    return EC_OK;
}
ec_status ec_event_record(ec_event* e) {
    EC_TRY(ensure());
    EC_CUDA_TRY(cudaEventRecord(e->ev, cur_stream()), "cudaEventRecord");
    return EC_OK;
}
ec_status ec_event_elapsed_ms(ec_event* a, ec_event* b, float* ms) {
    EC_CUDA_TRY(cudaEventSynchronize(b->ev), "cudaEventSynchronize");
    EC_CUDA_TRY(cudaEventElapsedTime(ms, a->ev, b->ev), "cudaEventElapsedTime");
    return EC_OK;
}
void ec_event_destroy(ec_event* e) {
    if (e) { cudaEventDestroy(e->ev); delete e; }
}
ec_status ec_host_alloc(size_t bytes, void** out) {
    EC_TRY(ensure());
    if (cudaError_t e = cudaMallocHost(out, bytes ? bytes : 1)) { cuda_fail(e, "cudaMallocHost"); return EC_OOM; }
    return EC_OK;
}
void ec_host_free(void* p) { if (p) cudaFreeHost(p); }
ec_status ec_host_register(void* p, size_t bytes) {
    EC_TRY(ensure());
    EC_CUDA_TRY(cudaHostRegister(p, bytes, cudaHostRegisterDefault), "cudaHostRegister");
    return EC_OK;
}
ec_status ec_host_unregister(void* p) {
    EC_CUDA_TRY(cudaHostUnregister(p), "cudaHostUnregister");
    return EC_OK;
}

// ---- CellType ----------------------------------------------------------------------------------------
uint8_t ec_ctype_union(uint8_t a, uint8_t b) { return (ct_ok(a) && ct_ok(b)) ? ct_union(a, b) : EC_FLOAT64; }
int ec_ctype_can_fit_into(uint8_t a, uint8_t b) { return ct_ok(a) && ct_ok(b) && ct_fits(a, b); }
size_t ec_ctype_size_of(uint8_t ct) { return ct_ok(ct) ? kSize[ct] : 0; }
int ec_ctype_is_integral(uint8_t ct) { return ct_ok(ct) && ct_integral(ct); }
int ec_ctype_is_signed(uint8_t ct) { return ct_ok(ct) && ct_signed(ct); }
const char* ec_ctype_name(uint8_t ct) { return ct_ok(ct) ? kName[ct] : "?"; }
ec_status ec_ctype_from_name(const char* s, uint8_t* out) {
    for (uint8_t i = 0; i < 10; ++i)
        if (strcmp(s, kName[i]) == 0) { *out = i; return EC_OK; }
    set_error("Unable to parse %s as a CellType", s);  // src/error.rs:20-21
    return EC_PARSE;
}
ec_status ec_ctype_min_value(uint8_t ct, ec_value* out) { if (!ct_ok(ct)) return invalid("cell type"); *out = value_min(ct); return EC_OK; }
ec_status ec_ctype_max_value(uint8_t ct, ec_value* out) { if (!ct_ok(ct)) return invalid("cell type"); *out = value_max(ct); return EC_OK; }
ec_status ec_ctype_zero(uint8_t ct, ec_value* out) { if (!ct_ok(ct)) return invalid("cell type"); *out = value_of_int(ct, 0); return EC_OK; }
ec_status ec_ctype_one(uint8_t ct, ec_value* out) { if (!ct_ok(ct)) return invalid("cell type"); *out = value_of_int(ct, 1); return EC_OK; }

// ---- CellValue ---------------------------------------------------------------------------------------
ec_status ec_value_convert(const ec_value* v, uint8_t ct, ec_value* out) {
    if (!ct_ok(v->ct) || !ct_ok(ct)) return invalid("cell type");
    if (!ct_fits(v->ct, ct)) return narrowing(v->ct, ct);
    *out = value_widen(*v, ct);
    return EC_OK;
}
ec_status ec_value_binary(int op, const ec_value* l, const ec_value* r, ec_value* out) {
    if (!ct_ok(l->ct) || !ct_ok(r->ct) || op < 0 || op > 3) return invalid("cell type / op");
    // unify (src/value.rs:103-107) is value-exact for every pair, so `as f64` of the operands is the
    // same number the reference feeds to the f64 op
    const uint8_t u = ct_union(l->ct, r->ct);
    *out = tagged<double>(EC_FLOAT64, host_f64_op(op, value_as_f64(value_widen(*l, u)), value_as_f64(value_widen(*r, u))));
    return EC_OK;
}
ec_status ec_value_neg(const ec_value* v, ec_value* out) {
    if (!ct_ok(v->ct)) return invalid("cell type");
    const Widened w = widen(*v);
    switch (v->ct) {
        case EC_UINT8: *out = tagged<int16_t>(EC_INT16, (int16_t)(-(int)w.u)); break;
        case EC_UINT16: *out = tagged<int32_t>(EC_INT32, -(int32_t)w.u); break;
        case EC_UINT32: case EC_UINT64: {
            uint64_t b; const double d = (double)w.u; memcpy(&b, &d, 8);
            *out = tagged<uint64_t>(EC_FLOAT64, b ^ 0x8000000000000000ull);
            break;
        }
        case EC_INT8: *out = tagged<int8_t>(v->ct, (int8_t)(0u - (uint8_t)w.i)); break;
        case EC_INT16: *out = tagged<int16_t>(v->ct, (int16_t)(0u - (uint16_t)w.i)); break;
        case EC_INT32: *out = tagged<int32_t>(v->ct, (int32_t)(0u - (uint32_t)w.i)); break;
        case EC_INT64: *out = tagged<int64_t>(v->ct, (int64_t)(0ull - (uint64_t)w.i)); break;
        case EC_FLOAT32: *out = tagged<uint32_t>(v->ct, payload<uint32_t>(*v) ^ 0x80000000u); break;
        default: *out = tagged<uint64_t>(v->ct, v->bits ^ 0x8000000000000000ull); break;
    }
    return EC_OK;
}
ec_status ec_value_cmp(const ec_value* l, const ec_value* r, int* ordering) {
    if (!ct_ok(l->ct) || !ct_ok(r->ct)) return invalid("cell type");
    *ordering = value_cmp(*l, *r);
    return EC_OK;
}
ec_status ec_value_to_f64(const ec_value* v, double* out, int* is_some) {
    if (!ct_ok(v->ct)) return invalid("cell type");
    *out = value_as_f64(*v);
    *is_some = 1;
    return EC_OK;
}
// num-traits 0.2.17 float->int: Some(trunc) iff inside the representable open interval
ec_status ec_value_to_i64(const ec_value* v, int64_t* out, int* is_some) {
    if (!ct_ok(v->ct)) return invalid("cell type");
    const Widened w = widen(*v);
    if (w.is_float) {
        *is_some = (w.f >= -9223372036854775808.0 && w.f < 9223372036854775808.0);
        *out = *is_some ? (int64_t)w.f : 0;
    } else if (v->ct == EC_UINT64) {
        *is_some = w.u <= (uint64_t)std::numeric_limits<int64_t>::max();
        *out = *is_some ? (int64_t)w.u : 0;
    } else { *is_some = 1; *out = w.i; }
    return EC_OK;
}
ec_status ec_value_to_u64(const ec_value* v, uint64_t* out, int* is_some) {
    if (!ct_ok(v->ct)) return invalid("cell type");
    const Widened w = widen(*v);
    if (w.is_float) {
        *is_some = (w.f > -1.0 && w.f < 18446744073709551616.0);
        *out = *is_some ? (uint64_t)w.f : 0;
    } else { *is_some = !w.is_neg_int; *out = *is_some ? w.u : 0; }
    return EC_OK;
}

// `self.to_<p>()` on a CellValue (src/value.rs:92 call shape; also Extend, src/buffer.rs:212, and the GDAL nodata
// conversion, src/gdal/mod.rs:59): the VALUE-checked num-traits chain — ints through i64/u64 with range checks,
// floats truncated iff inside the target's range, anything -> f32 through f64.
ec_status ec_value_to_prim(const ec_value* v, uint8_t ct, ec_value* out, int* is_some) {
    if (!ct_ok(v->ct) || !ct_ok(ct)) return invalid("cell type");
    const Widened w = widen(*v);
    *is_some = 1;
    if (ct == EC_FLOAT64) { *out = tagged<double>(ct, value_as_f64(*v)); return EC_OK; }
    if (ct == EC_FLOAT32) { *out = tagged<float>(ct, (float)value_as_f64(*v)); return EC_OK; }
    if (ct_signed(ct)) {
        int64_t t = 0;
        int some = 0;
        EC_TRY(ec_value_to_i64(v, &t, &some));
        const int bits = int(kSize[ct]) * 8;
        const int64_t hi = bits == 64 ? std::numeric_limits<int64_t>::max() : (int64_t(1) << (bits - 1)) - 1, lo = -hi - 1;
        if (!some || t < lo || t > hi) { *is_some = 0; return EC_OK; }
        *out = tagged<int64_t>(ct, t);
        out->bits &= bits == 64 ? ~0ull : ((1ull << bits) - 1);
        return EC_OK;
    }
    uint64_t t = 0;
    int some = 0;
    EC_TRY(ec_value_to_u64(v, &t, &some));
    const int bits = int(kSize[ct]) * 8;
    const uint64_t hi = bits == 64 ? ~0ull : ((1ull << bits) - 1);
    if (!some || t > hi) { *is_some = 0; return EC_OK; }
    *out = tagged<uint64_t>(ct, t);
    (void)w;
    return EC_OK;
}

// ---- CellBuffer --------------------------------------------------------------------------------------
struct BufOwner {  // a half-built result: an early error return frees it
    ec_buf* b = nullptr;
    BufOwner() = default;
    BufOwner(const BufOwner&) = delete;
    BufOwner& operator=(const BufOwner&) = delete;
    ~BufOwner() { if (b) ec_buf_free(b); }
    ec_buf* release() { ec_buf* q = b; b = nullptr; return q; }
};
// a Vec<T> (pageable memory) into a fresh plain buffer, through the staging pipeline; the upload stream is ordered behind the
// work still queued on the block (it may be recycled) and is drained before the call returns
static ec_status staged_upload(ec_buf* b, const void* host) {
    const cudaStream_t up = g_ctx.dev[t_dev].upload;
    cudaEvent_t after = nullptr;
    EC_CUDA_TRY(cudaEventCreateWithFlags(&after, cudaEventDisableTiming), "cudaEventCreate");
    cudaError_t e = cudaEventRecord(after, cur_stream());
    if (e == cudaSuccess) e = cudaStreamWaitEvent(up, after, 0);
    cudaEventDestroy(after);
    if (e != cudaSuccess) return cuda_fail(e, "cudaStreamWaitEvent");
    const HostSeg seg{t_dev, b->dptr, const_cast<void*>(host), b->len * kSize[b->ct], up};
    return staged_transfer(&seg, 1, true);
}
ec_status ec_buf_from_host(uint8_t ct, const void* host, size_t len, ec_buf** out) {
    EC_TRY(ensure());
    if (!ct_ok(ct)) return invalid("cell type");
    if (shard_policy(len)) return sh_from_host(ct, host, len, false, out);
    BufOwner o;
    EC_TRY(new_buf(ct, len, &o.b));
    if (len && staged_wanted(host, len * kSize[ct])) EC_TRY(staged_upload(o.b, host));
    else if (len) EC_CUDA_TRY(cudaMemcpyAsync(o.b->dptr, host, len * kSize[ct], cudaMemcpyHostToDevice, cur_stream()), "cudaMemcpyAsync(H2D)");
    *out = o.release();
    return EC_OK;
}
ec_status ec_buf_from_host_async(uint8_t ct, const void* host, size_t len, ec_buf** out) {
    EC_TRY(ensure());
    if (!ct_ok(ct)) return invalid("cell type");
    if (shard_policy(len)) return sh_from_host(ct, host, len, true, out);
    BufOwner o;
    EC_TRY(new_buf(ct, len, &o.b));
    ec_buf* b = o.b;
    if (len && staged_wanted(host, len * kSize[ct])) {
        EC_TRY(staged_upload(b, host));  // pageable memory: nothing to wait for afterwards
    } else if (len) {
        const cudaStream_t up = g_ctx.dev[t_dev].upload;
        EC_CUDA_TRY(cudaEventCreateWithFlags(&b->ready, cudaEventDisableTiming), "cudaEventCreate");
        // the block may have been recycled from work still queued on the current stream: upload after it
        EC_CUDA_TRY(cudaEventRecord(b->ready, cur_stream()), "cudaEventRecord");
        EC_CUDA_TRY(cudaStreamWaitEvent(up, b->ready, 0), "cudaStreamWaitEvent");
        EC_CUDA_TRY(cudaMemcpyAsync(b->dptr, host, len * kSize[ct], cudaMemcpyHostToDevice, up), "cudaMemcpyAsync(H2D)");
        EC_CUDA_TRY(cudaEventRecord(b->ready, up), "cudaEventRecord");
    }
    *out = o.release();
    return EC_OK;
}
ec_status ec_buf_wait(const ec_buf* b) {
    if (is_sharded(b)) {
        for (const ec_buf* part : b->parts) EC_TRY(ec_buf_wait(part));
        return EC_OK;
    }
    EC_TRY(resolve(b));
    if (b->ready) EC_CUDA_TRY(cudaEventSynchronize(b->ready), "cudaEventSynchronize");
    return EC_OK;
}
ec_status ec_buf_with_defaults(size_t len, uint8_t ct, ec_buf** out) {
    EC_TRY(ensure());
    if (!ct_ok(ct)) return invalid("cell type");
    if (shard_policy(len)) return sh_generate(ct, len, [&](size_t, size_t n, ec_buf** o) { return ec_buf_with_defaults(n, ct, o); }, out);
    BufOwner o;
    EC_TRY(new_buf(ct, len, &o.b));
    if (len) EC_CUDA_TRY(cudaMemsetAsync(o.b->dptr, 0, len * kSize[ct], cur_stream()), "cudaMemsetAsync");
    *out = o.release();
    return EC_OK;
}
ec_status ec_buf_fill(size_t len, const ec_value* value, ec_buf** out) {
    EC_TRY(ensure());
    if (!ct_ok(value->ct)) return invalid("cell type");
    if (shard_policy(len)) return sh_generate(value->ct, len, [&](size_t, size_t n, ec_buf** o) { return ec_buf_fill(n, value, o); }, out);
    BufOwner o;
    EC_TRY(new_buf(value->ct, len, &o.b));
    if (len) EC_LAUNCH(launch_fill(launch_ctx(), o.b->ct, o.b->dptr, len, value->bits), "fill");
    *out = o.release();
    return EC_OK;
}
ec_status ec_buf_wrap_device(uint8_t ct, void* device_ptr, size_t len, ec_buf** out) {
    EC_TRY(ensure());
    if (!ct_ok(ct)) return invalid("cell type");
    if (reinterpret_cast<uintptr_t>(device_ptr) % 32 != 0) return invalid("device pointer must be 32-byte aligned (row strips start on 128-cell boundaries)");
    ec_buf* b = new ec_buf;
    b->ct = ct; b->owned = false; b->len = len; b->capacity_bytes = len * kSize[ct]; b->dptr = device_ptr; b->dev = t_dev;
    *out = b;
    return EC_OK;
}
ec_status ec_buf_view(const ec_buf* b, size_t offset_cells, size_t len, ec_buf** out) {
    EC_TRY(ensure());
    if (offset_cells > b->len || len > b->len - offset_cells) { set_error("view [%zu, %zu) outside a buffer of %zu cells", offset_cells, offset_cells + len, b->len); return EC_OOB; }
    if ((offset_cells * kSize[b->ct]) % 32 != 0) return invalid("a view must start on a 32-byte boundary (row strips start on 128-cell boundaries)");
    if (is_sharded(b)) return sh_view(b, offset_cells, len, out);
    EC_TRY(resolve(b));
    ec_buf* v = new ec_buf;
    v->ct = b->ct; v->owned = b->owned; v->len = len; v->capacity_bytes = len * kSize[b->ct]; v->dev = b->dev;
    v->dptr = static_cast<char*>(b->dptr) + offset_cells * kSize[b->ct];
    v->blk = b->blk;
    if (v->blk) { v->blk->views.fetch_add(1); v->is_view = true; }
    const_cast<ec_buf*>(b)->mm_known = false;  // the cells can now change behind b's back
    if (b->ready) cudaStreamWaitEvent(device_stream(b->dev), b->ready, 0);  // the view has no event of its own: order it after the upload now
    *out = v;
    return EC_OK;
}
ec_status ec_buf_clone(const ec_buf* b, ec_buf** out) {
    EC_TRY(ensure());
    if (is_sharded(b)) return sh_map1(b, b->ct, [](const ec_buf* part, ec_buf** o) { return ec_buf_clone(part, o); }, out);
    EC_TRY(resolve(b));
    ec_buf* c;
    EC_TRY(new_buf(b->ct, b->len, &c));
    if (b->len) {
        if (cudaError_t e = launch_copy(launch_ctx(), (int)kSize[b->ct], rd(b), c->dptr, b->len)) { ec_buf_free(c); return cuda_fail(e, "clone"); }
        note_launch("clone");
    }
    *out = c;
    return EC_OK;
}
void ec_buf_free(ec_buf* b) {
    if (!b) return;
    for (ec_buf* part : b->parts) ec_buf_free(part);
    if (b->is_view && b->blk) b->blk->views.fetch_sub(1);
    if (b->ready) {
        cudaEventSynchronize(b->ready);
        cudaEventDestroy(b->ready);
    }
    delete b;  // the block goes back to the allocator when its last user (this handle, a view or a pending Expr) lets go
}
size_t ec_buf_len(const ec_buf* b) { return b->len; }
uint8_t ec_buf_ctype(const ec_buf* b) { return b->ct; }
void* ec_buf_device_ptr(const ec_buf* b) {
    if (is_sharded(b)) { set_error("a sharded buffer has one device pointer per strip: ec_buf_shard()"); return nullptr; }
    return resolve(b) == EC_OK ? b->dptr : nullptr;
}
ec_status ec_buf_to_host(const ec_buf* b, void* host, size_t host_bytes) {
    EC_TRY(ensure());
    const size_t bytes = b->len * kSize[b->ct];
    if (host_bytes < bytes) return invalid("ec_buf_to_host: host buffer too small");
    if (is_sharded(b)) return sh_to_host(b, host);
    DevScope on(b->dev);
    EC_TRY(resolve(b));
    if (bytes && staged_wanted(host, bytes)) {
        const HostSeg seg{b->dev, const_cast<void*>(rd(b)), host, bytes, cur_stream()};
        return staged_transfer(&seg, 1, false);
    }
    if (bytes) EC_CUDA_TRY(cudaMemcpyAsync(host, rd(b), bytes, cudaMemcpyDeviceToHost, cur_stream()), "cudaMemcpyAsync(D2H)");
    return sync_stream();
}
ec_status ec_buf_get(const ec_buf* b, size_t index, ec_value* out) {
    EC_TRY(ensure());
    if (index >= b->len) { set_error("index out of bounds: the len is %zu but the index is %zu", b->len, index); return EC_OOB; }
    if (is_sharded(b)) { const int g = sh_part_of(b->offs, index); DevScope on(g); return ec_buf_get(b->parts[g], index - b->offs[g], out); }
    DevScope on(b->dev);
    EC_TRY(resolve(b));
    uint64_t* pin;
    EC_TRY(pinned_words(&pin));
    pin[0] = 0;
    EC_CUDA_TRY(cudaMemcpyAsync(pin, static_cast<const char*>(rd(b)) + index * kSize[b->ct], kSize[b->ct], cudaMemcpyDeviceToHost, cur_stream()), "cudaMemcpyAsync(D2H)");
    EC_TRY(sync_stream());
    *out = tagged<uint64_t>(b->ct, pin[0]);
    return EC_OK;
}
ec_status ec_buf_put(ec_buf* b, size_t index, const ec_value* value) {
    EC_TRY(ensure());
    if (!ct_ok(value->ct)) return invalid("cell type");
    if (!ct_fits(value->ct, b->ct)) return narrowing(value->ct, b->ct);  // convert()? happens before the index (src/buffer.rs:137)
    if (index >= b->len) { set_error("index out of bounds: the len is %zu but the index is %zu", b->len, index); return EC_OOB; }
    b->mm_known = false;
    if (is_sharded(b)) { const int g = sh_part_of(b->offs, index); DevScope on(g); return ec_buf_put(b->parts[g], index - b->offs[g], value); }
    DevScope on(b->dev);
    const ec_value c = value_widen(*value, b->ct);
    EC_TRY(resolve(b));
    EC_TRY(ensure_unique(b));
    uint64_t* pin;
    EC_TRY(pinned_words(&pin));
    pin[1] = c.bits;
    EC_CUDA_TRY(cudaMemcpyAsync(static_cast<char*>(rd(b)) + index * kSize[b->ct], &pin[1], kSize[b->ct], cudaMemcpyHostToDevice, cur_stream()), "cudaMemcpyAsync(H2D)");
    return sync_stream();
}
ec_status ec_buf_extend_host(ec_buf* b, uint8_t ct, const void* host, size_t n) {
    EC_TRY(ensure());
    if (!ct_ok(ct)) return invalid("cell type");
    if (!b->owned) return invalid("cannot extend a wrapped buffer");
    if (n == 0) return EC_OK;
    b->mm_known = false;
    if (is_sharded(b)) {  // the appended cells join the last strip
        const int g = static_cast<int>(b->parts.size()) - 1;
        DevScope on(g);
        EC_TRY(ec_buf_extend_host(b->parts[g], ct, host, n));
        b->len += n;
        b->offs.back() = b->len;
        b->capacity_bytes = b->len * kSize[b->ct];
        return EC_OK;
    }
    DevScope on(b->dev);
    EC_TRY(resolve(b));
    const size_t new_len = b->len + n, sz = kSize[b->ct];
    Scratch grown;
    EC_TRY(grown.alloc(new_len * sz));
    if (b->len) EC_CUDA_TRY(cudaMemcpyAsync(grown.p, rd(b), b->len * sz, cudaMemcpyDeviceToDevice, cur_stream()), "cudaMemcpyAsync(D2D)");
    char* dst = static_cast<char*>(grown.p) + b->len * sz;
    if (ct == b->ct && ct != EC_FLOAT32) {  // same type: to_<p>() is the identity (f32 NaNs still pass through f64, below)
        EC_CUDA_TRY(cudaMemcpyAsync(dst, host, n * sz, cudaMemcpyHostToDevice, cur_stream()), "cudaMemcpyAsync(H2D)");
        EC_TRY(sync_stream());
    } else {
        // `c.into_cell_value().to_<p>().unwrap()` (src/buffer.rs:212): value-checked on the device
        Scratch stage, conv;  // the appended run starts at an arbitrary cell offset: cast into an aligned temp
        EC_TRY(stage.alloc(n * kSize[ct]));
        EC_TRY(conv.alloc(n * sz));
        StreamScratch* sc;
        EC_TRY(stream_scratch(&sc));
        EC_CUDA_TRY(cudaMemcpyAsync(stage.p, host, n * kSize[ct], cudaMemcpyHostToDevice, cur_stream()), "cudaMemcpyAsync(H2D)");
        EC_CUDA_TRY(cudaMemsetAsync(sc->rs.result, 0, 8, cur_stream()), "cudaMemsetAsync");
        EC_LAUNCH(launch_checked_cast(launch_ctx(), ct, stage.p, b->ct, conv.p, n, reinterpret_cast<unsigned int*>(sc->rs.result)), "checked_cast");
        EC_CUDA_TRY(cudaMemcpyAsync(dst, conv.p, n * sz, cudaMemcpyDeviceToDevice, cur_stream()), "cudaMemcpyAsync(D2D)");
        uint64_t* pin;
        EC_TRY(pinned_words(&pin));
        EC_CUDA_TRY(cudaMemcpyAsync(pin, sc->rs.result, 8, cudaMemcpyDeviceToHost, cur_stream()), "cudaMemcpyAsync(D2H)");
        EC_TRY(sync_stream());  // `host` may be reused by the caller as soon as we return
        if ((pin[0] & 0xFFFFFFFFu) != 0) return narrowing(ct, b->ct);  // a cell did not fit: the reference panics in `unwrap()`; the buffer is left untouched
    }
    void* p = grown.release();
    b->blk = own_block(p);  // the old block returns to the allocator once no view or pending Expr needs it
    b->dptr = p;
    b->len = new_len;
    b->capacity_bytes = new_len * sz;
    return EC_OK;
}

// an empty result is UInt8([]) — FromIterator<CellValue>, src/buffer.rs:233-236
static ec_status empty_result(ec_buf** out) { return new_buf(EC_UINT8, 0, out); }

static ec_status pending_result(std::shared_ptr<Expr> e, ec_buf** out) {
    e->depth = 1 + std::max(depth_of(e->l), depth_of(e->r));
    *out = pending_buf(std::move(e));
    return EC_OK;
}
ec_status ec_buf_binary(int op, const ec_buf* l, const ec_buf* r, ec_buf** out) {
    EC_TRY(ensure());
    if (op < 0 || op > 3) return invalid("op");
    const size_t n = std::min(l->len, r->len);  // zip (src/buffer.rs:327)
    if (n == 0) return empty_result(out);
    if (is_sharded(l) || is_sharded(r) || shard_policy(n) || l->dev != t_dev || r->dev != t_dev)
        return sh_map2(l, r, n, EC_FLOAT64, [op](const ec_buf* a, const ec_buf* b, ec_buf** o) { return ec_buf_binary(op, a, b, o); }, out);
    if (t_lazy && lazy_capable(l) && lazy_capable(r)) {
        auto e = std::make_shared<Expr>();
        e->kind = EX_BIN; e->op = op; e->s = 0; e->n = n;
        EC_TRY(snapshot(l, &e->l));
        EC_TRY(snapshot(r, &e->r));
        return pending_result(std::move(e), out);
    }
    EC_TRY(resolve(l));
    EC_TRY(resolve(r));
    ec_buf* o;
    EC_TRY(new_buf(EC_FLOAT64, n, &o));
    if (cudaError_t e = launch_binary(launch_ctx(), op, l->ct, rd(l), r->ct, rd(r), static_cast<double*>(o->dptr), n, nullptr, nullptr, nullptr, kNoCount)) {
        ec_buf_free(o);
        return cuda_fail(e, "binary");
    }
    note_launch("binary");
    *out = o;
    return EC_OK;
}
ec_status ec_buf_scalar(int op, const ec_buf* l, const ec_value* r, ec_buf** out) {
    EC_TRY(ensure());
    if (op < 0 || op > 3 || !ct_ok(r->ct)) return invalid("op / cell type");
    if (l->len == 0) return empty_result(out);
    if (is_sharded(l) || l->dev != t_dev) return sh_map1(l, EC_FLOAT64, [op, r](const ec_buf* part, ec_buf** o) { return ec_buf_scalar(op, part, r, o); }, out);
    // unify() is value-exact, so the rhs the reference feeds to the f64 op is `r as f64`
    const double s = value_as_f64(*r);
    if (t_lazy && lazy_capable(l)) {
        auto e = std::make_shared<Expr>();
        e->kind = EX_SCALAR; e->op = op; e->s = s; e->n = l->len;
        EC_TRY(snapshot(l, &e->l));
        return pending_result(std::move(e), out);
    }
    EC_TRY(resolve(l));
    ec_buf* o;
    EC_TRY(new_buf(EC_FLOAT64, l->len, &o));
    if (cudaError_t e = launch_scalar(launch_ctx(), op, l->ct, rd(l), s, static_cast<double*>(o->dptr), l->len)) {
        ec_buf_free(o);
        return cuda_fail(e, "scalar");
    }
    note_launch("scalar");
    *out = o;
    return EC_OK;
}
static const uint8_t kNegOut[10] = {EC_INT16, EC_INT32, EC_FLOAT64, EC_FLOAT64, EC_INT8, EC_INT16, EC_INT32, EC_INT64, EC_FLOAT32, EC_FLOAT64};
ec_status ec_buf_neg(const ec_buf* b, ec_buf** out) {
    EC_TRY(ensure());
    if (b->len == 0) return empty_result(out);
    if (is_sharded(b) || b->dev != t_dev) return sh_map1(b, kNegOut[b->ct], [](const ec_buf* part, ec_buf** o) { return ec_buf_neg(part, o); }, out);
    EC_TRY(resolve(b));
    ec_buf* o;
    EC_TRY(new_buf(kNegOut[b->ct], b->len, &o));
    if (cudaError_t e = launch_neg(launch_ctx(), b->ct, rd(b), o->dptr, b->len)) { ec_buf_free(o); return cuda_fail(e, "neg"); }
    note_launch("neg");
    *out = o;
    return EC_OK;
}
ec_status ec_buf_convert(const ec_buf* b, uint8_t ct, ec_buf** out) {
    EC_TRY(ensure());
    if (!ct_ok(ct)) return invalid("cell type");
    if (ct == b->ct) return ec_buf_clone(b, out);            // src/buffer.rs:151-153
    if (!ct_fits(b->ct, ct)) return narrowing(b->ct, ct);    // src/buffer.rs:155-159, before any launch
    if (b->len == 0) return empty_result(out);               // collect() of nothing, src/buffer.rs:234
    if (is_sharded(b) || b->dev != t_dev) return sh_map1(b, ct, [ct](const ec_buf* part, ec_buf** o) { return ec_buf_convert(part, ct, o); }, out);
    EC_TRY(resolve(b));
    ec_buf* o;
    EC_TRY(new_buf(ct, b->len, &o));
    if (cudaError_t e = launch_convert(launch_ctx(), b->ct, rd(b), ct, o->dptr, b->len)) { ec_buf_free(o); return cuda_fail(e, "convert"); }
    note_launch("convert");
    *out = o;
    return EC_OK;
}
static bool g_mm_cache = env_int("EC_MIN_MAX_CACHE", 1) != 0;
static bool mm_cacheable(const ec_buf* b) {  // nothing is remembered about cells another handle can change
    if (!g_mm_cache) return false;
    if (b->blk && b->blk->views.load() != 0) return false;
    for (const ec_buf* part : b->parts)
        if (!mm_cacheable(part)) return false;
    return b->owned;
}
int ec_set_min_max_cache(int on) { const int prev = g_mm_cache; g_mm_cache = on != 0; return prev; }
static ec_value key_value(uint8_t ct, uint64_t key) { return tagged<uint64_t>(ct, key_to_bits(ct, key)); }
ec_status ec_buf_min_max(const ec_buf* b, const ec_mask* m, ec_value* mn, ec_value* mx) {
    EC_TRY(ensure());
    if (m && m->len != b->len) { set_error("Mask and buffer must have the same length."); return EC_LEN_MISMATCH; }
    uint64_t k[2];
    key_seeds(b->ct, &k[0], &k[1]);
    const bool cacheable = !m && mm_cacheable(b);
    if (cacheable && b->mm_known) {
        k[0] = b->mm_k0; k[1] = b->mm_k1;
    } else if (is_sharded(b) || (m && is_sharded(m))) {
        EC_TRY(sh_min_max(b, m, &k[0], &k[1]));
    } else if (b->len) {
        DevScope on(b->dev);
        PendingReduce pend;
        EC_TRY(min_max_begin(b, m, nullptr, &pend, nullptr));
        EC_TRY(reduce_end(pend, &k[0], &k[1]));
    }
    if (cacheable && !b->mm_known) {  // a cache: concurrent readers store the same values
        ec_buf* w = const_cast<ec_buf*>(b);
        w->mm_k0 = k[0]; w->mm_k1 = k[1];
        __atomic_store_n(&w->mm_known, true, __ATOMIC_RELEASE);
    }
    *mn = key_value(b->ct, k[0]);
    *mx = key_value(b->ct, k[1]);
    return EC_OK;
}
// ---- statistics extension (no reference counterpart): definition in ec_stats.cuh / DESIGN.md §4.6 ----
ec_status ec_statistics_plan(const ec_value* mn, const ec_value* mx, int* kind, double* pivot, int* exp2) {
    if (!ct_ok(mn->ct) || mn->ct != mx->ct) return invalid("ec_statistics_plan: min and max must share one cell type");
    *pivot = 0.0;
    *exp2 = 0;
    if (value_cmp(*mn, *mx) > 0) { *kind = EC_STATS_EMPTY; return EC_OK; }  // the (T::MAX, T::MIN) seeds survived
    const double lo = value_as_f64(*mn), hi = value_as_f64(*mx);
    if (!std::isfinite(lo) || !std::isfinite(hi)) { *kind = EC_STATS_NONFINITE; return EC_OK; }
    if (mn->ct == EC_FLOAT32) {  // quantised route: no pivot on the device, exp2 = E with every |cell| < 2^E
        const double top = std::fabs(lo) > std::fabs(hi) ? std::fabs(lo) : std::fabs(hi);
        int e = 0;
        if (top != 0) (void)std::frexp(top, &e);
        *kind = EC_STATS_REGULAR;
        *exp2 = e;
        return EC_OK;
    }
    const volatile double half_lo = lo * 0.5, half_hi = hi * 0.5;  // one rounding per step, as the oracle states it
    const double p = half_lo + half_hi;
    const double up = hi - p, down = p - lo;
    const double d = up > down ? up : down;
    int e = 0;
    if (d != 0) (void)std::frexp(d, &e);
    *kind = EC_STATS_REGULAR;
    *pivot = p;
    *exp2 = e < -1000 ? -1000 : (e > 1024 ? 1024 : e);
    return EC_OK;
}
// Integer cells of at most 32 bits take the exact integer route: raw = {count, A = sum x, B = sum x^2, 0, 0}
// (pivot-free), everything else the FP64 window route: raw = {count, X1, X2, Z1, Z2}.
ec_status ec_buf_moments(const ec_buf* b, const ec_mask* m, double pivot, int exp2, uint64_t* raw) {
    EC_TRY(ensure());
    if (m && m->len != b->len) { set_error("Mask and buffer must have the same length."); return EC_LEN_MISMATCH; }
    if (exp2 < -1000 || exp2 > 1024) return invalid("ec_buf_moments: exp2 out of range");
    memset(raw, 0, EC_MOMENT_WORDS * sizeof(uint64_t));
    if (b->len == 0) return EC_OK;
    if (is_sharded(b) || (m && is_sharded(m))) return sh_moments(b, m, pivot, exp2, raw);
    DevScope on(b->dev);
    StatsPending pend;
    uint64_t w[EC_MOMENT_WORDS];
    if (stats_integer_route(b->ct) || b->ct == EC_FLOAT32) {
        if (b->ct == EC_FLOAT32) EC_TRY(quant_stats_begin(b, m, exp2, &pend));
        else EC_TRY(int_stats_begin(b, m, &pend));
        EC_TRY(stats_end(pend, w));
        memcpy(raw, w, 5 * sizeof(uint64_t));
        return EC_OK;
    }
    EC_TRY(moments_begin(b, m, pivot, exp2, &pend));
    return stats_end(pend, raw);
}
ec_status ec_statistics_finish(const uint64_t* raws, size_t n_parts, const ec_value* mn, const ec_value* mx, ec_statistics* out) {
    int kind, e;
    double p;
    EC_TRY(ec_statistics_plan(mn, mx, &kind, &p, &e));
    unsigned __int128 tot[4] = {0, 0, 0, 0};  // two's complement, modulo 2^128
    uint64_t count = 0;
    for (size_t i = 0; i < n_parts; ++i) {
        const uint64_t* r = raws + i * EC_MOMENT_WORDS;
        count += r[0];
        for (int k = 0; k < 4; ++k) tot[k] += (static_cast<unsigned __int128>(r[2 + 2 * k]) << 64) | r[1 + 2 * k];
    }
    const double qnan = std::numeric_limits<double>::quiet_NaN();
    out->count = count;
    out->min = *mn;
    out->max = *mx;
    out->mean = out->stddev = qnan;
    if (kind == EC_STATS_EMPTY || count == 0) return EC_OK;
    if (kind == EC_STATS_NONFINITE) {
        const double lo = value_as_f64(*mn), hi = value_as_f64(*mx);
        if (!std::isnan(lo) && !std::isnan(hi) && !(std::isinf(lo) && std::isinf(hi))) out->mean = std::isinf(lo) ? lo : hi;
        return EC_OK;
    }
    if (mn->ct == EC_FLOAT32) {
        // quantised route: the sums are those of the int32 raster q = rint(x * 2^k), k = 26 - E. Finish that integer
        // raster (min / max quantise monotonically) and scale the result back by 2^-k (exact).
        const int k = 26 - e;
        const ec_value qmn = tagged<int32_t>(EC_INT32, static_cast<int32_t>(std::nearbyint(std::ldexp(value_as_f64(*mn), k))));
        const ec_value qmx = tagged<int32_t>(EC_INT32, static_cast<int32_t>(std::nearbyint(std::ldexp(value_as_f64(*mx), k))));
        ec_statistics q;
        EC_TRY(ec_statistics_finish(raws, n_parts, &qmn, &qmx, &q));
        out->mean = std::ldexp(q.mean, -k);
        out->stddev = std::ldexp(q.stddev, -k);
        return EC_OK;
    }
    // one rounding per operation (host code is built with -ffp-contract=off; int128 -> double is round-to-nearest-even)
    volatile double s1, s2;
    if (stats_integer_route(mn->ct)) {
        // y = x - pivot = Y/2 with Y = 2x - s, s = 2 * pivot = min + max (an integer): the sums of y' = y * 2^-E and
        // of y'^2 are exact rationals, rounded once
        const __int128 cnt = static_cast<__int128>(count), A = static_cast<__int128>(tot[0]), B = static_cast<__int128>(tot[1]);
        const __int128 s = static_cast<__int128>(p * 2.0);
        s1 = std::ldexp(static_cast<double>(2 * A - cnt * s), -1 - e);
        s2 = std::ldexp(static_cast<double>(4 * B - 4 * s * A + cnt * s * s), -2 - 2 * e);
    } else {
        volatile double f[4];
        for (int k = 0; k < 4; ++k) f[k] = static_cast<double>(static_cast<__int128>(tot[k]));
        s1 = std::ldexp(f[0], -47) + std::ldexp(f[1], -95);
        s2 = std::ldexp(f[2], -47) + std::ldexp(f[3], -95);
    }
    const double n = static_cast<double>(count);
    const volatile double m1 = s1 / n, m2 = s2 / n;
    const volatile double sq = m1 * m1;
    double var = m2 - sq;
    if (var < 0) var = 0.0;
    out->mean = p + std::ldexp(m1, e);
    out->stddev = std::ldexp(std::sqrt(var), e);
    return EC_OK;
}
ec_status ec_buf_statistics(const ec_buf* b, const ec_mask* m, ec_statistics* out) {
    EC_TRY(ensure());
    if (m && m->len != b->len) { set_error("Mask and buffer must have the same length."); return EC_LEN_MISMATCH; }
    if (is_sharded(b) || (m && is_sharded(m))) return sh_statistics(b, m, out);
    DevScope on(b->dev);
    ec_value mn, mx;
    uint64_t raw[EC_MOMENT_WORDS] = {0};
    if (stats_integer_route(b->ct) && b->len) {
        // one pass: min, max, count and both integer sums; the pivot is only needed by the finish
        uint64_t w[8];
        StatsPending pend;
        EC_TRY(int_stats_begin(b, m, &pend));
        EC_TRY(stats_end(pend, w));
        memcpy(raw, w, 5 * sizeof(uint64_t));
        int_stats_min_max(b->ct, w, &mn, &mx);
        return ec_statistics_finish(raw, 1, &mn, &mx, out);
    }
    EC_TRY(ec_buf_min_max(b, m, &mn, &mx));
    int kind, e;
    double p;
    EC_TRY(ec_statistics_plan(&mn, &mx, &kind, &p, &e));
    if (kind == EC_STATS_REGULAR) {
        EC_TRY(ec_buf_moments(b, m, p, e, raw));
    } else if (m) {  // nothing to sum, but the count of valid cells is still reported
        size_t d, nd;
        EC_TRY(ec_mask_counts(m, &d, &nd));
        raw[0] = d;
    } else {
        raw[0] = b->len;
    }
    return ec_statistics_finish(raw, 1, &mn, &mx, out);
}
ec_status ec_buf_cmp(const ec_buf* l, const ec_buf* r, int* ordering) {
    EC_TRY(ensure());
    if (l->ct != r->ct) { *ordering = l->ct < r->ct ? -1 : 1; return EC_OK; }  // src/buffer.rs:395-398
    const size_t n = std::min(l->len, r->len);
    if (n) {
        uint64_t idx = ~0ull;
        if (is_sharded(l) || is_sharded(r) || l->dev != r->dev) {
            EC_TRY(sh_first_diff(l, r, n, &idx));
        } else {
            DevScope on(l->dev);
            PendingReduce pend;
            EC_TRY(first_diff_begin(l, r, n, &pend));
            uint64_t unused;
            EC_TRY(reduce_end(pend, &idx, &unused));
        }
        if (idx != ~0ull) {
            ec_value a, b;
            EC_TRY(ec_buf_get(l, idx, &a));
            EC_TRY(ec_buf_get(r, idx, &b));
            *ordering = value_cmp(a, b);
            return EC_OK;
        }
    }
    *ordering = l->len < r->len ? -1 : (l->len > r->len ? 1 : 0);
    return EC_OK;
}

// ---- fused chains --------------------------------------------------------------------------------------
ec_status ec_buf_normalized_difference(const ec_buf* a, const ec_buf* b, ec_buf** out) {
    EC_TRY(ensure());
    const size_t n = std::min(a->len, b->len);
    if (n == 0) return empty_result(out);
    if (is_sharded(a) || is_sharded(b) || shard_policy(n) || a->dev != t_dev || b->dev != t_dev)
        return sh_map2(a, b, n, EC_FLOAT64, [](const ec_buf* x, const ec_buf* y, ec_buf** o) { return ec_buf_normalized_difference(x, y, o); }, out);
    EC_TRY(resolve(a));
    EC_TRY(resolve(b));
    ec_buf* o;
    EC_TRY(new_buf(EC_FLOAT64, n, &o));
    if (cudaError_t e = launch_normdiff(launch_ctx(), a->ct, rd(a), b->ct, rd(b), static_cast<double*>(o->dptr), n)) { ec_buf_free(o); return cuda_fail(e, "normdiff"); }
    note_launch("normalized_difference");
    *out = o;
    return EC_OK;
}
ec_status ec_buf_binary_scalar(int op1, const ec_buf* l, const ec_buf* r, int op2, const ec_value* s, ec_buf** out) {
    EC_TRY(ensure());
    if (op1 < 0 || op1 > 3 || op2 < 0 || op2 > 3 || !ct_ok(s->ct)) return invalid("op / cell type");
    const size_t n = std::min(l->len, r->len);
    if (n == 0) return empty_result(out);
    if (is_sharded(l) || is_sharded(r) || shard_policy(n) || l->dev != t_dev || r->dev != t_dev)
        return sh_map2(l, r, n, EC_FLOAT64, [=](const ec_buf* x, const ec_buf* y, ec_buf** o) { return ec_buf_binary_scalar(op1, x, y, op2, s, o); }, out);
    EC_TRY(resolve(l));
    EC_TRY(resolve(r));
    ec_buf* o;
    EC_TRY(new_buf(EC_FLOAT64, n, &o));
    if (cudaError_t e = launch_binary_scalar(launch_ctx(), op1, l->ct, rd(l), r->ct, rd(r), op2, value_as_f64(*s), static_cast<double*>(o->dptr), n)) { ec_buf_free(o); return cuda_fail(e, "binary_scalar"); }
    note_launch("binary_scalar");
    *out = o;
    return EC_OK;
}

// ---- Mask ------------------------------------------------------------------------------------------------
using MaskOwner = std::unique_ptr<ec_mask, void (*)(ec_mask*)>;  // a half-built result: an early error return frees it
ec_status ec_mask_from_bools(const uint8_t* host_bools, size_t len, ec_mask** out) {
    EC_TRY(ensure());
    if (shard_policy(len)) return shm_generate(len, [&](size_t off, size_t n, ec_mask** o) { return ec_mask_from_bools(host_bools + off, n, o); }, out);
    ec_mask* m;
    EC_TRY(new_mask(len, &m));
    MaskOwner hold(m, ec_mask_free);
    if (len) {
        Scratch stage;
        EC_TRY(stage.alloc(len));
        if (staged_wanted(host_bools, len)) {  // a Vec<bool>: the staged copy, on the stream the pack kernel follows on
            const HostSeg seg{t_dev, stage.p, const_cast<uint8_t*>(host_bools), len, cur_stream()};
            EC_TRY(staged_transfer(&seg, 1, true));
        } else {
            EC_CUDA_TRY(cudaMemcpyAsync(stage.p, host_bools, len, cudaMemcpyHostToDevice, cur_stream()), "cudaMemcpyAsync(H2D)");
        }
        EC_LAUNCH(launch_mask_build(launch_ctx(), 1, stage.p, len, 0, true, m->words, arm_count(m)), "mask_pack");
    }
    *out = hold.release();
    return EC_OK;
}
ec_status ec_mask_fill(size_t len, int value, ec_mask** out) {
    EC_TRY(ensure());
    if (shard_policy(len)) return shm_generate(len, [&](size_t, size_t n, ec_mask** o) { return ec_mask_fill(n, value, o); }, out);
    ec_mask* m;
    EC_TRY(new_mask(len, &m));
    MaskOwner hold(m, ec_mask_free);
    if (len) EC_LAUNCH(launch_mask_fill(launch_ctx(), m->words, len, value != 0), "mask_fill");
    m->ones = value ? len : 0;
    m->cnt_known = true;
    *out = hold.release();
    return EC_OK;
}
ec_status ec_mask_to_bools(const ec_mask* m, uint8_t* host_bools, size_t capacity) {
    EC_TRY(ensure());
    if (capacity < m->len) return invalid("ec_mask_to_bools: host buffer too small");
    if (m->len == 0) return EC_OK;
    if (is_sharded(m)) {
        for (size_t g = 0; g < m->parts.size(); ++g) EC_TRY(ec_mask_to_bools(m->parts[g], host_bools + m->offs[g], m->parts[g]->len));
        return EC_OK;
    }
    DevScope on(m->dev);
    Scratch stage;
    EC_TRY(stage.alloc(m->len));
    EC_LAUNCH(launch_mask_unpack(launch_ctx(), rdm(m), m->len, static_cast<uint8_t*>(stage.p)), "mask_unpack");
    if (staged_wanted(host_bools, m->len)) {
        const HostSeg seg{m->dev, stage.p, host_bools, m->len, cur_stream()};
        return staged_transfer(&seg, 1, false);
    }
    EC_CUDA_TRY(cudaMemcpyAsync(host_bools, stage.p, m->len, cudaMemcpyDeviceToHost, cur_stream()), "cudaMemcpyAsync(D2H)");
    return sync_stream();
}
// Clone shares the words (refcounted): a Mask is immutable except through put / extend, which copy first. This is what
// makes the mask of `masked * scalar` and `-masked` (src/masked/masked_buffer.rs:353-364, :372-383) free.
ec_status ec_mask_clone(const ec_mask* m, ec_mask** out) {
    EC_TRY(ensure());
    ec_mask* c = new ec_mask;
    c->len = m->len; c->capacity_bytes = m->capacity_bytes; c->words = m->words; c->blk = m->blk; c->dev = m->dev;
    c->cnt = m->cnt; c->cnt_seq = m->cnt_seq; c->cnt_known = m->cnt_known; c->ones = m->ones;
    c->offs = m->offs;
    for (const ec_mask* part : m->parts) {
        ec_mask* pc = nullptr;
        if (ec_status s = ec_mask_clone(part, &pc)) { ec_mask_free(c); return s; }
        c->parts.push_back(pc);
    }
    *out = c;
    return EC_OK;
}
ec_status ec_mask_slice(const ec_mask* m, size_t offset_cells, size_t len, ec_mask** out) {
    EC_TRY(ensure());
    if (offset_cells > m->len || len > m->len - offset_cells) { set_error("slice [%zu, %zu) outside a mask of %zu cells", offset_cells, offset_cells + len, m->len); return EC_OOB; }
    if (offset_cells % 128 != 0) return invalid("a mask slice must start on a 128-cell boundary (row strips do)");
    if (is_sharded(m)) return shm_slice(m, offset_cells, len, out);
    DevScope on(m->dev);
    ec_mask* o;
    EC_TRY(new_mask(len, &o));
    MaskOwner hold(o, ec_mask_free);
    // x & x with the strip's words as both operands: the bit-op kernel already clears the bits past `len` in the last word
    const uint32_t* src = rdm(m) + offset_cells / 32;
    if (len) EC_LAUNCH(launch_mask_bitop(launch_ctx(), 1, src, src, len, o->words, arm_count(o)), "mask_slice");
    *out = hold.release();
    return EC_OK;
}
void ec_mask_free(ec_mask* m) {
    if (!m) return;
    for (ec_mask* part : m->parts) ec_mask_free(part);
    delete m;  // the words return to the allocator when the last clone lets go
}
size_t ec_mask_len(const ec_mask* m) { return m->len; }
void* ec_mask_device_words(const ec_mask* m) {
    if (is_sharded(m)) { set_error("a sharded mask has one word array per strip: ec_mask_shard()"); return nullptr; }
    return m->words;
}
static ec_status mask_word(const ec_mask* m, size_t w, uint32_t* out) {
    uint64_t* pin;
    EC_TRY(pinned_words(&pin));
    EC_CUDA_TRY(cudaMemcpyAsync(pin, rdm(m) + w, 4, cudaMemcpyDeviceToHost, cur_stream()), "cudaMemcpyAsync(D2H)");
    EC_TRY(sync_stream());
    *out = static_cast<uint32_t>(pin[0]);
    return EC_OK;
}
ec_status ec_mask_get(const ec_mask* m, size_t index, int* out) {
    EC_TRY(ensure());
    if (index >= m->len) { set_error("index out of bounds: the len is %zu but the index is %zu", m->len, index); return EC_OOB; }
    if (is_sharded(m)) { const int g = sh_part_of(m->offs, index); return ec_mask_get(m->parts[g], index - m->offs[g], out); }
    DevScope on(m->dev);
    uint32_t w;
    EC_TRY(mask_word(m, index / 32, &w));
    *out = (w >> (index % 32)) & 1u;
    return EC_OK;
}
ec_status ec_mask_put(ec_mask* m, size_t index, int value) {
    EC_TRY(ensure());
    if (index >= m->len) { set_error("index out of bounds: the len is %zu but the index is %zu", m->len, index); return EC_OOB; }
    if (is_sharded(m)) { const int g = sh_part_of(m->offs, index); return ec_mask_put(m->parts[g], index - m->offs[g], value); }
    DevScope on(m->dev);
    if (m->cnt_seq != 0 && !m->cnt_known) { uint64_t unused; EC_TRY(mask_ones(m, &unused)); }  // settle the producer's count before it goes stale
    EC_TRY(mask_unique(m));
    uint32_t w;
    EC_TRY(mask_word(m, index / 32, &w));
    const uint32_t bit = 1u << (index % 32);
    const uint32_t nw = value ? (w | bit) : (w & ~bit);
    if (m->cnt_known) m->ones += __builtin_popcount(nw) - __builtin_popcount(w);  // the cached count follows the mutation
    m->cnt_seq = 0;
    uint64_t* pin;
    EC_TRY(pinned_words(&pin));
    pin[1] = nw;
    EC_CUDA_TRY(cudaMemcpyAsync(m->words + index / 32, &pin[1], 4, cudaMemcpyHostToDevice, cur_stream()), "cudaMemcpyAsync(H2D)");
    return sync_stream();
}
ec_status ec_mask_extend_host(ec_mask* m, const uint8_t* host_bools, size_t n) {
    EC_TRY(ensure());
    if (n == 0) return EC_OK;
    if (is_sharded(m)) {
        EC_TRY(ec_mask_extend_host(m->parts.back(), host_bools, n));
        m->len += n;
        m->offs.back() = m->len;
        return EC_OK;
    }
    DevScope on(m->dev);
    // unpack -> append -> repack on the device (Extend is not a bulk path in the reference either)
    const size_t new_len = m->len + n;
    Scratch stage, words;
    EC_TRY(stage.alloc(new_len));
    if (m->len) EC_LAUNCH(launch_mask_unpack(launch_ctx(), rdm(m), m->len, static_cast<uint8_t*>(stage.p)), "mask_unpack");
    EC_CUDA_TRY(cudaMemcpyAsync(static_cast<uint8_t*>(stage.p) + m->len, host_bools, n, cudaMemcpyHostToDevice, cur_stream()), "cudaMemcpyAsync(H2D)");
    EC_TRY(words.alloc(mask_bytes(new_len)));
    m->len = new_len;
    m->cnt = nullptr;
    EC_LAUNCH(launch_mask_build(launch_ctx(), 1, stage.p, new_len, 0, true, static_cast<uint32_t*>(words.p), arm_count(m)), "mask_pack");
    EC_TRY(sync_stream());  // `host_bools` may be reused by the caller as soon as we return
    m->words = static_cast<uint32_t*>(words.release());
    m->blk = own_block(m->words);
    m->capacity_bytes = mask_bytes(new_len);
    return EC_OK;
}
static ec_status mask_bitop(int mop, const ec_mask* l, const ec_mask* r, ec_mask** out) {
    EC_TRY(ensure());
    const size_t n = r ? std::min(l->len, r->len) : l->len;  // zip (src/masked/mask.rs:133-137)
    if (is_sharded(l) || (r && is_sharded(r)) || shard_policy(n) || l->dev != t_dev || (r && r->dev != t_dev))
        return shm_map2(l, r, n, [mop](const ec_mask* a, const ec_mask* b, ec_mask** o) { return mask_bitop(mop, a, b, o); }, out);
    ec_mask* o;
    EC_TRY(new_mask(n, &o));
    MaskOwner hold(o, ec_mask_free);
    if (n) EC_LAUNCH(launch_mask_bitop(launch_ctx(), mop, rdm(l), r ? rdm(r) : nullptr, n, o->words, arm_count(o)), "mask_bitop");
    *out = hold.release();
    return EC_OK;
}
ec_status ec_mask_not(const ec_mask* m, ec_mask** out) { return mask_bitop(0, m, nullptr, out); }
ec_status ec_mask_and(const ec_mask* l, const ec_mask* r, ec_mask** out) { return mask_bitop(1, l, r, out); }
ec_status ec_mask_or(const ec_mask* l, const ec_mask* r, ec_mask** out) { return mask_bitop(2, l, r, out); }
ec_status ec_mask_counts(const ec_mask* m, size_t* data, size_t* nodata) {
    EC_TRY(ensure());
    uint64_t ones = 0;
    if (is_sharded(m)) {  // strips that still need a popcount pass run it side by side
        const size_t G = m->parts.size();
        std::vector<PendingReduce> pend(G);
        std::vector<char> launched(G, 0);
        for (size_t g = 0; g < G; ++g) {
            DevScope on(m->parts[g]->dev);
            bool l = false;
            EC_TRY(mask_ones_begin(m->parts[g], &pend[g], &l));
            launched[g] = l;
        }
        for (size_t g = 0; g < G; ++g) {
            DevScope on(m->parts[g]->dev);
            uint64_t c = 0;
            EC_TRY(mask_ones_end(m->parts[g], pend[g], launched[g] != 0, &c));
            ones += c;
        }
    } else {
        EC_TRY(mask_ones(m, &ones));
    }
    *data = ones;
    *nodata = m->len - ones;
    return EC_OK;
}
ec_status ec_mask_all(const ec_mask* m, int value, int* out) {
    size_t d, nd;
    EC_TRY(ec_mask_counts(m, &d, &nd));
    *out = value ? (nd == 0) : (d == 0);
    return EC_OK;
}
ec_status ec_mask_cmp(const ec_mask* l, const ec_mask* r, int* ordering) {
    EC_TRY(ensure());
    const size_t n = std::min(l->len, r->len);
    if (n) {
        if (is_sharded(l) || is_sharded(r) || l->dev != r->dev) {
            uint64_t bit = ~0ull;
            EC_TRY(shm_first_diff(l, r, n, &bit));
            if (bit != ~0ull) {
                int a = 0;
                EC_TRY(ec_mask_get(l, bit, &a));
                *ordering = a ? 1 : -1;  // false < true
                return EC_OK;
            }
        } else {
            DevScope on(l->dev);
            uint64_t bit = ~0ull;
            EC_TRY(mask_first_diff(l, r, n, &bit));
            if (bit != ~0ull) {
                int a = 0;
                EC_TRY(ec_mask_get(l, bit, &a));
                *ordering = a ? 1 : -1;
                return EC_OK;
            }
        }
    }
    *ordering = l->len < r->len ? -1 : (l->len > r->len ? 1 : 0);
    return EC_OK;
}

// ---- MaskedCellBuffer / NoData ------------------------------------------------------------------------------
ec_status ec_nodata_value(int kind, uint8_t ct, const ec_value* v, ec_value* out, int* has_value) {
    if (!ct_ok(ct) || kind < 0 || kind > 2 || (kind == EC_NODATA_VALUE && !v)) return invalid("nodata");
    *has_value = nodata_sentinel(kind, ct, v, out);
    return EC_OK;
}
ec_status ec_mask_from_nodata(const ec_buf* b, int kind, const ec_value* v, ec_mask** out) {
    EC_TRY(ensure());
    if (kind < 0 || kind > 2 || (kind == EC_NODATA_VALUE && !v)) return invalid("nodata");
    ec_value nd;
    const bool has = nodata_sentinel(kind, b->ct, v, &nd);
    if (has && nd.ct != b->ct) return invalid("NoData<T>: T must be the buffer's cell type");
    if (is_sharded(b) || b->dev != t_dev) return sh_buf_to_mask(b, [=](const ec_buf* part, ec_mask** o) { return ec_mask_from_nodata(part, kind, v, o); }, out);
    if (!has) return ec_mask_fill(b->len, 1, out);  // NoData::None: all valid
    EC_TRY(resolve(b));
    ec_mask* m;
    EC_TRY(new_mask(b->len, &m));
    MaskOwner hold(m, ec_mask_free);
    if (b->len) EC_LAUNCH(launch_mask_build(launch_ctx(), (int)kSize[b->ct], rd(b), b->len, nd.bits, false, m->words, arm_count(m)), "mask_from_nodata");
    *out = hold.release();
    return EC_OK;
}
ec_status ec_buf_fill_nodata(const ec_buf* b, const ec_mask* m, uint8_t dst_ct, int kind, const ec_value* v, ec_buf** out) {
    EC_TRY(ensure());
    if (!ct_ok(dst_ct) || kind < 0 || kind > 2 || (kind == EC_NODATA_VALUE && !v)) return invalid("nodata / cell type");
    if (!ct_fits(b->ct, dst_ct)) return narrowing(b->ct, dst_ct);  // to_vec::<T>()? first (src/masked/masked_buffer.rs:142)
    ec_value nd;
    if (!nodata_sentinel(kind, dst_ct, v, &nd)) return ec_buf_convert(b, dst_ct, out);
    if (nd.ct != dst_ct) return invalid("NoData<T>: T must be the target cell type");
    if (m->len != b->len) { set_error("Mask and buffer must have the same length."); return EC_LEN_MISMATCH; }
    if (is_sharded(b) || is_sharded(m) || b->dev != t_dev || m->dev != t_dev)
        return sh_map_bm(b, m, dst_ct, [=](const ec_buf* bp, const ec_mask* mp, ec_buf** o) { return ec_buf_fill_nodata(bp, mp, dst_ct, kind, v, o); }, out);
    EC_TRY(resolve(b));
    ec_buf* o;
    EC_TRY(new_buf(dst_ct, b->len, &o));
    if (b->len) {
        if (cudaError_t e = launch_fill_nodata(launch_ctx(), b->ct, rd(b), rdm(m), dst_ct, o->dptr, b->len, nd.bits)) { ec_buf_free(o); return cuda_fail(e, "fill_nodata"); }
        note_launch("fill_nodata");
    }
    *out = o;
    return EC_OK;
}
ec_status ec_masked_binary(int op, const ec_buf* lbuf, const ec_mask* lmask, const ec_buf* rbuf, const ec_mask* rmask,
                           ec_buf** out_buf, ec_mask** out_mask) {
    EC_TRY(ensure());
    if (op < 0 || op > 3) return invalid("op");
    if (lmask->len != lbuf->len || rmask->len != rbuf->len) { set_error("Mask and buffer must have the same length."); return EC_LEN_MISMATCH; }
    const size_t n = std::min(lbuf->len, rbuf->len);
    if (n > 0 && (is_sharded(lbuf) || is_sharded(rbuf) || is_sharded(lmask) || is_sharded(rmask) || shard_policy(n) || lbuf->dev != t_dev ||
                  rbuf->dev != t_dev || lmask->dev != t_dev || rmask->dev != t_dev))
        return sh_masked_binary(op, lbuf, lmask, rbuf, rmask, n, out_buf, out_mask);
    ec_mask* om;
    EC_TRY(new_mask(n, &om));
    MaskOwner hold(om, ec_mask_free);
    if (n == 0) {
        EC_TRY(empty_result(out_buf));
        *out_mask = hold.release();
        return EC_OK;
    }
    if (t_lazy && lazy_capable(lbuf) && lazy_capable(rbuf)) {  // data deferred (fusable), mask AND now
        EC_LAUNCH(launch_mask_bitop(launch_ctx(), 1, rdm(lmask), rdm(rmask), n, om->words, arm_count(om)), "mask_bitop");
        EC_TRY(ec_buf_binary(op, lbuf, rbuf, out_buf));
        *out_mask = hold.release();
        return EC_OK;
    }
    EC_TRY(resolve(lbuf));
    EC_TRY(resolve(rbuf));
    BufOwner o;
    EC_TRY(new_buf(EC_FLOAT64, n, &o.b));
    // The shorter operand's mask has no bits past n, so `&` leaves the last word's tail zero.
    EC_LAUNCH(launch_binary(launch_ctx(), op, lbuf->ct, rd(lbuf), rbuf->ct, rd(rbuf), static_cast<double*>(o.b->dptr), n,
                            rdm(lmask), rdm(rmask), om->words, arm_count(om)), "masked_binary");
    *out_buf = o.release();
    *out_mask = hold.release();
    return EC_OK;
}

// ---- sharding ---------------------------------------------------------------------------------------------------
ec_status ec_row_strip(size_t width, size_t height, int n_shards, int shard, size_t* cell_offset, size_t* cell_len) {
    if (n_shards <= 0 || shard < 0 || shard >= n_shards) return invalid("shard index");
    const size_t total = width * height, g = (size_t)shard, G = (size_t)n_shards;
    if (G == 1) { *cell_offset = 0; *cell_len = total; return EC_OK; }
    if (width % 128 == 0) {  // whole rows, remainder rows to the last strip
        const size_t rows = height / G, r0 = g * rows, r1 = (g + 1 == G) ? height : r0 + rows;
        *cell_offset = r0 * width;
        *cell_len = (r1 - r0) * width;
    } else {  // 128-cell aligned cell ranges
        const size_t per = (total / G) & ~size_t(127), c0 = g * per, c1 = (g + 1 == G) ? total : c0 + per;
        *cell_offset = c0;
        *cell_len = c1 - c0;
    }
    return EC_OK;
}
int ec_buf_shard_count(const ec_buf* b) { return static_cast<int>(b->parts.size()); }
int ec_mask_shard_count(const ec_mask* m) { return static_cast<int>(m->parts.size()); }
ec_status ec_buf_shard(const ec_buf* b, int shard, ec_shard_info* info, const ec_buf** strip) {
    const int G = is_sharded(b) ? static_cast<int>(b->parts.size()) : 1;
    if (shard < 0 || shard >= G) return invalid("shard index");
    const ec_buf* part = is_sharded(b) ? b->parts[shard] : b;
    if (!is_sharded(part)) { DevScope on(part->dev); EC_TRY(resolve(part)); }
    if (info) {
        info->logical_device = part->dev;
        info->cuda_device = g_ctx.inited ? g_ctx.dev[part->dev].phys : -1;
        info->offset = is_sharded(b) ? b->offs[shard] : 0;
        info->len = part->len;
        info->device_ptr = part->dptr;
    }
    if (strip) *strip = part;
    return EC_OK;
}
ec_status ec_mask_shard(const ec_mask* m, int shard, ec_shard_info* info, const ec_mask** strip) {
    const int G = is_sharded(m) ? static_cast<int>(m->parts.size()) : 1;
    if (shard < 0 || shard >= G) return invalid("shard index");
    const ec_mask* part = is_sharded(m) ? m->parts[shard] : m;
    if (info) {
        info->logical_device = part->dev;
        info->cuda_device = g_ctx.inited ? g_ctx.dev[part->dev].phys : -1;
        info->offset = is_sharded(m) ? m->offs[shard] : 0;
        info->len = part->len;
        info->device_ptr = part->words;
    }
    if (strip) *strip = part;
    return EC_OK;
}
ec_status ec_buf_min_max_keys(const ec_buf* b, const ec_mask* m, int64_t* device_keys2) {
    EC_TRY(ensure());
    if (is_sharded(b) || (m && is_sharded(m))) return invalid("ec_buf_min_max_keys takes one strip (a sharded buffer finishes its reductions itself: ec_buf_min_max)");
    if (m && m->len != b->len) { set_error("Mask and buffer must have the same length."); return EC_LEN_MISMATCH; }
    if (b->len == 0) {  // the seeds alone (src/buffer.rs:170)
        uint64_t seed[2];
        key_seeds(b->ct, &seed[0], &seed[1]);
        uint64_t* pin;
        EC_TRY(pinned_words(&pin));
        pin[2] = (uint64_t)key_to_signed(seed[0]);
        pin[3] = (uint64_t)~key_to_signed(seed[1]);
        EC_CUDA_TRY(cudaMemcpyAsync(device_keys2, &pin[2], 16, cudaMemcpyHostToDevice, cur_stream()), "cudaMemcpyAsync(H2D)");
        return sync_stream();
    }
    ReduceScratch sc;
    EC_TRY(min_max_begin(b, m, nullptr, nullptr, &sc));
    // the finishing CTA also wrote {skey(min), ~skey(max)}: one MIN all-reduce finishes both. No host round trip.
    EC_CUDA_TRY(cudaMemcpyAsync(device_keys2, sc.result + 2, 16, cudaMemcpyDeviceToDevice, cur_stream()), "cudaMemcpyAsync(D2D)");
    return EC_OK;
}
ec_status ec_min_max_from_keys(uint8_t ct, const int64_t* k, ec_value* mn, ec_value* mx) {
    if (!ct_ok(ct)) return invalid("cell type");
    *mn = tagged<uint64_t>(ct, key_to_bits(ct, key_from_signed(k[0])));
    *mx = tagged<uint64_t>(ct, key_to_bits(ct, key_from_signed(~k[1])));
    return EC_OK;
}

ec_status ec_min_max_to_keys(const ec_value* mn, const ec_value* mx, int64_t* k) {
    if (!ct_ok(mn->ct) || mn->ct != mx->ct) return invalid("cell type");
    k[0] = key_to_signed(key_from_bits(mn->ct, mn->bits));
    k[1] = ~key_to_signed(key_from_bits(mx->ct, mx->bits));
    return EC_OK;
}

ec_status ec_buf_synth(uint8_t ct, size_t len, uint64_t seed, uint64_t index_offset, int kind, double lo, double hi,
                       uint64_t period, const ec_value* sentinel, ec_buf** out) {
    EC_TRY(ensure());
    if (!ct_ok(ct) || kind < 0 || kind > 2) return invalid("cell type / kind");
    if (shard_policy(len))
        return sh_generate(ct, len, [&](size_t off, size_t n, ec_buf** o) { return ec_buf_synth(ct, n, seed, index_offset + off, kind, lo, hi, period, sentinel, o); }, out);
    ec_buf* b;
    EC_TRY(new_buf(ct, len, &b));
    if (len) {
        if (cudaError_t e = launch_synth(launch_ctx(), ct, b->dptr, len, seed, index_offset, kind, lo, hi,
                                         sentinel ? period : 0, sentinel ? sentinel->bits : 0)) { ec_buf_free(b); return cuda_fail(e, "synth"); }
        note_launch("synth");
    }
    *out = b;
    return EC_OK;
}

}  // extern "C"
#include "ec_shard.inc"
#include "ec_ingest.inc"
