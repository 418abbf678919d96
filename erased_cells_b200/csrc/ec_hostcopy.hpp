// Host-side copy pool of the staged pageable-memory path (ec_ingest.inc): plain C++, no CUDA, so that the CPU tests can
// drive it (tests/cpp/test_hostcopy.cpp).
#pragma once
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace ec {
struct CopyJob {
    char* dst;
    const char* src;
    size_t bytes, piece;
    int pieces, workers;
    std::atomic<int> next{0}, done{0};
};
#if defined(__x86_64__)
// streaming stores: the destination (staging about to be read by the DMA engine, or the caller's fresh Vec) is not read
// back by this core, so skip the read-for-ownership and keep the caches for the source
__attribute__((target("avx2"))) static void copy_stream_avx2(char* d, const char* s, size_t n) {
    while (n && (reinterpret_cast<uintptr_t>(d) & 31)) { *d++ = *s++; --n; }
    size_t v = n / 128;
    for (; v; --v, d += 128, s += 128) {
        const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s));
        const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + 32));
        const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + 64));
        const __m256i e = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + 96));
        _mm256_stream_si256(reinterpret_cast<__m256i*>(d), a);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(d + 32), b);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(d + 64), c);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(d + 96), e);
    }
    _mm_sfence();
    n &= 127;
    if (n) memcpy(d, s, n);
}
#endif
static std::atomic<int> g_copy_nt{-1};  // -1: not decided yet; 0 memcpy; 1 streaming stores (several threads may decide at once: same answer)
static bool copy_stream_env() {
    const char* v = getenv("EC_HOST_COPY_STREAM");
    return !v || atoi(v) != 0;
}
static void copy_piece(char* d, const char* s, size_t n) {
#if defined(__x86_64__)
    int nt = g_copy_nt.load(std::memory_order_relaxed);
    if (nt < 0) {
        nt = copy_stream_env() && __builtin_cpu_supports("avx2");
        g_copy_nt.store(nt, std::memory_order_relaxed);
    }
    if (nt) { copy_stream_avx2(d, s, n); return; }
#endif
    memcpy(d, s, n);
}
struct CopyPool {
    std::mutex job_mu;  // one chunk at a time: concurrent transfers take turns, each at the full width of the pool
    std::mutex mu;
    std::condition_variable cv;
    std::shared_ptr<CopyJob> cur;
    std::atomic<uint64_t> gen{0};
    std::vector<std::thread> th;
    int sleepers = 0;
    static void work(CopyJob& j) {
        for (;;) {
            const int i = j.next.fetch_add(1, std::memory_order_acq_rel);
            if (i >= j.pieces) return;
            const size_t off = size_t(i) * j.piece;
            copy_piece(j.dst + off, j.src + off, std::min(j.piece, j.bytes - off));
            j.done.fetch_add(1, std::memory_order_release);
        }
    }
    void worker(int index) {
        uint64_t seen = 0;
        for (;;) {
            int spins = 0;
            while (gen.load(std::memory_order_acquire) == seen) {
                if (++spins < 20000) {
#if defined(__x86_64__)
                    __builtin_ia32_pause();
#endif
                } else {
                    std::unique_lock<std::mutex> lk(mu);
                    ++sleepers;
                    cv.wait(lk, [&] { return gen.load(std::memory_order_acquire) != seen; });
                    --sleepers;
                }
            }
            std::shared_ptr<CopyJob> j;
            {
                std::lock_guard<std::mutex> lk(mu);
                j = cur;
                seen = gen.load(std::memory_order_acquire);
            }
            if (j && index < j->workers) work(*j);
        }
    }
    void ensure_threads(int want) {  // the caller is one of the `threads`
        while (static_cast<int>(th.size()) < want) {
            const int index = static_cast<int>(th.size());
            th.emplace_back([this, index] { worker(index); });
            th.back().detach();  // parked for the life of the process
        }
    }
    void copy(void* dst, const void* src, size_t bytes, int threads) {
        if (threads <= 1 || bytes < (size_t(1) << 20)) { copy_piece(static_cast<char*>(dst), static_cast<const char*>(src), bytes); return; }
        std::lock_guard<std::mutex> turn(job_mu);
        ensure_threads(threads - 1);
        auto j = std::make_shared<CopyJob>();
        j->dst = static_cast<char*>(dst); j->src = static_cast<const char*>(src); j->bytes = bytes;
        j->pieces = 2 * threads;  // two pieces per thread: a late starter leaves its second piece to the others
        j->piece = ((bytes + j->pieces - 1) / j->pieces + 4095) & ~size_t(4095);
        j->pieces = static_cast<int>((bytes + j->piece - 1) / j->piece);
        j->workers = threads - 1;
        bool wake;
        {
            std::lock_guard<std::mutex> lk(mu);
            cur = j;
            gen.fetch_add(1, std::memory_order_release);
            wake = sleepers > 0;
        }
        if (wake) cv.notify_all();
        work(*j);
        while (j->done.load(std::memory_order_acquire) < j->pieces) {
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
        }
    }
};
// The pool lives on the heap for the life of the process: its threads are detached and may sleep on the condition variable
// when the process exits, and destroying a condition variable that has waiters blocks.
inline CopyPool& copy_pool() {
    static CopyPool* pool = new CopyPool;
    return *pool;
}
}  // namespace ec
