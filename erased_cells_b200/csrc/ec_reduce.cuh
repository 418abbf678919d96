// Reductions: min_max (src/buffer.rs:169-173, masked: src/masked/masked_buffer.rs:208-217),
// Mask::counts / Mask::all (src/masked/mask.rs:67-80) and the first-difference search behind
// Ord/Eq for CellBuffer (src/buffer.rs:390-436). Read-only streams: algorithmic bytes per cell =
// size_of(T) (+ 1/8 with a mask).
//
// Shape: 256-bit loads, UNROLL independent requests per thread -> per-thread accumulators on
// order-preserving unsigned keys -> warp redux/shuffle -> shared-memory block tree -> one partial
// per CTA -> the last CTA to finish (atomic ticket) folds the partials and writes the result. One
// launch, no float atomics, exactly associative, so the answer does not depend on the grid.
#pragma once
#include "ec_common.cuh"

namespace ec {

// Cross-GPU finish over NVLink peer memory (row-strip sharding, one rank per GPU). When `peers` is set, the CTA
// that folds this GPU's partials also exchanges the 16-byte partial result with every peer — remote stores into
// each peer's mailbox, then a bounded spin on its own mailbox — and reduces the n_ranks pairs, so a sharded
// reduction is ONE kernel per GPU with no separate collective launch. Mailbox slot (32 bytes) of sender r for
// epoch e lives at word 4 * ((e & 1) * n_ranks + r): four words {tag(e) | 32 payload bits}. Two epochs of slots suffice:
// a rank can only be one collective ahead of the slowest peer, because finishing epoch e needs every peer's
// epoch-e message.
struct PeerExchange {
    unsigned long long* const* peers;  // device array [n_ranks]: every rank's mailbox, peer-mapped (null = single GPU)
    int n_ranks, rank;
    unsigned long long epoch;          // >= 1, identical on all ranks for one collective
    unsigned long long spin_limit;     // clock64() ticks before giving up on a peer
};
struct ReduceScratch {
    uint64_t* partials;     // 2 * max_blocks
    unsigned int* ticket;   // zero between launches (the finishing CTA resets it)
    uint64_t* result;       // 4 words: [0..1] raw result; [2..3] min_max as {skey(min), ~skey(max)} for a MIN all-reduce
    uint64_t* host_result;  // optional device alias of mapped pinned host memory: five words {tag | 32 bits}: r0 lo/hi, r1 lo/hi, status
    uint64_t host_seq;      // the 32-bit tag of this call (top bit set): the host polls for it instead of synchronising the stream
    int early_trigger;      // experiment knob: let the next grid be scheduled at once (griddepcontrol.launch_dependents) or only when this one ends
    uint64_t* trace;        // null, or 8 words of mapped pinned memory the kernel stamps with %globaltimer (ec_set_reduce_trace): where a call's time goes
    PeerExchange px;
};

__device__ __forceinline__ uint64_t global_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// trace slots: 0 first CTA past the dependency wait, 1 last CTA of the grid has its partial, 2 this GPU's result folded,
// 3 partial sent to every peer, 4 every peer's partial received and folded, 5 result published to the host
__device__ __forceinline__ void stamp(const ReduceScratch& s, int slot) {
    if (s.trace != nullptr) { volatile uint64_t* t = s.trace; t[slot] = global_ns(); }
}
__device__ __forceinline__ uint32_t warp_min(uint32_t v) { return __reduce_min_sync(0xFFFFFFFFu, v); }
__device__ __forceinline__ uint32_t warp_max(uint32_t v) { return __reduce_max_sync(0xFFFFFFFFu, v); }
__device__ __forceinline__ uint64_t warp_min(uint64_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { uint64_t x = __shfl_xor_sync(0xFFFFFFFFu, v, o); v = x < v ? x : v; }
    return v;
}
__device__ __forceinline__ uint64_t warp_max(uint64_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { uint64_t x = __shfl_xor_sync(0xFFFFFFFFu, v, o); v = x > v ? x : v; }
    return v;
}
__device__ __forceinline__ uint64_t warp_sum(uint64_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}

enum : int { RED_MINMAX = 0, RED_SUM = 1, RED_MIN = 2 };

template <int MODE> __device__ __forceinline__ void combine(uint64_t& a0, uint64_t& a1, uint64_t b0, uint64_t b1) {
    if constexpr (MODE == RED_MINMAX) { a0 = b0 < a0 ? b0 : a0; a1 = b1 > a1 ? b1 : a1; }
    else if constexpr (MODE == RED_SUM) { a0 += b0; a1 += b1; }
    else { a0 = b0 < a0 ? b0 : a0; a1 = b1 < a1 ? b1 : a1; }
}

// Block tree + cross-CTA finish. (k0,k1) are this thread's warp-reduced values (valid in lane 0).
template <int MODE, int THREADS>
__device__ __forceinline__ void block_finish(uint64_t k0, uint64_t k1, uint64_t id0, uint64_t id1, ReduceScratch s) {
    __shared__ uint64_t sh0[THREADS / 32], sh1[THREADS / 32];
    __shared__ bool last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { sh0[warp] = k0; sh1[warp] = k1; }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t a0 = sh0[0], a1 = sh1[0];
#pragma unroll
        for (int w = 1; w < THREADS / 32; ++w) combine<MODE>(a0, a1, sh0[w], sh1[w]);
        s.partials[2 * blockIdx.x] = a0;
        s.partials[2 * blockIdx.x + 1] = a1;
        __threadfence();
        const unsigned int done = atomicAdd(s.ticket, 1u);
        last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    if (threadIdx.x == 0) stamp(s, 1);
    uint64_t a0 = id0, a1 = id1;
    for (unsigned int i = threadIdx.x; i < gridDim.x; i += THREADS) {
        combine<MODE>(a0, a1, __ldcg(s.partials + 2 * i), __ldcg(s.partials + 2 * i + 1));
    }
    if constexpr (MODE == RED_MINMAX) { a0 = warp_min(a0); a1 = warp_max(a1); }
    else if constexpr (MODE == RED_SUM) { a0 = warp_sum(a0); a1 = warp_sum(a1); }
    else { a0 = warp_min(a0); a1 = warp_min(a1); }
    __syncthreads();
    if (lane == 0) { sh0[warp] = a0; sh1[warp] = a1; }
    __syncthreads();
    __shared__ uint64_t xr[2];
    __shared__ unsigned int xstatus;
    if (threadIdx.x == 0) {
        a0 = sh0[0]; a1 = sh1[0];
#pragma unroll
        for (int w = 1; w < THREADS / 32; ++w) combine<MODE>(a0, a1, sh0[w], sh1[w]);
        xr[0] = a0; xr[1] = a1;
        xstatus = 0;
        stamp(s, 2);
    }
    if (s.px.peers != nullptr) {
        // ---- exchange with the peer GPUs (uniform branch: every thread of the CTA takes it) ----
        // A message is four 64-bit words, each carrying 32 bits of payload under a 32-bit tag made from the epoch: a word
        // is valid exactly when its tag is the current one, and an aligned 64-bit store arrives whole — so the sender
        // needs no system-scope fence between "data" and "flag" (there is no separate flag), and the receiver simply
        // polls until all four tags match. One NVLink crossing instead of a fenced round trip plus a crossing.
        __syncthreads();
        const int n = s.px.n_ranks;
        const unsigned long long base = 4ull * ((s.px.epoch & 1ull) * n);
        const unsigned long long tag = (0x80000000ull | (s.px.epoch & 0x7FFFFFFFull)) << 32;
        if (threadIdx.x < n) {  // thread t talks to rank t: send ...
            volatile unsigned long long* slot = s.px.peers[threadIdx.x] + base + 4ull * s.px.rank;
            slot[0] = tag | (xr[0] & 0xFFFFFFFFull);
            slot[1] = tag | (xr[0] >> 32);
            slot[2] = tag | (xr[1] & 0xFFFFFFFFull);
            slot[3] = tag | (xr[1] >> 32);
            if (threadIdx.x == 0) stamp(s, 3);
        }
        uint64_t p0 = id0, p1 = id1;
        if (threadIdx.x < n) {  // ... then receive rank t's partial from our own mailbox
            volatile unsigned long long* slot = s.px.peers[s.px.rank] + base + 4ull * threadIdx.x;
            const long long t0 = clock64();
            bool ok = true;
            unsigned long long w0, w1, w2, w3;
            for (;;) {
                w0 = slot[0]; w1 = slot[1]; w2 = slot[2]; w3 = slot[3];
                if (((w0 ^ tag) >> 32) == 0 && ((w1 ^ tag) >> 32) == 0 && ((w2 ^ tag) >> 32) == 0 && ((w3 ^ tag) >> 32) == 0) break;
                if (static_cast<unsigned long long>(clock64() - t0) > s.px.spin_limit) { ok = false; break; }
            }
            if (ok) { p0 = (w0 & 0xFFFFFFFFull) | (w1 << 32); p1 = (w2 & 0xFFFFFFFFull) | (w3 << 32); } else { atomicOr(&xstatus, 1u); }
        }
        // fold the n_ranks pairs (n_ranks <= 32: one warp)
        if (threadIdx.x < 32) {
            if constexpr (MODE == RED_MINMAX) { p0 = warp_min(p0); p1 = warp_max(p1); }
            else if constexpr (MODE == RED_SUM) { p0 = warp_sum(p0); p1 = warp_sum(p1); }
            else { p0 = warp_min(p0); p1 = warp_min(p1); }
            if (threadIdx.x == 0) { xr[0] = p0; xr[1] = p1; stamp(s, 4); }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        a0 = xr[0]; a1 = xr[1];
        s.result[0] = a0;
        s.result[1] = a1;
        if constexpr (MODE == RED_MINMAX) {  // order-preserving signed keys, max negated
            s.result[2] = a0 ^ 0x8000000000000000ull;
            s.result[3] = ~(a1 ^ 0x8000000000000000ull);
        }
        if (s.host_result != nullptr) {
            // mapped pinned memory, same self-validating layout: five words {tag | 32 payload bits} — a0, a1 in halves and
            // the status. The host polls until all five carry this call's tag: no D2H copy, no stream sync, no fence.
            volatile uint64_t* hr = s.host_result;
            const uint64_t htag = s.host_seq << 32;
            hr[0] = htag | (a0 & 0xFFFFFFFFull);
            hr[1] = htag | (a0 >> 32);
            hr[2] = htag | (a1 & 0xFFFFFFFFull);
            hr[3] = htag | (a1 >> 32);
            hr[4] = htag | xstatus;
            stamp(s, 5);
        }
        *s.ticket = 0;  // ready for the next launch on this stream
    }
}

// ---- exact sums across the GPUs (statistics of a row-strip sharded raster) -----------------------------------------
// After a strip's statistics kernel has left its raw accumulators in device memory — word 0 a plain count, then
// `pairs` 128-bit two's complement sums as (lo, hi) words — ONE CTA sends them to every peer's mailbox, receives every
// peer's, adds them exactly (carries included) and publishes the totals to mapped pinned host memory. Same
// self-validating words as the min_max exchange: 32 payload bits under the epoch's tag, no fences. It is enqueued right
// behind the statistics kernel, so its launch latency hides under that kernel.
constexpr int kSumWords = 10;  // values per message (count + 4 pairs = 9 used)
template <int THREADS>
__global__ void __launch_bounds__(THREADS) exchange_sums_kernel(const unsigned long long* __restrict__ local, int K, int pairs, PeerExchange px,
                                                           unsigned long long region_off, uint64_t* host_words, uint64_t host_seq) {
    __shared__ unsigned int half[32][2 * kSumWords];
    __shared__ unsigned int status;
    const int n = px.n_ranks, W = 2 * K;
    const unsigned long long tag = (0x80000000ull | (px.epoch & 0x7FFFFFFFull)) << 32;
    const unsigned long long slot_words = 2ull * kSumWords;
    if (threadIdx.x == 0) status = 0u;
    __syncthreads();
    for (int i = threadIdx.x; i < n * W; i += blockDim.x) {  // send: word w of my message into slot `rank` of peer r
        const int r = i / W, w = i % W;
        const unsigned long long v = local[w >> 1];
        volatile unsigned long long* slot = px.peers[r] + region_off + ((px.epoch & 1ull) * n + px.rank) * slot_words + w;
        *slot = tag | ((w & 1) ? (v >> 32) : (v & 0xFFFFFFFFull));
    }
    for (int i = threadIdx.x; i < n * W; i += blockDim.x) {  // receive: word w of rank r's message from my own mailbox
        const int r = i / W, w = i % W;
        volatile unsigned long long* slot = px.peers[px.rank] + region_off + ((px.epoch & 1ull) * n + r) * slot_words + w;
        const long long t0 = clock64();
        unsigned long long x;
        for (;;) {
            x = *slot;
            if (((x ^ tag) >> 32) == 0) break;
            if (static_cast<unsigned long long>(clock64() - t0) > px.spin_limit) { atomicOr(&status, 1u); x = 0; break; }
        }
        half[r][w] = static_cast<unsigned int>(x);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long tot[kSumWords];
        for (int k = 0; k < K; ++k) tot[k] = 0;
        for (int r = 0; r < n; ++r) {
            tot[0] += (static_cast<unsigned long long>(half[r][1]) << 32) | half[r][0];
            for (int p = 0; p < pairs; ++p) {  // 128-bit add with carry
                const unsigned long long lo = (static_cast<unsigned long long>(half[r][2 * (1 + 2 * p) + 1]) << 32) | half[r][2 * (1 + 2 * p)];
                const unsigned long long hi = (static_cast<unsigned long long>(half[r][2 * (2 + 2 * p) + 1]) << 32) | half[r][2 * (2 + 2 * p)];
                const unsigned long long s = tot[1 + 2 * p] + lo;
                tot[2 + 2 * p] += hi + (s < lo ? 1ull : 0ull);
                tot[1 + 2 * p] = s;
            }
        }
        volatile uint64_t* hr = host_words;
        const uint64_t htag = host_seq << 32;
        for (int k = 0; k < K; ++k) {
            hr[2 * k] = htag | (tot[k] & 0xFFFFFFFFull);
            hr[2 * k + 1] = htag | (tot[k] >> 32);
        }
        hr[W] = htag | status;
    }
}

// ---- min_max ---------------------------------------------------------------------------------
// Unmasked 8/16-bit cells are reduced two-at-a-time in packed 16-bit lanes (VIMNMX.U16x2); wider
// cells and the masked flavour go cell by cell on 32/64-bit keys.
template <class T, bool MASKED, int VB, int UNROLL, int THREADS>
__global__ void __launch_bounds__(THREADS) min_max_kernel(const T* __restrict__ a, const uint32_t* __restrict__ m,
                                                          size_t n, okey_t<T> seed_min, okey_t<T> seed_max,
                                                          ReduceScratch s) {
    using K = okey_t<T>;
    constexpr int V = VB / sizeof(T);
    constexpr size_t TILE = size_t(THREADS) * V * UNROLL;
    const size_t full = n / TILE;
    K kmin = seed_min, kmax = seed_max;
    if (s.early_trigger) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (blockIdx.x == 0 && threadIdx.x == 0) stamp(s, 0);

    if constexpr (MASKED && sizeof(T) == 1 && VB == 32) {
        // 8-bit cells with a mask: a thread's 32 cells are exactly one mask word. Invalid bytes are forced to the
        // identity of the biased unsigned domain (0xFF for min, 0x00 for max — which are also the seeds T::MAX /
        // T::MIN), then the same packed 16-bit-lane min/max as the unmasked path.
        constexpr bool SG = std::is_signed<T>::value;
        constexpr uint32_t BIAS = SG ? 0x80808080u : 0u;
        uint32_t pmin = 0xFFFFFFFFu, pmax = 0u;
        for (size_t t = blockIdx.x; t < full; t += gridDim.x) {
            const size_t base = t * TILE + size_t(threadIdx.x) * V;
            Vec<uint32_t, 8> w[UNROLL];
            uint32_t mw[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const size_t c = base + size_t(u) * THREADS * V;
                w[u] = ld_stream<uint32_t, 8>(reinterpret_cast<const uint32_t*>(a + c));
                mw[u] = __ldg(m + c / 32);
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint32_t nib = (mw[u] >> (4 * j)) & 0xFu;
                    const uint32_t sel = ((nib * 0x00204081u) & 0x01010101u) * 0xFFu;  // 4 mask bits -> 4 byte masks
                    const uint32_t x = w[u].v[j] ^ BIAS;
                    const uint32_t xmin = x | ~sel, xmax = x & sel;
                    pmin = __vimin3_u16x2(pmin, __byte_perm(xmin, 0u, 0x4240), __byte_perm(xmin, 0u, 0x4341));
                    pmax = __vimax3_u16x2(pmax, __byte_perm(xmax, 0u, 0x4240), __byte_perm(xmax, 0u, 0x4341));
                }
            }
        }
        constexpr uint32_t LB = SG ? 0x80u : 0u;
        const uint32_t lo = min(pmin & 0xFFFFu, pmin >> 16), hi = max(pmax & 0xFFFFu, pmax >> 16);
        if (full > 0 && blockIdx.x < full) {
            const T vlo = static_cast<T>(static_cast<bits_t<T>>(lo ^ LB)), vhi = static_cast<T>(static_cast<bits_t<T>>(hi ^ LB));
            const K klo = to_key<T>(vlo), khi = to_key<T>(vhi);
            kmin = klo < kmin ? klo : kmin;
            kmax = khi > kmax ? khi : kmax;
        }
    } else if constexpr (!MASKED && sizeof(T) <= 2) {
        constexpr bool SG = std::is_signed<T>::value;
        constexpr uint32_t BIAS = sizeof(T) == 1 ? (SG ? 0x80808080u : 0u) : (SG ? 0x80008000u : 0u);
        uint32_t pmin = 0xFFFFFFFFu, pmax = 0u;  // two 16-bit lanes of biased keys
        for (size_t t = blockIdx.x; t < full; t += gridDim.x) {
            const size_t base = t * TILE + size_t(threadIdx.x) * V;
            Vec<uint32_t, VB / 4> w[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
                w[u] = ld_stream<uint32_t, VB / 4>(reinterpret_cast<const uint32_t*>(a + base + size_t(u) * THREADS * V));
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
#pragma unroll
                for (int j = 0; j < VB / 4; ++j) {
                    const uint32_t x = w[u].v[j] ^ BIAS;
                    if constexpr (sizeof(T) == 2) {
                        pmin = __vminu2(pmin, x);
                        pmax = __vmaxu2(pmax, x);
                    } else {
                        const uint32_t e = __byte_perm(x, 0u, 0x4240), o = __byte_perm(x, 0u, 0x4341);
                        pmin = __vimin3_u16x2(pmin, e, o);
                        pmax = __vimax3_u16x2(pmax, e, o);
                    }
                }
            }
        }
        // unbias back into the generic 32-bit key domain
        constexpr uint32_t LB = sizeof(T) == 1 ? (SG ? 0x80u : 0u) : (SG ? 0x8000u : 0u);
        const uint32_t lo = min(pmin & 0xFFFFu, pmin >> 16), hi = max(pmax & 0xFFFFu, pmax >> 16);
        if (full > 0 && blockIdx.x < full) {
            // biased narrow key -> value -> 32-bit key
            const T vlo = static_cast<T>(static_cast<bits_t<T>>(lo ^ LB)), vhi = static_cast<T>(static_cast<bits_t<T>>(hi ^ LB));
            const K klo = to_key<T>(vlo), khi = to_key<T>(vhi);
            kmin = klo < kmin ? klo : kmin;
            kmax = khi > kmax ? khi : kmax;
        }
    } else {
        constexpr uint32_t VMASK = V >= 32 ? 0xFFFFFFFFu : ((1u << (V & 31)) - 1u);
        for (size_t t = blockIdx.x; t < full; t += gridDim.x) {
            const size_t base = t * TILE + size_t(threadIdx.x) * V;
            Vec<T, V> va[UNROLL];
            uint32_t w[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const size_t c = base + size_t(u) * THREADS * V;
                va[u] = ld_stream<T, V>(a + c);
                if constexpr (MASKED) w[u] = (__ldg(m + c / 32) >> (c % 32)) & VMASK;
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    const K k = to_key<T>(va[u].v[j]);
                    if constexpr (MASKED) {
                        const bool valid = (w[u] >> j) & 1u;
                        kmin = (valid && k < kmin) ? k : kmin;
                        kmax = (valid && k > kmax) ? k : kmax;
                    } else {
                        kmin = k < kmin ? k : kmin;
                        kmax = k > kmax ? k : kmax;
                    }
                }
            }
        }
    }
    if (blockIdx.x == full % gridDim.x) {
        for (size_t i = full * TILE + threadIdx.x; i < n; i += THREADS) {
            bool valid = true;
            if constexpr (MASKED) valid = (m[i / 32] >> (i % 32)) & 1u;
            const K k = to_key<T>(a[i]);
            kmin = (valid && k < kmin) ? k : kmin;
            kmax = (valid && k > kmax) ? k : kmax;
        }
    }
    kmin = warp_min(kmin);
    kmax = warp_max(kmax);
    block_finish<RED_MINMAX, THREADS>(kmin, kmax, seed_min, seed_max, s);
}

// ---- Mask::counts: popcount over packed words (tail bits beyond len are kept zero) ---------------
template <int THREADS>
__global__ void __launch_bounds__(THREADS) popcount_kernel(const uint32_t* __restrict__ m, size_t words, ReduceScratch s,
                                                           uint64_t second_word) {
    uint64_t c = 0;
    const size_t groups = words / 4;  // 16-byte groups (allocation is padded to 16 bytes, pad is zero)
    if (s.early_trigger) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    for (size_t g = blockIdx.x * size_t(THREADS) + threadIdx.x; g < groups; g += size_t(gridDim.x) * THREADS) {
        const Vec<uint32_t, 4> w = ld_stream<uint32_t, 4>(m + 4 * g);
        c += __popc(w.v[0]) + __popc(w.v[1]) + __popc(w.v[2]) + __popc(w.v[3]);
    }
    if (blockIdx.x == 0 && threadIdx.x < words % 4) c += __popc(m[groups * 4 + threadIdx.x]);
    c = warp_sum(c);
    // the second word of the pair is a caller-supplied constant (the strip length for sharded counts), added once
    block_finish<RED_SUM, THREADS>(c, (blockIdx.x == 0 && threadIdx.x == 0) ? second_word : 0, 0, 0, s);
}

// ---- first differing cell of two buffers of the same type (bitwise == total-order equality) -------
template <class T, int VB, int THREADS>
__global__ void __launch_bounds__(THREADS) first_diff_kernel(const T* __restrict__ a, const T* __restrict__ b, size_t n,
                                                             ReduceScratch s) {
    constexpr int V = VB / sizeof(T);
    uint64_t first = ~0ull;
    const size_t groups = n / V;
    if (s.early_trigger) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // compare 32 bytes as four 64-bit words; only a differing word is examined cell by cell (lowest differing byte)
    for (size_t g = blockIdx.x * size_t(THREADS) + threadIdx.x; g < groups; g += size_t(gridDim.x) * THREADS) {
        const Raw<VB> x = ld_stream_raw<VB>(a + g * V), y = ld_stream_raw<VB>(b + g * V);
        static_assert(VB == 32, "first_diff_kernel compares 4 x 64-bit words per thread");
#pragma unroll
        for (int w = 3; w >= 0; --w) {
            const uint64_t d = x.r[w] ^ y.r[w];
            if (d != 0) {
                const uint64_t i = g * V + (w * 8 + (__ffsll(static_cast<long long>(d)) - 1) / 8) / sizeof(T);
                first = i < first ? i : first;
            }
        }
    }
    if (blockIdx.x == 0)
        for (size_t i = groups * V + threadIdx.x; i < n; i += THREADS)
            if (to_bits(a[i]) != to_bits(b[i])) first = i < first ? i : first;
    first = warp_min(first);
    block_finish<RED_MIN, THREADS>(first, ~0ull, ~0ull, ~0ull, s);
}

}  // namespace ec
