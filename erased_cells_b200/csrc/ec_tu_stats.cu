// Moments pass of the statistics extension (10 cell types x masked/unmasked); see ec_stats.cuh.
#include "ec_internal.hpp"
#include "ec_stats.cuh"

#ifndef EC_VB
#define EC_VB 32
#endif

namespace ec {

// FP64-issue-bound for cells narrower than 8 bytes: two 32-byte loads in flight per thread are enough, and the
// register budget goes to the eight accumulators.
constexpr int kStatThreads = 256;
constexpr int kStatUnroll = 2;
constexpr int kStatCtasPerSm = 8;

template <class T>
static cudaError_t moments_t(const Launch& Lc, const void* a, const uint32_t* mask, size_t n, double pivot, double scale,
                             unsigned long long* acc) {
    constexpr int V = EC_VB / sizeof(T);
    constexpr size_t TILE = size_t(kStatThreads) * V * kStatUnroll;
    size_t grid = n / TILE;
    const size_t cap = size_t(Lc.sm_count) * kStatCtasPerSm;
    if (grid > cap) grid = cap;
    if (grid == 0) grid = 1;
    if (mask)
        moments_kernel<T, true, EC_VB, kStatUnroll, kStatThreads><<<int(grid), kStatThreads, 0, Lc.stream>>>(
            static_cast<const T*>(a), mask, n, pivot, scale, acc);
    else
        moments_kernel<T, false, EC_VB, kStatUnroll, kStatThreads><<<int(grid), kStatThreads, 0, Lc.stream>>>(
            static_cast<const T*>(a), nullptr, n, pivot, scale, acc);
    return cudaGetLastError();
}

cudaError_t launch_moments(const Launch& Lc, int ct, const void* a, const uint32_t* mask, size_t n, double pivot, double scale,
                           unsigned long long* acc) {
    switch (ct) {
#define X(id, p) case id: return moments_t<p>(Lc, a, mask, n, pivot, scale, acc);
        EC_WITH_CT(X)
#undef X
    }
    return cudaErrorInvalidValue;
}

// 8/16-bit cells: integer moments, no FP64 (see ec_stats.cuh). A CTA must see fewer than 2^28 cells for its 64-bit
// partials: the grid is never capped below n / 2^28 tiles' worth.
constexpr int kIntStatUnroll = 4;
template <class T>
static cudaError_t int_moments_t(const Launch& Lc, const void* a, const uint32_t* mask, size_t n, unsigned long long* acc) {
    constexpr int V = EC_VB / sizeof(T);
    constexpr size_t TILE = size_t(kStatThreads) * V * kIntStatUnroll;
    size_t grid = n / TILE;
    size_t cap = size_t(Lc.sm_count) * kStatCtasPerSm;
    if (cap < (n >> 28) + 1) cap = (n >> 28) + 1;
    if (grid > cap) grid = cap;
    if (grid == 0) grid = 1;
    if (mask)
        int_moments_kernel<T, true, EC_VB, kIntStatUnroll, kStatThreads><<<int(grid), kStatThreads, 0, Lc.stream>>>(
            static_cast<const T*>(a), mask, n, acc);
    else
        int_moments_kernel<T, false, EC_VB, kIntStatUnroll, kStatThreads><<<int(grid), kStatThreads, 0, Lc.stream>>>(
            static_cast<const T*>(a), nullptr, n, acc);
    return cudaGetLastError();
}
cudaError_t launch_int_moments(const Launch& Lc, int ct, const void* a, const uint32_t* mask, size_t n, unsigned long long* acc) {
    switch (ct) {
        case EC_UINT8: return int_moments_t<uint8_t>(Lc, a, mask, n, acc);
        case EC_UINT16: return int_moments_t<uint16_t>(Lc, a, mask, n, acc);
        case EC_INT8: return int_moments_t<int8_t>(Lc, a, mask, n, acc);
        case EC_INT16: return int_moments_t<int16_t>(Lc, a, mask, n, acc);
    }
    return cudaErrorInvalidValue;
}

}  // namespace ec
