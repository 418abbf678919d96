// Statistics extension: FP64 window moments (64-bit integers, floats) and the one-pass integer kernel (integers of at
// most 32 bits), masked and unmasked; see ec_stats.cuh.
#include "ec_internal.hpp"
#include "ec_stats.cuh"

#include <cmath>

#ifndef EC_VB
#define EC_VB 32
#endif

namespace ec {

// FP64-issue-bound for cells narrower than 8 bytes: two 32-byte loads in flight per thread are enough, and the
// register budget goes to the eight accumulators.
constexpr int kStatThreads = 256;
constexpr int kStatUnroll = 2;

// Persistent grids are sized to what is actually resident (register use differs between the instantiations: a grid of
// 8 CTAs per SM over a kernel that fits 6 runs a second, one-third-full wave).
template <class K> static size_t resident_ctas(K kernel, const Launch& Lc) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kStatThreads, 0) != cudaSuccess || per_sm < 1) {
        (void)cudaGetLastError();
        per_sm = 1;
    }
    return size_t(Lc.sm_count) * per_sm;
}

template <class T>
static cudaError_t moments_t(const Launch& Lc, const void* a, const uint32_t* mask, size_t n, double pivot, double scale,
                             unsigned long long* acc) {
    constexpr int V = EC_VB / sizeof(T);
    constexpr size_t TILE = size_t(kStatThreads) * V * kStatUnroll;
    auto masked = moments_kernel<T, true, EC_VB, kStatUnroll, kStatThreads>;
    auto plain = moments_kernel<T, false, EC_VB, kStatUnroll, kStatThreads>;
    static const size_t cap_masked = resident_ctas(masked, Lc), cap_plain = resident_ctas(plain, Lc);
    size_t grid = n / TILE;
    const size_t cap = mask ? cap_masked : cap_plain;
    if (grid > cap) grid = cap;
    if (grid == 0) grid = 1;
    if (mask) masked<<<int(grid), kStatThreads, 0, Lc.stream>>>(static_cast<const T*>(a), mask, n, pivot, scale, acc);
    else plain<<<int(grid), kStatThreads, 0, Lc.stream>>>(static_cast<const T*>(a), nullptr, n, pivot, scale, acc);
    return cudaGetLastError();
}

cudaError_t launch_moments(const Launch& Lc, int ct, const void* a, const uint32_t* mask, size_t n, double pivot, double scale,
                           unsigned long long* acc) {
    switch (ct) {  // 64-bit integers and floats; narrower integers take the exact integer route (launch_int_stats)
        case EC_UINT64: return moments_t<uint64_t>(Lc, a, mask, n, pivot, scale, acc);
        case EC_INT64: return moments_t<int64_t>(Lc, a, mask, n, pivot, scale, acc);
        case EC_FLOAT64: return moments_t<double>(Lc, a, mask, n, pivot, scale, acc);
    }
    return cudaErrorInvalidValue;
}

// Integer cells of at most 32 bits: min, max, count, sum x, sum x^2 in one pass (see ec_stats.cuh). acc[7], with
// acc[5] preset to all ones (min) and acc[6] to 0 (max); both hold biased (x ^ sign bit) cells. A CTA must see fewer
// than 2^28 cells for its 64-bit partials: the grid is never capped below n / 2^28.
constexpr int kIntStatUnroll = 4;
template <class T>
static cudaError_t int_stats_t(const Launch& Lc, const void* a, const uint32_t* mask, size_t n, unsigned long long* acc) {
    constexpr int V = EC_VB / sizeof(T);
    constexpr size_t TILE = size_t(kStatThreads) * V * kIntStatUnroll;
    auto masked = int_stats_kernel<T, true, EC_VB, kIntStatUnroll, kStatThreads>;
    auto plain = int_stats_kernel<T, false, EC_VB, kIntStatUnroll, kStatThreads>;
    static const size_t cap_masked = resident_ctas(masked, Lc), cap_plain = resident_ctas(plain, Lc);
    size_t grid = n / TILE;
    size_t cap = mask ? cap_masked : cap_plain;
    if (cap < (n >> 28) + 1) cap = (n >> 28) + 1;
    if (grid > cap) grid = cap;
    if (grid == 0) grid = 1;
    if (mask) masked<<<int(grid), kStatThreads, 0, Lc.stream>>>(static_cast<const T*>(a), mask, n, acc, 1.0f, 1.0f);
    else plain<<<int(grid), kStatThreads, 0, Lc.stream>>>(static_cast<const T*>(a), nullptr, n, acc, 1.0f, 1.0f);
    return cudaGetLastError();
}
// Float32 cells quantised to q = rint(x * qs1 * qs2) and summed as int32 (see int_stats_kernel)
static cudaError_t quant_stats(const Launch& Lc, const void* a, const uint32_t* mask, size_t n, unsigned long long* acc, float qs1, float qs2) {
    constexpr int V = EC_VB / 4;
    constexpr size_t TILE = size_t(kStatThreads) * V * kIntStatUnroll;
    auto masked = int_stats_kernel<int32_t, true, EC_VB, kIntStatUnroll, kStatThreads, true>;
    auto plain = int_stats_kernel<int32_t, false, EC_VB, kIntStatUnroll, kStatThreads, true>;
    static const size_t cap_masked = resident_ctas(masked, Lc), cap_plain = resident_ctas(plain, Lc);
    size_t grid = n / TILE;
    size_t cap = mask ? cap_masked : cap_plain;
    if (cap < (n >> 28) + 1) cap = (n >> 28) + 1;
    if (grid > cap) grid = cap;
    if (grid == 0) grid = 1;
    if (mask) masked<<<int(grid), kStatThreads, 0, Lc.stream>>>(static_cast<const int32_t*>(a), mask, n, acc, qs1, qs2);
    else plain<<<int(grid), kStatThreads, 0, Lc.stream>>>(static_cast<const int32_t*>(a), nullptr, n, acc, qs1, qs2);
    return cudaGetLastError();
}
cudaError_t launch_quant_stats(const Launch& Lc, const void* a, const uint32_t* mask, size_t n, unsigned long long* acc, int exp2) {
    // 2^(26 - E) as two f32 factors, each a normal power of two; x * qs1 stays below 2^26 (no overflow), and whatever
    // underflows on the way is far below one half and rounds to 0 either way
    const int k = 26 - exp2, k1 = k > 120 ? 120 : k;
    return quant_stats(Lc, a, mask, n, acc, std::ldexp(1.0f, k1), std::ldexp(1.0f, k - k1));
}
cudaError_t launch_int_stats(const Launch& Lc, int ct, const void* a, const uint32_t* mask, size_t n, unsigned long long* acc) {
    switch (ct) {
        case EC_UINT8: return int_stats_t<uint8_t>(Lc, a, mask, n, acc);
        case EC_UINT16: return int_stats_t<uint16_t>(Lc, a, mask, n, acc);
        case EC_UINT32: return int_stats_t<uint32_t>(Lc, a, mask, n, acc);
        case EC_INT8: return int_stats_t<int8_t>(Lc, a, mask, n, acc);
        case EC_INT16: return int_stats_t<int16_t>(Lc, a, mask, n, acc);
        case EC_INT32: return int_stats_t<int32_t>(Lc, a, mask, n, acc);
    }
    return cudaErrorInvalidValue;
}

}  // namespace ec
