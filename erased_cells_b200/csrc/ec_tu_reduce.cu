// Reductions and mask kernels: min_max (10 types x masked/unmasked), popcount, first difference,
// mask build / unpack / bit ops / fill, synthetic rasters.
#include "ec_internal.hpp"
#include "ec_mask.cuh"
#include "ec_reduce.cuh"

#ifndef EC_VB
#define EC_VB 32
#endif
#ifndef EC_RUNROLL
#define EC_RUNROLL 4
#endif

namespace ec {

// Geometry picked from the tools/ubench sweep on B200 (profiles/r01_ubench_sweep.md): unmasked
// min_max runs best with 512-thread CTAs, 4 x 32-byte loads in flight per thread and a grid of up to
// 32 CTAs per SM (6.86-6.97 TB/s for f32/u8/i16/f64); the masked flavour with 2 loads in flight and
// 8 CTAs per SM (6.77 TB/s).
constexpr int kRedThreads = 512;
constexpr int kRedUnroll = 4;
constexpr int kRedUnrollMasked = 2;
constexpr size_t kMaxReduceBlocks = 8192;  // size of ReduceScratch::partials (pairs), see ec_api.cu

static int reduce_grid(size_t n, size_t tile, const Launch& Lc, int ctas_per_sm = 32) {
    size_t full = n / tile;
    if (full == 0) full = 1;
    size_t cap = size_t(Lc.sm_count) * ctas_per_sm;
    if (cap > kMaxReduceBlocks) cap = kMaxReduceBlocks;
    return int(full < cap ? full : cap);
}

template <class T>
static cudaError_t min_max_t(const Launch& Lc, const void* a, const uint32_t* mask, size_t n, const ReduceScratch& s) {
    constexpr int V = EC_VB / sizeof(T);
    const okey_t<T> smin = to_key<T>(std::numeric_limits<T>::max()), smax = to_key<T>(std::numeric_limits<T>::lowest());
    if (mask) {
        constexpr size_t TILE = size_t(kRedThreads) * V * kRedUnrollMasked;
        min_max_kernel<T, true, EC_VB, kRedUnrollMasked, kRedThreads><<<reduce_grid(n, TILE, Lc, 8), kRedThreads, 0, Lc.stream>>>(
            static_cast<const T*>(a), mask, n, smin, smax, s);
    } else {
        constexpr size_t TILE = size_t(kRedThreads) * V * kRedUnroll;
        min_max_kernel<T, false, EC_VB, kRedUnroll, kRedThreads><<<reduce_grid(n, TILE, Lc, 32), kRedThreads, 0, Lc.stream>>>(
            static_cast<const T*>(a), nullptr, n, smin, smax, s);
    }
    return cudaGetLastError();
}
cudaError_t launch_min_max(const Launch& Lc, int ct, const void* a, const uint32_t* mask, size_t n, const ReduceScratch& s) {
    switch (ct) {
#define X(id, p) case id: return min_max_t<p>(Lc, a, mask, n, s);
        EC_WITH_CT(X)
#undef X
    }
    return cudaErrorInvalidValue;
}
void key_seeds(int ct, uint64_t* seed_min, uint64_t* seed_max) {
    switch (ct) {
#define X(id, p) case id: *seed_min = to_key<p>(std::numeric_limits<p>::max()); *seed_max = to_key<p>(std::numeric_limits<p>::lowest()); break;
        EC_WITH_CT(X)
#undef X
    }
}
uint64_t key_to_bits(int ct, uint64_t key) {
    switch (ct) {
#define X(id, p) case id: return static_cast<uint64_t>(to_bits<p>(from_key<p>(static_cast<okey_t<p>>(key))));
        EC_WITH_CT(X)
#undef X
    }
    return 0;
}
uint64_t key_from_bits(int ct, uint64_t bits) {
    switch (ct) {
#define X(id, p) case id: return static_cast<uint64_t>(to_key<p>(from_bits<p>(static_cast<bits_t<p>>(bits))));
        EC_WITH_CT(X)
#undef X
    }
    return 0;
}
int64_t key_to_signed(uint64_t key) { return static_cast<int64_t>(key ^ 0x8000000000000000ull); }
uint64_t key_from_signed(int64_t skey) { return static_cast<uint64_t>(skey) ^ 0x8000000000000000ull; }

cudaError_t launch_popcount(const Launch& Lc, const uint32_t* words, size_t nwords, const ReduceScratch& s) {
    const int grid = reduce_grid(nwords / 4, kThreads, Lc);
    popcount_kernel<kThreads><<<grid, kThreads, 0, Lc.stream>>>(words, nwords, s);
    return cudaGetLastError();
}

template <class U> static cudaError_t first_diff_u(const Launch& Lc, const void* a, const void* b, size_t n, const ReduceScratch& s) {
    const int grid = reduce_grid(n / (EC_VB / sizeof(U)), kThreads, Lc);
    first_diff_kernel<U, EC_VB, kThreads><<<grid, kThreads, 0, Lc.stream>>>(static_cast<const U*>(a), static_cast<const U*>(b), n, s);
    return cudaGetLastError();
}
cudaError_t launch_first_diff(const Launch& Lc, int cell_bytes, const void* a, const void* b, size_t n, const ReduceScratch& s) {
    switch (cell_bytes) {
        case 1: return first_diff_u<uint8_t>(Lc, a, b, n, s);
        case 2: return first_diff_u<uint16_t>(Lc, a, b, n, s);
        case 4: return first_diff_u<uint32_t>(Lc, a, b, n, s);
        default: return first_diff_u<uint64_t>(Lc, a, b, n, s);
    }
}

template <class U, bool PACK>
static cudaError_t mask_build_u(const Launch& Lc, const void* a, size_t n, uint64_t sentinel, uint32_t* out) {
    constexpr int V0 = EC_VB / sizeof(U);
    constexpr int V = V0 > 32 ? 32 : V0;
    constexpr size_t TILE = size_t(kThreads) * V * EC_RUNROLL;
    mask_build_kernel<U, PACK, EC_VB, EC_RUNROLL, kThreads><<<grid_for(n, TILE, Lc), kThreads, 0, Lc.stream>>>(
        static_cast<const U*>(a), n, static_cast<U>(sentinel), out);
    return cudaGetLastError();
}
cudaError_t launch_mask_build(const Launch& Lc, int cell_bytes, const void* a, size_t n, uint64_t sentinel_bits,
                              bool pack_bools, uint32_t* out) {
    if (pack_bools) return mask_build_u<uint8_t, true>(Lc, a, n, 0, out);
    switch (cell_bytes) {
        case 1: return mask_build_u<uint8_t, false>(Lc, a, n, sentinel_bits, out);
        case 2: return mask_build_u<uint16_t, false>(Lc, a, n, sentinel_bits, out);
        case 4: return mask_build_u<uint32_t, false>(Lc, a, n, sentinel_bits, out);
        default: return mask_build_u<uint64_t, false>(Lc, a, n, sentinel_bits, out);
    }
}
cudaError_t launch_mask_unpack(const Launch& Lc, const uint32_t* m, size_t n, uint8_t* out) {
    mask_unpack_kernel<kThreads><<<grid_for(n, size_t(kThreads) * 16, Lc), kThreads, 0, Lc.stream>>>(m, n, out);
    return cudaGetLastError();
}
cudaError_t launch_mask_bitop(const Launch& Lc, int mop, const uint32_t* l, const uint32_t* r, size_t n, uint32_t* out) {
    mask_bitop_kernel<kThreads><<<grid_for((n + 127) / 128, kThreads, Lc), kThreads, 0, Lc.stream>>>(mop, l, r, n, out);
    return cudaGetLastError();
}
cudaError_t launch_mask_fill(const Launch& Lc, uint32_t* out, size_t n, bool value) {
    mask_fill_kernel<kThreads><<<grid_for((n + 31) / 32, kThreads, Lc), kThreads, 0, Lc.stream>>>(out, n, value ? 0xFFFFFFFFu : 0u);
    return cudaGetLastError();
}

// ---- synthetic rasters (host mirror: erased_cells_b200/synth.py) ---------------------------------
template <class T>
__global__ void __launch_bounds__(kThreads) synth_kernel(T* __restrict__ out, size_t n, uint64_t seed, uint64_t off, int kind,
                                                         double lo, double hi, uint64_t period, T sentinel) {
    for (size_t i = blockIdx.x * size_t(kThreads) + threadIdx.x; i < n; i += size_t(gridDim.x) * kThreads) {
        const uint64_t h = splitmix64(seed ^ (off + i));
        T v;
        if (kind == 0) {  // uniform over the full bit range
            v = from_bits<T>(static_cast<bits_t<T>>(h));
        } else if (kind == 1) {  // uniform integer in [lo, hi]
            const int64_t a = static_cast<int64_t>(lo), b = static_cast<int64_t>(hi);
            const uint64_t span = static_cast<uint64_t>(b - a) + 1ull;
            const int64_t x = a + static_cast<int64_t>((h >> 11) % span);
            v = static_cast<T>(x);
        } else {  // uniform real in [lo, hi): lo + u * (hi - lo), u = (h >> 11) * 2^-53, each op rounded
            const double u = __dmul_rn(static_cast<double>(h >> 11), 0x1p-53);
            const double x = __dadd_rn(lo, __dmul_rn(u, __dsub_rn(hi, lo)));
            if constexpr (is_fp<T>) v = static_cast<T>(x);
            else v = static_cast<T>(static_cast<int64_t>(x));
        }
        if (period != 0 && splitmix64(h) % period == 0) v = sentinel;
        out[i] = v;
    }
}
cudaError_t launch_synth(const Launch& Lc, int ct, void* out, size_t n, uint64_t seed, uint64_t index_offset, int kind,
                         double lo, double hi, uint64_t period, uint64_t sentinel_bits) {
    const int grid = reduce_grid(n, kThreads * 4, Lc);
    switch (ct) {
#define X(id, p)                                                                                                    \
    case id:                                                                                                        \
        synth_kernel<p><<<grid, kThreads, 0, Lc.stream>>>(static_cast<p*>(out), n, seed, index_offset, kind, lo, hi, period, \
                                                          from_bits<p>(static_cast<bits_t<p>>(sentinel_bits)));    \
        break;
        EC_WITH_CT(X)
#undef X
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace ec
