// Reductions and mask kernels: min_max (10 types x masked/unmasked), popcount, first difference,
// mask build / unpack / bit ops / fill, synthetic rasters.
#include "ec_internal.hpp"
#include "ec_mask.cuh"
#include "ec_reduce.cuh"

#include <limits>

#ifndef EC_VB
#define EC_VB 32
#endif
#ifndef EC_RUNROLL
#define EC_RUNROLL 4
#endif

namespace ec {

// Geometry picked from the tools/ubench sweeps on B200 at 2^24, 2^26 and 2^28 cells (profiles/r01_ubench_*):
// unmasked min_max: 256-thread CTAs, 4 x 32-byte loads in flight per thread, persistent grid of 4 CTAs per SM
// (3.8 / 5.9 / 6.96 TB/s for f32 at the three sizes; larger persistent grids must match the resident capacity
// exactly or lose a partial wave); masked: 512 threads, 2 loads in flight, 8 CTAs per SM (6.77 TB/s at 2^28).
constexpr int kRedThreads = 256;
constexpr int kRedUnroll = 4;
constexpr int kRedCtasPerSm = 4;
constexpr int kRedThreadsMasked = 512;
constexpr int kRedUnrollMasked = 2;
constexpr int kRedCtasPerSmMasked = 8;
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
        constexpr size_t TILE = size_t(kRedThreadsMasked) * V * kRedUnrollMasked;
        return launch_k(Lc, min_max_kernel<T, true, EC_VB, kRedUnrollMasked, kRedThreadsMasked>, reduce_grid(n, TILE, Lc, kRedCtasPerSmMasked),
                        kRedThreadsMasked, static_cast<const T*>(a), mask, n, smin, smax, s);
    } else {
        constexpr size_t TILE = size_t(kRedThreads) * V * kRedUnroll;
        return launch_k(Lc, min_max_kernel<T, false, EC_VB, kRedUnroll, kRedThreads>, reduce_grid(n, TILE, Lc, kRedCtasPerSm), kRedThreads,
                        static_cast<const T*>(a), nullptr, n, smin, smax, s);
    }
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

cudaError_t launch_popcount(const Launch& Lc, const uint32_t* words, size_t nwords, const ReduceScratch& s, uint64_t second_word) {
    const int grid = reduce_grid(nwords / 4, kThreads, Lc);
    return launch_k(Lc, popcount_kernel<kThreads>, grid, kThreads, words, nwords, s, second_word);
}

cudaError_t launch_exchange_sums(const Launch& Lc, const unsigned long long* local, int words, int pairs, const PeerExchange& px,
                                 unsigned long long region_off, uint64_t* host_words, uint64_t host_seq) {
    exchange_sums_kernel<256><<<1, 256, 0, Lc.stream>>>(local, words, pairs, px, region_off, host_words, host_seq);
    return cudaGetLastError();
}

template <class U> static cudaError_t first_diff_u(const Launch& Lc, const void* a, const void* b, size_t n, const ReduceScratch& s) {
    const int grid = reduce_grid(n / (EC_VB / sizeof(U)), kThreads, Lc);
    return launch_k(Lc, first_diff_kernel<U, EC_VB, kThreads>, grid, kThreads, static_cast<const U*>(a), static_cast<const U*>(b), n, s);
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
static cudaError_t mask_build_u(const Launch& Lc, const void* a, size_t n, uint64_t sentinel, uint32_t* out, const MaskCount& mc) {
    constexpr int V0 = EC_VB / sizeof(U);
    constexpr int V = V0 > 32 ? 32 : V0;
    constexpr size_t TILE = size_t(kThreads) * V * EC_RUNROLL;
    return launch_k(Lc, mask_build_kernel<U, PACK, EC_VB, EC_RUNROLL, kThreads>, grid_for(n, TILE, Lc), kThreads,
                    static_cast<const U*>(a), n, static_cast<U>(sentinel), out, mc);
}
cudaError_t launch_mask_build(const Launch& Lc, int cell_bytes, const void* a, size_t n, uint64_t sentinel_bits,
                              bool pack_bools, uint32_t* out, const MaskCount& mc) {
    if (pack_bools) return mask_build_u<uint8_t, true>(Lc, a, n, 0, out, mc);
    switch (cell_bytes) {
        case 1: return mask_build_u<uint8_t, false>(Lc, a, n, sentinel_bits, out, mc);
        case 2: return mask_build_u<uint16_t, false>(Lc, a, n, sentinel_bits, out, mc);
        case 4: return mask_build_u<uint32_t, false>(Lc, a, n, sentinel_bits, out, mc);
        default: return mask_build_u<uint64_t, false>(Lc, a, n, sentinel_bits, out, mc);
    }
}
cudaError_t launch_mask_unpack(const Launch& Lc, const uint32_t* m, size_t n, uint8_t* out) {
    mask_unpack_kernel<kThreads><<<grid_for(n, size_t(kThreads) * 16, Lc), kThreads, 0, Lc.stream>>>(m, n, out);
    return cudaGetLastError();
}
cudaError_t launch_mask_bitop(const Launch& Lc, int mop, const uint32_t* l, const uint32_t* r, size_t n, uint32_t* out, const MaskCount& mc) {
    return launch_k(Lc, mask_bitop_kernel<kThreads>, grid_for((n + 127) / 128, kThreads, Lc), kThreads, mop, l, r, n, out, mc);
}
cudaError_t launch_mask_fill(const Launch& Lc, uint32_t* out, size_t n, bool value) {
    return launch_k(Lc, mask_fill_kernel<kThreads>, grid_for((n + 31) / 32, kThreads, Lc), kThreads, out, n, value ? 0xFFFFFFFFu : 0u);
}

// ---- Extend<C> (src/buffer.rs:205-221): value-checked cast `c.into_cell_value().to_<p>().unwrap()` ---------
// num-traits chain: ints go through i64/u64 with range checks, floats truncate iff inside the target's range,
// anything -> f32 goes through f64. A cell that does not fit raises the flag (the reference panics).
template <class S, class D> __device__ __forceinline__ bool checked_conv(S v, D& o) {
    if constexpr (std::is_same<D, double>::value) {
        o = as_f64(v);
        return true;
    } else if constexpr (std::is_same<D, float>::value) {
        const double x = as_f64(v);
        if (x != x) {  // cvtsd2ss: sign, quiet bit, top payload bits
            const uint64_t b = static_cast<uint64_t>(__double_as_longlong(x));
            o = __uint_as_float(static_cast<uint32_t>((b >> 32) & 0x80000000u) | 0x7FC00000u | static_cast<uint32_t>((b >> 29) & 0x003FFFFFu));
        } else {
            o = __double2float_rn(x);
        }
        return true;
    } else if constexpr (std::is_signed<D>::value) {
        int64_t t;
        if constexpr (is_fp<S>) {
            const double x = static_cast<double>(v);
            if (!(x >= -9223372036854775808.0 && x < 9223372036854775808.0)) return false;
            t = __double2ll_rz(x);
        } else if constexpr (std::is_same<S, uint64_t>::value) {
            if (v > static_cast<uint64_t>(INT64_MAX)) return false;
            t = static_cast<int64_t>(v);
        } else {
            t = static_cast<int64_t>(v);
        }
        constexpr int64_t hi = sizeof(D) == 8 ? INT64_MAX : (int64_t(1) << (8 * sizeof(D) - 1)) - 1, lo = -hi - 1;
        if (t < lo || t > hi) return false;
        o = static_cast<D>(t);
        return true;
    } else {
        uint64_t t;
        if constexpr (is_fp<S>) {
            const double x = static_cast<double>(v);
            if (!(x > -1.0 && x < 18446744073709551616.0)) return false;
            t = __double2ull_rz(x);
        } else if constexpr (std::is_signed<S>::value) {
            if (v < 0) return false;
            t = static_cast<uint64_t>(v);
        } else {
            t = static_cast<uint64_t>(v);
        }
        constexpr uint64_t hi = sizeof(D) == 8 ? UINT64_MAX : (uint64_t(1) << (8 * sizeof(D))) - 1;
        if (t > hi) return false;
        o = static_cast<D>(t);
        return true;
    }
}
template <class S>
__global__ void __launch_bounds__(kThreads) checked_cast_kernel(const S* __restrict__ a, int dct, void* __restrict__ out, size_t n,
                                                                unsigned int* __restrict__ fail) {
    bool ok = true;
    for (size_t i = blockIdx.x * size_t(kThreads) + threadIdx.x; i < n; i += size_t(gridDim.x) * kThreads) {
        const S v = a[i];
        switch (dct) {
#define X(id, p) case id: ok &= checked_conv<S, p>(v, static_cast<p*>(out)[i]); break;
            EC_WITH_CT(X)
#undef X
        }
    }
    if (!ok) atomicOr(fail, 1u);
}
cudaError_t launch_checked_cast(const Launch& Lc, int sct, const void* a, int dct, void* out, size_t n, unsigned int* fail) {
    const int grid = reduce_grid(n, kThreads * 4, Lc);
    switch (sct) {
#define X(id, p) case id: checked_cast_kernel<p><<<grid, kThreads, 0, Lc.stream>>>(static_cast<const p*>(a), dct, out, n, fail); break;
        EC_WITH_CT(X)
#undef X
        default: return cudaErrorInvalidValue;
    }
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
