// Statistics (count / mean / population stddev next to min_max) — an EXTENSION: the reference has no statistics
// beyond min_max and Mask::counts, the task's north star asks for them as reductions.
// The definition (restated on the CPU in oracle/oracle.py) makes the answer a function of the
// multiset of valid cells only, so it is bit-identical for any grid, any summation order and any number of GPUs:
//
//   y  = fl(fl(to_f64(cell) - pivot) * 2^-E)     |y| < 1   (pivot, E come from the min/max of the valid cells)
//   z  = fl(y * y)
//   each of y, z is cut into two 48-bit fixed-point windows — t1 = fl(v + 48) carries rint(v * 2^47) in its low
//   mantissa bits, the remainder r = v - (t1 - 48) is exact, t2 = fl(r + 48 * 2^-48) carries rint(r * 2^95) — and
//   the four integer streams are summed EXACTLY (128-bit two's complement).
//
// One streaming pass, algorithmic bytes per cell = size_of(T) (+ 1/8 with a mask); 12 FP64 operations per cell,
// which is what bounds the narrow cell types (see the kernel table in DESIGN.md).
#pragma once
#include "ec_reduce.cuh"

namespace ec {

constexpr int kMomentWords = 9;  // {count, X1.lo, X1.hi, X2.lo, X2.hi, Z1.lo, Z1.hi, Z2.lo, Z2.hi}

struct Acc128 {
    uint64_t lo;
    int64_t hi;
};
__device__ __forceinline__ void acc_add(Acc128& a, int64_t x) {
    const uint64_t ux = static_cast<uint64_t>(x);
    a.lo += ux;
    a.hi += (x >> 63) + (a.lo < ux ? 1 : 0);
}
__device__ __forceinline__ void acc_add(Acc128& a, const Acc128& b) {
    a.lo += b.lo;
    a.hi += b.hi + (a.lo < b.lo ? 1 : 0);
}
__device__ __forceinline__ Acc128 warp_sum(Acc128 v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Acc128 x;
        x.lo = __shfl_xor_sync(0xFFFFFFFFu, v.lo, o);
        x.hi = __shfl_xor_sync(0xFFFFFFFFu, v.hi, o);
        acc_add(v, x);
    }
    return v;
}

// raw mantissa bits of the two window sums of `v`, accumulated modulo 2^64 (the per-tile bias K1/K2 * cells is
// taken off once per tile, see below)
__device__ __forceinline__ void moment_windows(double v, uint64_t& w1, uint64_t& w2) {
    constexpr double C1 = 48.0;                   // ulp 2^-47
    constexpr double C2 = 48.0 * 0x1p-48;         // ulp 2^-95
    const double t1 = __dadd_rn(v, C1);
    const double r = __dsub_rn(v, __dsub_rn(t1, C1));
    const double t2 = __dadd_rn(r, C2);
    w1 += static_cast<uint64_t>(__double_as_longlong(t1));
    w2 += static_cast<uint64_t>(__double_as_longlong(t2));
}
constexpr uint64_t kMomentK1 = 0x4048000000000000ull;  // bits(48.0)
constexpr uint64_t kMomentK2 = 0x3D48000000000000ull;  // bits(48.0 * 2^-48)

template <class T, bool MASKED, int VB, int UNROLL, int THREADS>
__global__ void __launch_bounds__(THREADS) moments_kernel(const T* __restrict__ a, const uint32_t* __restrict__ m, size_t n,
                                                          double pivot, double scale, unsigned long long* __restrict__ acc) {
    constexpr int V = VB / sizeof(T);
    constexpr size_t TILE = size_t(THREADS) * V * UNROLL;
    constexpr uint32_t VMASK = V >= 32 ? 0xFFFFFFFFu : ((1u << (V & 31)) - 1u);
    const size_t full = n / TILE;
    Acc128 s[4] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
    uint64_t count = 0;

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
        uint64_t ts[4] = {0, 0, 0, 0};  // sums of raw bits over this thread's V * UNROLL cells, modulo 2^64
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
#pragma unroll
            for (int j = 0; j < V; ++j) {
                double y = __dmul_rn(__dsub_rn(as_f64(va[u].v[j]), pivot), scale);
                if constexpr (MASKED) y = ((w[u] >> j) & 1u) ? y : 0.0;  // a masked-out cell contributes 0 to every window
                moment_windows(y, ts[0], ts[1]);
                moment_windows(__dmul_rn(y, y), ts[2], ts[3]);
            }
            if constexpr (MASKED) count += __popc(w[u]);
        }
        // window sums of one tile fit int64 (64 cells x 2^47): take the bias off and widen
        constexpr uint64_t CELLS = uint64_t(V) * UNROLL;
        acc_add(s[0], static_cast<int64_t>(ts[0] - CELLS * kMomentK1));
        acc_add(s[1], static_cast<int64_t>(ts[1] - CELLS * kMomentK2));
        acc_add(s[2], static_cast<int64_t>(ts[2] - CELLS * kMomentK1));
        acc_add(s[3], static_cast<int64_t>(ts[3] - CELLS * kMomentK2));
        if constexpr (!MASKED) count += CELLS;
    }
    if (blockIdx.x == full % gridDim.x) {  // ragged tail, one cell per thread
        for (size_t i = full * TILE + threadIdx.x; i < n; i += THREADS) {
            bool valid = true;
            if constexpr (MASKED) valid = (m[i / 32] >> (i % 32)) & 1u;
            if (!valid) continue;
            const double y = __dmul_rn(__dsub_rn(as_f64(a[i]), pivot), scale);
            uint64_t ts[4] = {0, 0, 0, 0};
            moment_windows(y, ts[0], ts[1]);
            moment_windows(__dmul_rn(y, y), ts[2], ts[3]);
            acc_add(s[0], static_cast<int64_t>(ts[0] - kMomentK1));
            acc_add(s[1], static_cast<int64_t>(ts[1] - kMomentK2));
            acc_add(s[2], static_cast<int64_t>(ts[2] - kMomentK1));
            acc_add(s[3], static_cast<int64_t>(ts[3] - kMomentK2));
            ++count;
        }
    }

    // warp shuffle -> shared-memory block tree -> nine 64-bit atomics per CTA. Integer sums modulo 2^128: the carry
    // out of a low-word atomicAdd is recovered from the value it returns, so the total is exact in any order.
    __shared__ Acc128 sh[THREADS / 32][4];
    __shared__ uint64_t shc[THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 4; ++k) s[k] = warp_sum(s[k]);
    count = warp_sum(count);
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) sh[warp][k] = s[k];
        shc[warp] = count;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        Acc128 t = sh[0][threadIdx.x];
#pragma unroll
        for (int wv = 1; wv < THREADS / 32; ++wv) acc_add(t, sh[wv][threadIdx.x]);
        const unsigned long long old = atomicAdd(acc + 1 + 2 * threadIdx.x, static_cast<unsigned long long>(t.lo));
        const unsigned long long carry = (old + t.lo < old) ? 1ull : 0ull;
        atomicAdd(acc + 2 + 2 * threadIdx.x, static_cast<unsigned long long>(t.hi) + carry);
    } else if (threadIdx.x == 4) {
        uint64_t c = shc[0];
#pragma unroll
        for (int wv = 1; wv < THREADS / 32; ++wv) c += shc[wv];
        atomicAdd(acc, static_cast<unsigned long long>(c));
    }
}

// ---- 8/16-bit cells: the same raw sums from plain integer moments -------------------------------------------------
// For these types every step of the definition is exact (y is a multiple of 1/2 below 2^16, y*y needs 34 bits), the
// second windows are zero and the first ones are 2^(47-E) * sum(y) and 2^(47-2E) * sum(y^2). So the device only sums
// x and x^2 over the valid cells — packed dot products for 8-bit cells, 16x16 multiplies for 16-bit ones, no FP64 at
// all — and the host rewrites {count, A = sum x, B = sum x^2} into the window sums (ec_api.cu: moments_from_integer_sums).
// acc[5] = {count, A.lo, A.hi, B.lo, B.hi}, 128-bit two's complement.
template <class T, bool MASKED, int VB, int UNROLL, int THREADS>
__global__ void __launch_bounds__(THREADS) int_moments_kernel(const T* __restrict__ a, const uint32_t* __restrict__ m, size_t n,
                                                              unsigned long long* __restrict__ acc) {
    static_assert(sizeof(T) <= 2, "integer moments are for 8- and 16-bit cells");
    constexpr bool SG = std::is_signed<T>::value;
    constexpr int V = VB / sizeof(T);          // cells per 32-byte load
    constexpr int W = VB / 4;                  // 32-bit words per load
    constexpr int CPW = 4 / sizeof(T);         // cells per word
    constexpr size_t TILE = size_t(THREADS) * V * UNROLL;
    constexpr uint32_t VMASK = V >= 32 ? 0xFFFFFFFFu : ((1u << (V & 31)) - 1u);
    const size_t full = n / TILE;
    int64_t sum = 0;
    uint64_t sq = 0, count = 0;

    for (size_t t = blockIdx.x; t < full; t += gridDim.x) {
        const size_t base = t * TILE + size_t(threadIdx.x) * V;
        Vec<uint32_t, W> w[UNROLL];
        uint32_t mw[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const size_t c = base + size_t(u) * THREADS * V;
            w[u] = ld_stream<uint32_t, W>(reinterpret_cast<const uint32_t*>(a + c));
            if constexpr (MASKED) mw[u] = (__ldg(m + c / 32) >> (c % 32)) & VMASK;
        }
        int32_t ts = 0;      // |sum| of one thread's tile share: 128 cells x 2^15
        uint32_t tq8 = 0;    // 8-bit: 128 cells x 2^16
        uint64_t tq16 = 0;   // 16-bit squares reach 2^32 each
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
#pragma unroll
            for (int j = 0; j < W; ++j) {
                uint32_t x = w[u].v[j];
                if constexpr (MASKED) {
                    const uint32_t b = (mw[u] >> (CPW * j)) & ((1u << CPW) - 1u);
                    uint32_t sel;
                    if constexpr (sizeof(T) == 1) sel = ((b * 0x00204081u) & 0x01010101u) * 0xFFu;  // 4 bits -> 4 byte masks
                    else sel = ((b & 1u) * 0xFFFFu) | ((b >> 1) * 0xFFFF0000u);
                    x &= sel;  // a masked-out cell becomes 0: contributes nothing to either sum
                }
                if constexpr (sizeof(T) == 1) {
                    if constexpr (SG) {
                        ts = __dp4a(static_cast<int>(x), 0x01010101, ts);
                        tq8 = static_cast<uint32_t>(__dp4a(static_cast<int>(x), static_cast<int>(x), static_cast<int>(tq8)));
                    } else {
                        ts = static_cast<int32_t>(__dp4a(x, 0x01010101u, static_cast<uint32_t>(ts)));
                        tq8 = __dp4a(x, x, tq8);
                    }
                } else {
                    if constexpr (SG) {
                        const int lo = static_cast<int>(x << 16) >> 16, hi = static_cast<int>(x) >> 16;
                        ts += lo + hi;
                        tq16 += static_cast<uint64_t>(static_cast<uint32_t>(lo * lo)) + static_cast<uint32_t>(hi * hi);
                    } else {
                        const uint32_t lo = x & 0xFFFFu, hi = x >> 16;
                        ts += static_cast<int32_t>(lo + hi);
                        tq16 += static_cast<uint64_t>(lo * lo) + static_cast<uint64_t>(hi * hi);
                    }
                }
            }
            if constexpr (MASKED) count += __popc(mw[u]);
        }
        sum += ts;
        sq += sizeof(T) == 1 ? static_cast<uint64_t>(tq8) : tq16;
        if constexpr (!MASKED) count += uint64_t(V) * UNROLL;
    }
    if (blockIdx.x == full % gridDim.x) {  // ragged tail, one cell per thread
        for (size_t i = full * TILE + threadIdx.x; i < n; i += THREADS) {
            bool valid = true;
            if constexpr (MASKED) valid = (m[i / 32] >> (i % 32)) & 1u;
            if (!valid) continue;
            const int64_t x = static_cast<int64_t>(a[i]);
            sum += x;
            sq += static_cast<uint64_t>(x * x);
            ++count;
        }
    }
    // per-CTA totals fit 64 bits (a CTA sees fewer than 2^28 cells of at most 2^32 each); the cross-CTA sums are
    // 128-bit through carry-tracking atomics, as in moments_kernel
    sum = static_cast<int64_t>(warp_sum(static_cast<uint64_t>(sum)));
    sq = warp_sum(sq);
    count = warp_sum(count);
    __shared__ uint64_t sh[THREADS / 32][3];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { sh[warp][0] = static_cast<uint64_t>(sum); sh[warp][1] = sq; sh[warp][2] = count; }
    __syncthreads();
    if (threadIdx.x < 3) {
        uint64_t t = sh[0][threadIdx.x];
#pragma unroll
        for (int wv = 1; wv < THREADS / 32; ++wv) t += sh[wv][threadIdx.x];
        if (threadIdx.x == 2) {
            atomicAdd(acc, static_cast<unsigned long long>(t));
        } else {
            unsigned long long* dst = acc + 1 + 2 * threadIdx.x;
            const unsigned long long old = atomicAdd(dst, static_cast<unsigned long long>(t));
            const unsigned long long carry = (old + t < old) ? 1ull : 0ull;
            const unsigned long long ext = (threadIdx.x == 0 && static_cast<int64_t>(t) < 0) ? ~0ull : 0ull;  // sign of A
            atomicAdd(dst + 1, ext + carry);
        }
    }
}

}  // namespace ec
