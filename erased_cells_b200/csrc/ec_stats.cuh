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
// One streaming pass, algorithmic bytes per cell = size_of(T) (+ 1/8 with a mask); 12 FP64 operations per cell, hidden
// behind the HBM traffic of 64-bit cells (UInt64, Int64, Float64 — the only types that come here). Integer cells of at
// most 32 bits and Float32 cells take int_stats_kernel below: exact integer sums, no FP64.
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
// to_f64 for finite cells: the moments pass only runs when no valid cell is NaN or infinite, and a masked-out cell's
// value is discarded, so the payload-preserving f32 widening of as_f64 is not needed here
template <class T> __device__ __forceinline__ double finite_f64(T v) {
    if constexpr (std::is_same<T, float>::value) return static_cast<double>(v);
    else return as_f64(v);
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
                double y = __dmul_rn(__dsub_rn(finite_f64(va[u].v[j]), pivot), scale);
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
            const double y = __dmul_rn(__dsub_rn(finite_f64(a[i]), pivot), scale);
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

// ---- integer cells of at most 32 bits: everything in ONE pass, no FP64 ---------------------------------------------
// For these types y = x - pivot is exact, so the definition asks for the correctly rounded sums of y and y^2 — which
// follow on the host from {count, A = sum x, B = sum x^2} (exact integers) once the pivot is known. The device
// therefore needs no pivot: one read of the cells yields min, max, count, A and B, i.e. the whole ec_buf_statistics.
// 8-bit cells use packed dot products, 16-bit ones 16x16 multiplies, 32-bit ones one wide multiply whose two halves
// are summed separately (no carry chain); min/max run on biased unsigned values, two 16-bit lanes per register for
// the narrow types (as min_max_kernel does); a masked-out cell is forced to zero / to the identity of min and max.
// acc[7] = {count, A.lo, A.hi, B.lo, B.hi, min (biased, preset to all ones), max (biased, preset to 0)};
// A and B are 128-bit two's complement, summed across CTAs with carry-tracking atomics.
//
// QUANT: Float32 cells take the same route. A f32 raster whose valid cells are all below 2^E in magnitude is read as
// the int32 raster q = rint(x * 2^(26 - E)) (|q| <= 2^26, round half to even): cells within 2^3 of the largest
// magnitude are exact on that grid (their own ulp is coarser), smaller ones are rounded by at most 2^-27 of the
// largest magnitude — an eighth of the ulp the large cells carry. min, max, count, sum q and sum q^2 are then exact
// integers like for any int32 raster, and the host turns them into mean / stddev and scales by 2^(E - 26). Two FP32
// multiplies (the scale is split so that neither factor leaves the f32 range) and one cvt.rni per cell replace the
// twelve FP64 operations of the window route, which is what kept Float32 statistics off the HBM roofline.
template <class T, bool MASKED, int VB, int UNROLL, int THREADS, bool QUANT = false>
__global__ void __launch_bounds__(THREADS) int_stats_kernel(const T* __restrict__ a, const uint32_t* __restrict__ m, size_t n,
                                                            unsigned long long* __restrict__ acc, float qs1 = 1.0f, float qs2 = 1.0f) {
    static_assert(sizeof(T) <= 4 && std::is_integral<T>::value, "integer cells of at most 32 bits");
    static_assert(!QUANT || std::is_same<T, int32_t>::value, "quantised Float32 cells are read as int32");
    constexpr bool SG = std::is_signed<T>::value;
    constexpr int V = VB / sizeof(T);
    constexpr int W = VB / 4;
    constexpr int CPW = 4 / sizeof(T);
    constexpr size_t TILE = size_t(THREADS) * V * UNROLL;
    constexpr uint32_t VMASK = V >= 32 ? 0xFFFFFFFFu : ((1u << (V & 31)) - 1u);
    constexpr uint32_t BIAS = !SG ? 0u : (sizeof(T) == 1 ? 0x80808080u : (sizeof(T) == 2 ? 0x80008000u : 0x80000000u));
    constexpr uint32_t LB = !SG ? 0u : (sizeof(T) == 1 ? 0x80u : (sizeof(T) == 2 ? 0x8000u : 0x80000000u));
    const size_t full = n / TILE;
    int64_t sum = 0;
    uint64_t sum_u = 0;                      // 32-bit cells: zero-extended (biased when signed) sum of the tiled part
    uint64_t sq = 0, sq_hi = 0, count = 0;   // sq_hi: upper halves of the 32-bit cells' squares, weight 2^32
    uint32_t pmin = 0xFFFFFFFFu, pmax = 0u;  // 8/16-bit: two 16-bit lanes; 32-bit: one biased value

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
        int32_t ts = 0;      // 8/16-bit: one thread's share of a tile, at most 128 cells x 2^15
        uint32_t tq8 = 0;    // 8-bit squares: 128 cells x 2^16
        uint64_t tq16 = 0;   // 16-bit squares reach 2^32 each
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
#pragma unroll
            for (int j = 0; j < W; ++j) {
                uint32_t x = w[u].v[j];
                if constexpr (QUANT) x = static_cast<uint32_t>(__float2int_rn(__fmul_rn(__fmul_rn(__uint_as_float(x), qs1), qs2)));
                uint32_t kmin = x ^ BIAS, kmax = x ^ BIAS;
                if constexpr (MASKED) {
                    [[maybe_unused]] const uint32_t b = (mw[u] >> (CPW * j)) & ((1u << CPW) - 1u);
                    uint32_t sel;
                    if constexpr (sizeof(T) == 1) sel = ((b * 0x00204081u) & 0x01010101u) * 0xFFu;  // 4 bits -> 4 byte masks
                    else if constexpr (sizeof(T) == 2) {
                        // 2 bits -> 2 lane masks: one multiply parks the two bits in the sign positions of bytes 2 and 3,
                        // prmt's sign-replicate mode fans them out (the ALU pipe is what limits this kernel)
                        const uint32_t v = (mw[u] & (3u << (2 * j))) * ((1u << (23 - 2 * j)) | (1u << (30 - 2 * j)));
                        asm("prmt.b32 %0, %1, %1, 0xBBAA;" : "=r"(sel) : "r"(v));
                    }
                    else sel = 0u - b;
                    x &= sel;
                    kmin |= ~sel;
                    kmax &= sel;
                }
                if constexpr (sizeof(T) == 1) {
                    pmin = __vimin3_u16x2(pmin, __byte_perm(kmin, 0u, 0x4240), __byte_perm(kmin, 0u, 0x4341));
                    pmax = __vimax3_u16x2(pmax, __byte_perm(kmax, 0u, 0x4240), __byte_perm(kmax, 0u, 0x4341));
                    if constexpr (SG) {
                        ts = __dp4a(static_cast<int>(x), 0x01010101, ts);
                        tq8 = static_cast<uint32_t>(__dp4a(static_cast<int>(x), static_cast<int>(x), static_cast<int>(tq8)));
                    } else {
                        ts = static_cast<int32_t>(__dp4a(x, 0x01010101u, static_cast<uint32_t>(ts)));
                        tq8 = __dp4a(x, x, tq8);
                    }
                } else if constexpr (sizeof(T) == 2) {
                    pmin = __vminu2(pmin, kmin);
                    pmax = __vmaxu2(pmax, kmax);
                    // sums run over the biased unsigned lanes u = x ^ 0x8000 (kmax; 0 for a masked-out lane) — no sign
                    // extension; signed cells are put right after the loop: x = u - 2^15
                    const uint32_t lo = kmax & 0xFFFFu, hi = kmax >> 16;
                    ts = static_cast<int32_t>(__dp2a_lo(kmax, 0x00000101u, static_cast<uint32_t>(ts)));
                    tq16 = static_cast<uint64_t>(lo) * lo + tq16;
                    tq16 = static_cast<uint64_t>(hi) * hi + tq16;
                } else {
                    pmin = min(pmin, kmin);
                    pmax = max(pmax, kmax);
                    // signed cells are summed biased (kmax = (x ^ 2^31) & sel, zero for a masked-out cell) and squared
                    // through |x|: zero-extending adds and one unsigned 32x32->64 multiply per cell
                    uint32_t mag = x;
                    if constexpr (SG) { sum_u += kmax; mag = static_cast<uint32_t>(abs(static_cast<int>(x))); }
                    else sum_u += x;
                    const uint64_t q = static_cast<uint64_t>(mag) * mag;
                    sq += q & 0xFFFFFFFFull;
                    sq_hi += q >> 32;
                }
            }
            if constexpr (MASKED) count += __popc(mw[u]);
        }
        if constexpr (sizeof(T) <= 2) {
            sum += ts;
            sq += sizeof(T) == 1 ? static_cast<uint64_t>(tq8) : tq16;
        }
        if constexpr (!MASKED) count += uint64_t(V) * UNROLL;
    }
    uint32_t lo, hi;
    if constexpr (sizeof(T) <= 2) {
        lo = min(pmin & 0xFFFFu, pmin >> 16); hi = max(pmax & 0xFFFFu, pmax >> 16);
        if constexpr (SG && sizeof(T) == 2) {  // sums were taken over u = x + 2^15: x = u - 2^15, x^2 = u^2 - 2^16 u + 2^30
            sq = sq - (static_cast<uint64_t>(sum) << 16) + (count << 30);
            sum -= static_cast<int64_t>(count << 15);
        }
    } else {
        lo = pmin; hi = pmax;
        sum = static_cast<int64_t>(sum_u) - (SG ? static_cast<int64_t>(count << 31) : 0);  // take the bias off: 2^31 per valid cell
    }
    if (blockIdx.x == full % gridDim.x) {  // ragged tail, one cell per thread
        for (size_t i = full * TILE + threadIdx.x; i < n; i += THREADS) {
            bool valid = true;
            if constexpr (MASKED) valid = (m[i / 32] >> (i % 32)) & 1u;
            if (!valid) continue;
            T cell = a[i];
            if constexpr (QUANT) cell = static_cast<T>(__float2int_rn(__fmul_rn(__fmul_rn(__int_as_float(static_cast<int>(cell)), qs1), qs2)));
            const int64_t x = static_cast<int64_t>(cell);
            const uint32_t k = static_cast<uint32_t>(static_cast<bits_t<T>>(cell)) ^ LB;
            lo = min(lo, k);
            hi = max(hi, k);
            sum += x;
            const uint64_t q = static_cast<uint64_t>(x * x);
            sq += q & 0xFFFFFFFFull;
            sq_hi += q >> 32;
            ++count;
        }
    }
    // per-CTA totals fit 64 bits (a CTA sees fewer than 2^28 cells, see the launcher)
    sum = static_cast<int64_t>(warp_sum(static_cast<uint64_t>(sum)));
    sq = warp_sum(sq);
    sq_hi = warp_sum(sq_hi);
    count = warp_sum(count);
    lo = warp_min(lo);
    hi = warp_max(hi);
    __shared__ uint64_t sh[THREADS / 32][4];
    __shared__ uint32_t shk[THREADS / 32][2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        sh[warp][0] = static_cast<uint64_t>(sum); sh[warp][1] = sq; sh[warp][2] = sq_hi; sh[warp][3] = count;
        shk[warp][0] = lo; shk[warp][1] = hi;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        uint64_t t = sh[0][threadIdx.x];
#pragma unroll
        for (int wv = 1; wv < THREADS / 32; ++wv) t += sh[wv][threadIdx.x];
        if (threadIdx.x == 3) {
            atomicAdd(acc, static_cast<unsigned long long>(t));
        } else {
            // 128-bit add of {A: sign-extended t | B: t | B: t * 2^32} with the carry taken from the low word's old value
            unsigned long long add_lo = t, add_hi = 0;
            if (threadIdx.x == 0) add_hi = static_cast<int64_t>(t) < 0 ? ~0ull : 0ull;
            if (threadIdx.x == 2) { add_lo = t << 32; add_hi = t >> 32; }
            unsigned long long* dst = acc + (threadIdx.x == 0 ? 1 : 3);
            const unsigned long long old = atomicAdd(dst, add_lo);
            const unsigned long long carry = (old + add_lo < old) ? 1ull : 0ull;
            atomicAdd(dst + 1, add_hi + carry);
        }
    } else if (threadIdx.x == 4) {
        uint32_t k = shk[0][0];
#pragma unroll
        for (int wv = 1; wv < THREADS / 32; ++wv) k = min(k, shk[wv][0]);
        atomicMin(acc + 5, static_cast<unsigned long long>(k));
    } else if (threadIdx.x == 5) {
        uint32_t k = shk[0][1];
#pragma unroll
        for (int wv = 1; wv < THREADS / 32; ++wv) k = max(k, shk[wv][1]);
        atomicMax(acc + 6, static_cast<unsigned long long>(k));
    }
}

}  // namespace ec
