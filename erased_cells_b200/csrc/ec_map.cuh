// Streaming map kernels: one pass load -> promote -> op -> store (reference hot loops
// src/buffer.rs:324-329, :346-352, :360-371, :150-167 and src/masked/masked_buffer.rs:137-152,
// :326-336). All are HBM-bound: algorithmic bytes per cell = sum of the operand sizes + output size.
//
// Geometry: a CTA of THREADS threads walks tiles of THREADS*V*UNROLL cells (grid-stride, persistent
// grid sized from the SM count). V = VB / widest cell so that the widest stream moves VB (32 or 16)
// bytes per thread per access; narrower streams use proportionally narrower accesses, which stay
// sector-exact because a warp still covers >= 32*V contiguous cells. All UNROLL loads are issued
// before the first use so each thread keeps UNROLL independent requests per operand in flight.
#pragma once
#include "ec_common.cuh"

namespace ec {

template <int A, int B> __host__ __device__ constexpr int cmax() { return A > B ? A : B; }

// ---- functors: per-cell semantics ------------------------------------------------------------
template <class L, class R, int OP> struct BinaryF {  // src/value.rs:199-209
    using A = L; using B = R; using O = double;
    __device__ __forceinline__ double operator()(L a, R b) const {
        if constexpr (OP == OP_DIV) return f64_div_cells<L, R>(as_f64(a), as_f64(b));
        else return f64_op<OP, is_fp<L>, is_fp<R>>(as_f64(a), as_f64(b));
    }
};
template <class L> struct ScalarF {  // src/buffer.rs:346-352; rhs is `s as f64`, converted once on the host
    using A = L; using O = double;
    int op; double s;
    __device__ __forceinline__ double operator()(L a) const { return f64_op_rt<true, true>(op, as_f64(a), s); }
};
template <class T> struct NegF {  // src/value.rs:224-240
    using A = T; using O = typename neg_out<T>::type;
    __device__ __forceinline__ O operator()(T a) const { return neg_cell(a); }
};
template <class S, class D> struct CastF {  // src/value.rs:74-98
    using A = S; using O = D;
    __device__ __forceinline__ D operator()(S a) const { return cast_cell<S, D>(a); }
};
// `(a - b) / (a + b)` with each op rounded on its own — the NDVI chain of src/gdal/rasterband.rs:148
template <class L, class R> struct NormDiffF {
    using A = L; using B = R; using O = double;
    __device__ __forceinline__ double operator()(L a, R b) const {
        const double x = as_f64(a), y = as_f64(b);
        const double num = f64_op<OP_SUB, is_fp<L>, is_fp<R>>(x, y);
        const double den = f64_op<OP_ADD, is_fp<L>, is_fp<R>>(x, y);
        // integer bands: num and den are exact (or, for 64-bit cells, rounded) integers within 2^65 -> the guard-free quotient
        return f64_div_cells<L, R>(num, den);
    }
};
// `(l op1 r) op2 s`
template <class L, class R> struct BinaryScalarF {
    using A = L; using B = R; using O = double;
    int op1, op2; double s;
    __device__ __forceinline__ double operator()(L a, R b) const {
        const double t = f64_op_rt<is_fp<L>, is_fp<R>>(op1, as_f64(a), as_f64(b));
        return f64_op_rt<true, true>(op2, t, s);
    }
};

// the same with both ops known at compile time (one division body at most per op instead of a 4-way switch each):
// instantiated for same-typed operands and the README's u8/u16 pair (ec_tu_fused.cu), the rest use the runtime form
template <class L, class R, int OP1, int OP2> struct BinaryScalarT {
    using A = L; using B = R; using O = double;
    double s;
    __device__ __forceinline__ double operator()(L a, R b) const {
        double t;
        if constexpr (OP1 == OP_DIV) t = f64_div_cells<L, R>(as_f64(a), as_f64(b));
        else t = f64_op<OP1, is_fp<L>, is_fp<R>>(as_f64(a), as_f64(b));
        return f64_op<OP2, true, true>(t, s);
    }
};

// `(a op1 s1) op2 s2` — scale-and-offset chains (`dn * 0.0001 + 273.15`), both ops at compile time
template <class L, int OP1, int OP2> struct ScalarScalarT {
    using A = L; using O = double;
    double s1, s2;
    __device__ __forceinline__ double operator()(L a) const {
        const double t = f64_op<OP1, true, true>(as_f64(a), s1);
        return f64_op<OP2, true, true>(t, s2);
    }
};

// ---- one-input map -----------------------------------------------------------------------------
template <class F, int VB, int UNROLL, int THREADS>
__global__ void __launch_bounds__(THREADS) map1_kernel(const typename F::A* __restrict__ a,
                                                       typename F::O* __restrict__ o, size_t n, F f) {
    using A = typename F::A; using O = typename F::O;
    constexpr int V = VB / cmax<sizeof(A), sizeof(O)>();
    constexpr size_t TILE = size_t(THREADS) * V * UNROLL;
    const size_t full = n / TILE;
    overlap_prologue();
    for (size_t t = blockIdx.x; t < full; t += gridDim.x) {
        const size_t base = t * TILE + size_t(threadIdx.x) * V;
        Vec<A, V> va[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) va[u] = ld_stream<A, V>(a + base + size_t(u) * THREADS * V);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            Vec<O, V> vo;
#pragma unroll
            for (int j = 0; j < V; ++j) vo.v[j] = f(va[u].v[j]);
            st_stream<O, V>(o + base + size_t(u) * THREADS * V, vo);
        }
    }
    if (blockIdx.x == full % gridDim.x) {  // ragged tail (< TILE cells), one cell per thread
        for (size_t i = full * TILE + threadIdx.x; i < n; i += THREADS) o[i] = f(a[i]);
    }
}

// ---- two-input map, optionally carrying the validity mask (packed words) along ---------------------
// The MASKED flavour also ANDs the tile's mask words (32 cells per word) in the CTA that owns the tile and counts the
// set bits it wrote: MaskedCellBuffer op (src/masked/masked_buffer.rs:326-336) plus Mask::counts of the result in one
// launch. It is a separate instantiation so that the unmasked kernels (issue-limited for the division) carry none of it.
// Register budget: with narrow operands (<= 8 input bytes per cell) the loads in flight are small, so the
// kernel is held to 64 registers (>= 1024 resident threads per SM) — the f64 division path otherwise drifts
// to 66 registers = 3 CTAs/SM in the fused functors (ncu, profiles/). Wide operands keep UNROLL x 32 bytes
// per operand in registers and get the 128-register budget instead.
template <class F> constexpr int map2_min_ctas(int threads) {
    return (sizeof(typename F::A) + sizeof(typename F::B) <= 8 ? 1024 : 512) / threads;
}
template <class F, int VB, int UNROLL, int THREADS, bool MASKED = false>
__global__ void __launch_bounds__(THREADS, map2_min_ctas<F>(THREADS)) map2_kernel(const typename F::A* __restrict__ a,
                                                       const typename F::B* __restrict__ b,
                                                       typename F::O* __restrict__ o, size_t n, F f,
                                                       const uint32_t* __restrict__ lm,
                                                       const uint32_t* __restrict__ rm,
                                                       uint32_t* __restrict__ om, MaskCount mc) {
    using A = typename F::A; using B = typename F::B; using O = typename F::O;
    constexpr int V = VB / cmax<cmax<sizeof(A), sizeof(B)>(), sizeof(O)>();
    constexpr size_t TILE = size_t(THREADS) * V * UNROLL;
    static_assert(TILE % 128 == 0, "a tile must cover whole 16-byte groups of mask words");
    constexpr int TILE_WORDS = TILE / 32;
    const size_t full = n / TILE;
    unsigned int ones = 0;  // set bits this thread wrote into the result mask
    overlap_prologue();
    for (size_t t = blockIdx.x; t < full; t += gridDim.x) {
        const size_t base = t * TILE + size_t(threadIdx.x) * V;
        Vec<A, V> va[UNROLL];
        Vec<B, V> vb[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            va[u] = ld_stream<A, V>(a + base + size_t(u) * THREADS * V);
            vb[u] = ld_stream<B, V>(b + base + size_t(u) * THREADS * V);
        }
        uint4 wl, wr;
        const bool mask_lane = MASKED && threadIdx.x < TILE_WORDS / 4;
        if constexpr (MASKED) {
            if (mask_lane) {
                wl = *reinterpret_cast<const uint4*>(lm + t * TILE_WORDS + threadIdx.x * 4);
                wr = *reinterpret_cast<const uint4*>(rm + t * TILE_WORDS + threadIdx.x * 4);
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            Vec<O, V> vo;
#pragma unroll
            for (int j = 0; j < V; ++j) vo.v[j] = f(va[u].v[j], vb[u].v[j]);
            st_stream<O, V>(o + base + size_t(u) * THREADS * V, vo);
        }
        if constexpr (MASKED) {
            if (mask_lane) {
                const uint4 w = make_uint4(wl.x & wr.x, wl.y & wr.y, wl.z & wr.z, wl.w & wr.w);
                *reinterpret_cast<uint4*>(om + t * TILE_WORDS + threadIdx.x * 4) = w;
                ones += __popc(w.x) + __popc(w.y) + __popc(w.z) + __popc(w.w);
            }
        }
    }
    if (blockIdx.x == full % gridDim.x) {
        for (size_t i = full * TILE + threadIdx.x; i < n; i += THREADS) o[i] = f(a[i], b[i]);
        if constexpr (MASKED) {
            const size_t words = (n + 31) / 32;
            for (size_t w = full * TILE_WORDS + threadIdx.x; w < words; w += THREADS) {
                const uint32_t x = lm[w] & rm[w];
                om[w] = x;
                ones += __popc(x);
            }
        }
    }
    if constexpr (MASKED) {
        if (mc.acc != nullptr) {  // uniform over the grid: Mask::counts of the result comes for free
            const unsigned long long c = block_count<THREADS>(ones);
            if (threadIdx.x == 0) publish_count(mc, c);
        }
    }
}

// ---- fill: `vec![v; len]` (src/buffer.rs:79-88) — write-only ---------------------------------------
template <class T, int VB, int THREADS>
__global__ void __launch_bounds__(THREADS) fill_kernel(T* __restrict__ o, size_t n, T value) {
    constexpr int V = VB / sizeof(T);
    constexpr size_t TILE = size_t(THREADS) * V;
    Vec<T, V> vv;
#pragma unroll
    for (int j = 0; j < V; ++j) vv.v[j] = value;
    const size_t full = n / TILE;
    overlap_prologue();
    for (size_t t = blockIdx.x; t < full; t += gridDim.x) st_stream<T, V>(o + t * TILE + size_t(threadIdx.x) * V, vv);
    if (blockIdx.x == full % gridDim.x)
        for (size_t i = full * TILE + threadIdx.x; i < n; i += THREADS) o[i] = value;
}

// ---- to_vec_with_nodata (src/masked/masked_buffer.rs:137-152): convert + select in one pass ---------
// Thread owns V consecutive cells; their V validity bits sit inside one mask word.
template <class S, class D, int VB, int UNROLL, int THREADS>
__global__ void __launch_bounds__(THREADS) fill_nodata_kernel(const S* __restrict__ a, const uint32_t* __restrict__ m,
                                                              D* __restrict__ o, size_t n, D nodata) {
    constexpr int V0 = VB / cmax<sizeof(S), sizeof(D)>();
    constexpr int V = V0 > 32 ? 32 : V0;
    constexpr size_t TILE = size_t(THREADS) * V * UNROLL;
    constexpr uint32_t VMASK = V == 32 ? 0xFFFFFFFFu : ((1u << V) - 1u);
    const size_t full = n / TILE;
    overlap_prologue();
    for (size_t t = blockIdx.x; t < full; t += gridDim.x) {
        const size_t base = t * TILE + size_t(threadIdx.x) * V;
        Vec<S, V> va[UNROLL];
        uint32_t w[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const size_t c = base + size_t(u) * THREADS * V;
            va[u] = ld_stream<S, V>(a + c);
            w[u] = (__ldg(m + c / 32) >> (c % 32)) & VMASK;
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            Vec<D, V> vo;
#pragma unroll
            for (int j = 0; j < V; ++j) vo.v[j] = ((w[u] >> j) & 1u) ? cast_cell<S, D>(va[u].v[j]) : nodata;
            st_stream<D, V>(o + base + size_t(u) * THREADS * V, vo);
        }
    }
    if (blockIdx.x == full % gridDim.x) {
        for (size_t i = full * TILE + threadIdx.x; i < n; i += THREADS)
            o[i] = ((m[i / 32] >> (i % 32)) & 1u) ? cast_cell<S, D>(a[i]) : nodata;
    }
}

}  // namespace ec
