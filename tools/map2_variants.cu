// Code-generation variants of the two-input streaming kernel for the issue-limited functors (u8 / u16 division, u16
// normalized difference) — development tool. The library kernel sits at a 64-register cap where ptxas' schedule is
// fragile (a signature change moved u8 / u16 at 4096^2 from 34.0 to 37.5 us); this measures alternatives side by side:
//   LIB    the library kernel (map2_kernel<F, 32, 4, 256>)
//   SEQ    same geometry, a compiler barrier after each 256-bit store (one quotient group at a time)
//   R80    same, 3 CTAs / SM (up to 85 registers)
//   U2     UNROLL 2, 256 threads, 4 CTAs / SM;  U2T128: UNROLL 2, 128 threads, 8 CTAs / SM;  U8: UNROLL 8
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -prec-div=true -std=c++17 -Xptxas -v -I erased_cells_b200/csrc tools/map2_variants.cu -o tools/bin/map2_variants
#include <cstdio>
#include <cstdlib>

#include "ec_map.cuh"

using namespace ec;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

template <class F, int UNROLL, int THREADS, int MINCTAS, bool SEQ>
__global__ void __launch_bounds__(THREADS, MINCTAS) variant_kernel(const typename F::A* __restrict__ a, const typename F::B* __restrict__ b,
                                                                   double* __restrict__ o, size_t n, F f) {
    using A = typename F::A; using B = typename F::B;
    constexpr int V = 4;
    constexpr size_t TILE = size_t(THREADS) * V * UNROLL;
    const size_t full = n / TILE;
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
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            Vec<double, V> vo;
#pragma unroll
            for (int j = 0; j < V; ++j) vo.v[j] = f(va[u].v[j], vb[u].v[j]);
            st_stream<double, V>(o + base + size_t(u) * THREADS * V, vo);
            if constexpr (SEQ) asm volatile("" ::: "memory");
        }
    }
    if (blockIdx.x == full % gridDim.x)
        for (size_t i = full * TILE + threadIdx.x; i < n; i += THREADS) o[i] = f(a[i], b[i]);
}

// BAL: a persistent grid of exactly the resident CTAs (MINCTAS per SM), every CTA streaming one contiguous range of
// 1024-cell mini tiles, balanced to +-1 mini tile: all CTAs start together and end together (no partial last wave).
template <class F, int UNROLL, int THREADS, int MINCTAS>
__global__ void __launch_bounds__(THREADS, MINCTAS) balanced_kernel(const typename F::A* __restrict__ a, const typename F::B* __restrict__ b,
                                                                    double* __restrict__ o, size_t n, F f) {
    using A = typename F::A; using B = typename F::B;
    constexpr int V = 4;
    constexpr size_t MT = size_t(THREADS) * V;
    const size_t mts = n / MT;
    const size_t lo = size_t(blockIdx.x) * mts / gridDim.x, hi = size_t(blockIdx.x + 1) * mts / gridDim.x;
    overlap_prologue();
    size_t t = lo;
    for (; t + UNROLL <= hi; t += UNROLL) {
        const size_t base = t * MT + size_t(threadIdx.x) * V;
        Vec<A, V> va[UNROLL];
        Vec<B, V> vb[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            va[u] = ld_stream<A, V>(a + base + size_t(u) * MT);
            vb[u] = ld_stream<B, V>(b + base + size_t(u) * MT);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            Vec<double, V> vo;
#pragma unroll
            for (int j = 0; j < V; ++j) vo.v[j] = f(va[u].v[j], vb[u].v[j]);
            st_stream<double, V>(o + base + size_t(u) * MT, vo);
        }
    }
    for (; t < hi; ++t) {
        const size_t base = t * MT + size_t(threadIdx.x) * V;
        const Vec<A, V> va = ld_stream<A, V>(a + base);
        const Vec<B, V> vb = ld_stream<B, V>(b + base);
        Vec<double, V> vo;
#pragma unroll
        for (int j = 0; j < V; ++j) vo.v[j] = f(va.v[j], vb.v[j]);
        st_stream<double, V>(o + base, vo);
    }
    if (blockIdx.x == gridDim.x - 1)
        for (size_t i = mts * MT + threadIdx.x; i < n; i += THREADS) o[i] = f(a[i], b[i]);
}

static cudaEvent_t e0, e1;
static unsigned g_it = 0;
template <class Launch> static float timed(Launch&& go, int iters) {
    for (int i = 0; i < 3; ++i) go();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(e0));
        for (int i = 0; i < iters; ++i) go();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        best = ms / iters < best ? ms / iters : best;
    }
    return best;
}
template <class F, int UNROLL, int THREADS, int MINCTAS, bool SEQ>
static float run_variant(const typename F::A* a, const typename F::B* b, double* o, size_t n, size_t slots, int iters) {
    constexpr size_t TILE = size_t(THREADS) * 4 * UNROLL;
    const int grid = int(n / TILE ? n / TILE : 1);
    return timed([&] { const size_t s = (g_it++ % slots) * n; variant_kernel<F, UNROLL, THREADS, MINCTAS, SEQ><<<grid, THREADS>>>(a + s, b + s, o + s, n, F{}); }, iters);
}
template <class F, int UNROLL, int MINCTAS>
static float run_balanced(const typename F::A* a, const typename F::B* b, double* o, size_t n, size_t slots, int iters, int sms) {
    const int grid = sms * MINCTAS;
    return timed([&] { const size_t s = (g_it++ % slots) * n; balanced_kernel<F, UNROLL, 256, MINCTAS><<<grid, 256>>>(a + s, b + s, o + s, n, F{}); }, iters);
}
template <class F> static float run_lib(const typename F::A* a, const typename F::B* b, double* o, size_t n, size_t slots, int iters) {
    constexpr size_t TILE = size_t(256) * 4 * 4;
    const int grid = int(n / TILE ? n / TILE : 1);
    return timed([&] { const size_t s = (g_it++ % slots) * n; map2_kernel<F, 32, 4, 256><<<grid, 256>>>(a + s, b + s, o + s, n, F{}, nullptr, nullptr, nullptr, MaskCount{nullptr, nullptr, 0}); }, iters);
}
template <class F> static void sweep(const char* name, const typename F::A* a, const typename F::B* b, double* o, double bpc) {
    for (size_t n : {size_t(1) << 24, size_t(1) << 26, size_t(1) << 28}) {
        const size_t slots = (size_t(1) << 29) / n;  // rotate: 2^29 cells of arena per operand
        const int iters = n >= (size_t(1) << 28) ? 4 : 32;
        const float t[] = {run_lib<F>(a, b, o, n, slots, iters),
                           run_variant<F, 4, 256, 4, true>(a, b, o, n, slots, iters),
                           run_variant<F, 4, 256, 3, false>(a, b, o, n, slots, iters),
                           run_variant<F, 4, 256, 3, true>(a, b, o, n, slots, iters),
                           run_variant<F, 2, 256, 4, false>(a, b, o, n, slots, iters),
                           run_variant<F, 2, 128, 8, false>(a, b, o, n, slots, iters),
                           run_variant<F, 8, 256, 4, false>(a, b, o, n, slots, iters),
                           run_variant<F, 4, 256, 4, false>(a, b, o, n, slots, iters),
                           run_variant<F, 2, 256, 4, true>(a, b, o, n, slots, iters),
                           run_balanced<F, 4, 4>(a, b, o, n, slots, iters, 148), run_balanced<F, 8, 4>(a, b, o, n, slots, iters, 148),
                           run_balanced<F, 2, 4>(a, b, o, n, slots, iters, 148), run_balanced<F, 4, 3>(a, b, o, n, slots, iters, 148)};
        printf("%s n=2^%d us:  LIB %.2f  SEQ %.2f  R80 %.2f  R80SEQ %.2f  U2 %.2f  U2T128 %.2f  U8 %.2f  PLAIN(U4,64r) %.2f  U2SEQ %.2f  BAL_U4 %.2f  BAL_U8 %.2f  BAL_U2 %.2f  BAL_U4_3cta %.2f   (ideal at 6.53 TB/s: %.2f)\n", name,
               63 - __builtin_clzll(n), t[0] * 1e3, t[1] * 1e3, t[2] * 1e3, t[3] * 1e3, t[4] * 1e3, t[5] * 1e3, t[6] * 1e3, t[7] * 1e3, t[8] * 1e3, t[9] * 1e3, t[10] * 1e3, t[11] * 1e3, t[12] * 1e3,
               bpc * n / 6532.5e9 * 1e6);
    }
}
__global__ void fill(uint8_t* p, size_t n, uint64_t seed) {
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) p[i] = uint8_t(splitmix64(seed ^ i) >> 13);
}
int main() {
    const size_t N = size_t(1) << 29;
    uint8_t *a, *b; double* o;
    CK(cudaMalloc(&a, N * 2)); CK(cudaMalloc(&b, N * 2)); CK(cudaMalloc(&o, N * 8));
    fill<<<4096, 256>>>(a, N * 2, 1); fill<<<4096, 256>>>(b, N * 2, 2);
    CK(cudaDeviceSynchronize());
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    sweep<BinaryF<uint8_t, uint16_t, OP_DIV>>("div_u8_u16  ", a, reinterpret_cast<const uint16_t*>(b), o, 11);
    sweep<NormDiffF<uint16_t, uint16_t>>("normdiff_u16", reinterpret_cast<const uint16_t*>(a), reinterpret_cast<const uint16_t*>(b), o, 12);
    sweep<BinaryF<int16_t, int16_t, OP_SUB>>("sub_i16_i16 ", reinterpret_cast<const int16_t*>(a), reinterpret_cast<const int16_t*>(b), o, 12);
    sweep<BinaryScalarT<uint8_t, uint16_t, OP_DIV, OP_MUL>>("div_mul_u8u16", a, reinterpret_cast<const uint16_t*>(b), o, 11);
    return 0;
}
