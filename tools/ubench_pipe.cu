// Round-2 experiment, ready to run (development tool, not part of the library): a PERSISTENT, SOFTWARE-PIPELINED
// variant of map2_kernel against the library's one-tile-per-CTA kernel, for the functors that are issue-limited
// rather than HBM-limited (the f64 division: u8/u16, the NDVI quotient, f32/f32).
//
// Why it might help: with one tile per CTA a thread's life is load -> wait -> 16 quotients -> store -> exit; while
// a CTA computes it has no loads in flight, and at 4 CTAs/SM (64 registers) the SM's bytes in flight sag whenever
// two of the four are in their compute phase. Here a CTA stays resident, and the loads of tile i+1 are issued before
// the quotients of tile i are computed; operands are narrow (u8/u16/f32), so the second register set is 12-32 bytes
// per thread. Grid = resident CTAs exactly (SMs x CTAs/SM), tiles assigned round-robin.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -prec-div=true -std=c++17 \
//        -I erased_cells_b200/csrc tools/ubench_pipe.cu -o tools/bin/ubench_pipe && tools/bin/ubench_pipe
#include <cstdio>
#include <cstdlib>

#include "ec_map.cuh"

using namespace ec;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

template <class F, int VB, int UNROLL, int THREADS, int MIN_CTAS>
__global__ void __launch_bounds__(THREADS, MIN_CTAS) map2_pipe_kernel(const typename F::A* __restrict__ a, const typename F::B* __restrict__ b,
                                                                      typename F::O* __restrict__ o, size_t n, F f) {
    using A = typename F::A; using B = typename F::B; using O = typename F::O;
    constexpr int V = VB / cmax<cmax<sizeof(A), sizeof(B)>(), sizeof(O)>();
    constexpr size_t TILE = size_t(THREADS) * V * UNROLL;
    const size_t full = n / TILE;
    Vec<A, V> va[UNROLL], na[UNROLL];
    Vec<B, V> vb[UNROLL], nb[UNROLL];
    size_t t = blockIdx.x;
    if (t < full) {
        const size_t base = t * TILE + size_t(threadIdx.x) * V;
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            va[u] = ld_stream<A, V>(a + base + size_t(u) * THREADS * V);
            vb[u] = ld_stream<B, V>(b + base + size_t(u) * THREADS * V);
        }
    }
    while (t < full) {
        const size_t tn = t + gridDim.x;
        if (tn < full) {  // next tile's loads go out before this tile's arithmetic
            const size_t nbase = tn * TILE + size_t(threadIdx.x) * V;
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                na[u] = ld_stream<A, V>(a + nbase + size_t(u) * THREADS * V);
                nb[u] = ld_stream<B, V>(b + nbase + size_t(u) * THREADS * V);
            }
        }
        const size_t base = t * TILE + size_t(threadIdx.x) * V;
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            Vec<O, V> vo;
#pragma unroll
            for (int j = 0; j < V; ++j) vo.v[j] = f(va[u].v[j], vb[u].v[j]);
            st_stream<O, V>(o + base + size_t(u) * THREADS * V, vo);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) { va[u] = na[u]; vb[u] = nb[u]; }
        t = tn;
    }
    if (blockIdx.x == full % gridDim.x)
        for (size_t i = full * TILE + threadIdx.x; i < n; i += THREADS) o[i] = f(a[i], b[i]);
}

static cudaEvent_t e0, e1;
template <class K> static float timed(K&& launch, int iters) {
    CK(cudaEventRecord(e0));
    for (int i = 0; i < iters; ++i) launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / iters;
}

template <class F, int UNROLL, int MIN_CTAS>
static void compare(const char* name, const typename F::A* a, const typename F::B* b, double* o, size_t n, double bpc, int sms) {
    constexpr int V = 32 / 8;
    constexpr size_t TILE_LIB = size_t(256) * V * 4, TILE_PIPE = size_t(256) * V * UNROLL;
    const int grid_lib = int(n / TILE_LIB ? n / TILE_LIB : 1);
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, map2_pipe_kernel<F, 32, UNROLL, 256, MIN_CTAS>, 256, 0));
    size_t grid_pipe = size_t(sms) * per_sm;
    if (grid_pipe > n / TILE_PIPE) grid_pipe = n / TILE_PIPE ? n / TILE_PIPE : 1;
    const int iters = n >= (size_t(1) << 28) ? 5 : 40;
    float tl = 0, tp = 0;
    for (int rep = 0; rep < 4; ++rep) {  // alternate; the first repetition warms up
        const float x = timed([&] { map2_kernel<F, 32, 4, 256><<<grid_lib, 256>>>(a, b, o, n, F{}, nullptr, nullptr, nullptr, MaskCount{nullptr, nullptr, 0}); }, iters);
        const float y = timed([&] { map2_pipe_kernel<F, 32, UNROLL, 256, MIN_CTAS><<<int(grid_pipe), 256>>>(a, b, o, n, F{}); }, iters);
        if (rep) { tl += x / 3; tp += y / 3; }
    }
    CK(cudaGetLastError());
    printf("%-16s n=2^%-2d unroll=%d ctas/sm=%d  library %.4f ms (%5.0f GB/s)   pipelined %.4f ms (%5.0f GB/s)   %+.1f %%\n", name, 63 - __builtin_clzll(n),
           UNROLL, per_sm, tl, bpc * n / tl / 1e6, tp, bpc * n / tp / 1e6, (tl / tp - 1) * 100);
}

__global__ void fill16(uint16_t* p, size_t n, uint64_t seed) {
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
        const uint64_t h = splitmix64(seed ^ i);
        p[i] = (h % 1000 == 0) ? 0 : uint16_t(5000 + (h >> 20) % 35001);
    }
}
__global__ void fill8(uint8_t* p, size_t n, uint64_t seed) {
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) p[i] = uint8_t(splitmix64(seed ^ i));
}
__global__ void fillf(float* p, size_t n, uint64_t seed) {
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
        const uint64_t h = splitmix64(seed ^ i);
        p[i] = (h % 1000 == 0) ? 0.0f : float((h >> 11) * 0x1p-53);
    }
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    const size_t N = size_t(1) << 29;
    uint16_t *a16, *b16; uint8_t* c8; float *fa, *fb; double* o;
    CK(cudaMalloc(&a16, N * 2)); CK(cudaMalloc(&b16, N * 2)); CK(cudaMalloc(&c8, N)); CK(cudaMalloc(&fa, N * 4)); CK(cudaMalloc(&fb, N * 4)); CK(cudaMalloc(&o, N * 8));
    fill16<<<4096, 256>>>(a16, N, 1); fill16<<<4096, 256>>>(b16, N, 2); fill8<<<4096, 256>>>(c8, N, 3); fillf<<<4096, 256>>>(fa, N, 4); fillf<<<4096, 256>>>(fb, N, 5);
    CK(cudaDeviceSynchronize());
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    printf("# %s, %d SMs; library = map2_kernel<F,32,4,256>, one tile per CTA; pipelined = persistent grid, next tile's loads ahead of this tile's arithmetic\n", prop.name, sms);
    for (size_t n : {size_t(1) << 24, size_t(1) << 26, size_t(1) << 29}) {
        compare<BinaryF<uint8_t, uint16_t, OP_DIV>, 4, 4>("div_u8_u16", c8, b16, o, n, 11, sms);
        compare<BinaryF<uint8_t, uint16_t, OP_DIV>, 2, 4>("div_u8_u16", c8, b16, o, n, 11, sms);
        compare<BinaryF<uint8_t, uint16_t, OP_DIV>, 4, 3>("div_u8_u16", c8, b16, o, n, 11, sms);
        compare<NormDiffF<uint16_t, uint16_t>, 4, 4>("normdiff_u16", a16, b16, o, n, 12, sms);
        compare<NormDiffF<uint16_t, uint16_t>, 2, 4>("normdiff_u16", a16, b16, o, n, 12, sms);
        compare<NormDiffF<float, float>, 4, 4>("normdiff_f32", fa, fb, o, n, 16, sms);
        compare<NormDiffF<float, float>, 4, 3>("normdiff_f32", fa, fb, o, n, 16, sms);
        compare<BinaryF<float, float, OP_DIV>, 4, 4>("div_f32_f32", fa, fb, o, n, 16, sms);
        compare<BinaryF<int16_t, int16_t, OP_SUB>, 4, 4>("sub_i16_i16", reinterpret_cast<const int16_t*>(a16), reinterpret_cast<const int16_t*>(b16), o, n, 12, sms);
    }
    return 0;
}
