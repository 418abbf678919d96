// A/B of the guard-free integer-operand division (div_int_operands, ec_common.cuh) against div.rn.f64's generic
// expansion, same box, same buffers, alternating launches (development tool, not part of the library).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -prec-div=true -std=c++17 -I erased_cells_b200/csrc tools/div_ab.cu -o tools/bin/div_ab
#include <cstdio>
#include <cstdlib>

#include "ec_map.cuh"

using namespace ec;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

template <class L, class R> struct NormDiffGeneric {
    using A = L; using B = R; using O = double;
    __device__ __forceinline__ double operator()(L a, R b) const {
        const double x = as_f64(a), y = as_f64(b);
        double r = __ddiv_rn(__dsub_rn(x, y), __dadd_rn(x, y));
        if (r != r) r = x86_nan_result(__dsub_rn(x, y), __dadd_rn(x, y));
        return r;
    }
};
template <class L, class R> struct DivGeneric {
    using A = L; using B = R; using O = double;
    __device__ __forceinline__ double operator()(L a, R b) const {
        double r = __ddiv_rn(as_f64(a), as_f64(b));
        if (r != r) r = __longlong_as_double(static_cast<long long>(0xFFF8000000000000ull));
        return r;
    }
};

template <class L, class R> struct DivGenericFp {
    using A = L; using B = R; using O = double;
    __device__ __forceinline__ double operator()(L a, R b) const {
        const double x = as_f64(a), y = as_f64(b);
        double r = __ddiv_rn(x, y);
        if (r != r) r = x86_nan_result(x, y);
        return r;
    }
};

template <class F> static float run(const typename F::A* a, const typename F::B* b, double* o, size_t n, int iters, cudaEvent_t e0, cudaEvent_t e1) {
    constexpr int V = 32 / 8;
    constexpr size_t TILE = size_t(256) * V * 4;
    const int grid = int(n / TILE ? n / TILE : 1);
    CK(cudaEventRecord(e0));
    for (int i = 0; i < iters; ++i) map2_kernel<F, 32, 4, 256><<<grid, 256>>>(a, b, o, n, F{}, nullptr, nullptr, nullptr, MaskCount{nullptr, nullptr, 0});
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / iters;
}

__global__ void fill16(uint16_t* p, size_t n, uint64_t seed, uint32_t lo, uint32_t span) {
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x)
        p[i] = uint16_t(lo + splitmix64(seed ^ i) % span);
}
__global__ void fillf(float* p, size_t n, uint64_t seed) {  // reflectance-like: (0, 1], a few exact zeros
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
        const uint64_t h = splitmix64(seed ^ i);
        p[i] = (h % 1000 == 0) ? 0.0f : float((h >> 11) * 0x1p-53);
    }
}
__global__ void fill8(uint8_t* p, size_t n, uint64_t seed) {
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) p[i] = uint8_t(splitmix64(seed ^ i));
}

int main() {
    const size_t N = size_t(1) << 30;
    uint16_t *a, *b; uint8_t* c; double* o;
    CK(cudaMalloc(&a, N * 2)); CK(cudaMalloc(&b, N * 2)); CK(cudaMalloc(&c, N)); CK(cudaMalloc(&o, N * 8));
    fill16<<<4096, 256>>>(a, N, 1, 5000, 35001); fill16<<<4096, 256>>>(b, N, 2, 5000, 35001); fill8<<<4096, 256>>>(c, N, 3);
    CK(cudaDeviceSynchronize());
    float *fa = reinterpret_cast<float*>(a), *fb = reinterpret_cast<float*>(b);  // the same arenas, 2^29 f32 cells each
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (size_t n : {size_t(1) << 24, size_t(1) << 26, size_t(1) << 28, size_t(1) << 30}) {
        const int iters = n >= (size_t(1) << 28) ? 5 : 40;
        float t[4] = {0, 0, 0, 0};
        for (int rep = 0; rep < 4; ++rep) {  // alternate the variants; first repetition is the warm-up
            const float x0 = run<NormDiffF<uint16_t, uint16_t>>(a, b, o, n, iters, e0, e1);
            const float x1 = run<NormDiffGeneric<uint16_t, uint16_t>>(a, b, o, n, iters, e0, e1);
            const float x2 = run<BinaryF<uint8_t, uint16_t, OP_DIV>>(c, b, o, n, iters, e0, e1);
            const float x3 = run<DivGeneric<uint8_t, uint16_t>>(c, b, o, n, iters, e0, e1);
            if (rep) { t[0] += x0 / 3; t[1] += x1 / 3; t[2] += x2 / 3; t[3] += x3 / 3; }
        }
        printf("n=2^%d  normdiff_u16: guard-free %.4f ms (%.0f GB/s)  generic %.4f ms (%.0f GB/s)   div_u8_u16: guard-free %.4f ms (%.0f GB/s)  generic %.4f ms (%.0f GB/s)\n",
               63 - __builtin_clzll(n), t[0], 12.0 * n / t[0] / 1e6, t[1], 12.0 * n / t[1] / 1e6, t[2], 11.0 * n / t[2] / 1e6, t[3], 11.0 * n / t[3] / 1e6);
    }
    fillf<<<4096, 256>>>(fa, N / 2, 4); fillf<<<4096, 256>>>(fb, N / 2, 5);
    CK(cudaDeviceSynchronize());
    for (size_t n : {size_t(1) << 24, size_t(1) << 26, size_t(1) << 28, size_t(1) << 29}) {
        const int iters = n >= (size_t(1) << 28) ? 5 : 40;
        float t[4] = {0, 0, 0, 0};
        for (int rep = 0; rep < 4; ++rep) {
            const float x0 = run<NormDiffF<float, float>>(fa, fb, o, n, iters, e0, e1);
            const float x1 = run<NormDiffGeneric<float, float>>(fa, fb, o, n, iters, e0, e1);
            const float x2 = run<BinaryF<float, float, OP_DIV>>(fa, fb, o, n, iters, e0, e1);
            const float x3 = run<DivGenericFp<float, float>>(fa, fb, o, n, iters, e0, e1);
            if (rep) { t[0] += x0 / 3; t[1] += x1 / 3; t[2] += x2 / 3; t[3] += x3 / 3; }
        }
        printf("n=2^%d  normdiff_f32: guard-free %.4f ms (%.0f GB/s)  generic %.4f ms (%.0f GB/s)   div_f32_f32: guard-free %.4f ms (%.0f GB/s)  generic %.4f ms (%.0f GB/s)\n",
               63 - __builtin_clzll(n), t[0], 16.0 * n / t[0] / 1e6, t[1], 16.0 * n / t[1] / 1e6, t[2], 16.0 * n / t[2] / 1e6, t[3], 16.0 * n / t[3] / 1e6);
    }
    return 0;
}
