// Experiment (development tool): does TMA / shared-memory staging of the narrow operands beat the direct
// 256-bit LDG/STG path of ec_map.cuh for mixed-width maps?  Persistent CTAs, a STAGES-deep ring of 1-D
// cp.async.bulk (global -> shared, mbarrier complete_tx) loads issued by one elected thread, consumers read
// their cells from shared memory, results go out either as direct 256-bit stores or staged in shared memory
// and written with cp.async.bulk (shared -> global). Same functors, same arithmetic as the library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -prec-div=true -std=c++17 -I erased_cells_b200/csrc tools/ubench_tma.cu -o tools/bin/ubench_tma
#include <cstdio>
#include <cstdlib>

#include "ec_map.cuh"

using namespace ec;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_store(void* gdst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N> __device__ __forceinline__ void bulk_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <class F, int TILE, int STAGES, int THREADS, bool BULK_STORE>
__global__ void __launch_bounds__(THREADS) map2_tma_kernel(const typename F::A* __restrict__ a, const typename F::B* __restrict__ b,
                                                           double* __restrict__ o, size_t n_tiles, F f) {
    using A = typename F::A; using B = typename F::B;
    constexpr int V = 4;  // 4 cells -> one 32-byte f64 store per thread
    constexpr int ITERS = TILE / (THREADS * V);
    extern __shared__ __align__(128) unsigned char smem[];
    A* sa = reinterpret_cast<A*>(smem);
    B* sb = reinterpret_cast<B*>(smem + size_t(STAGES) * TILE * sizeof(A));
    double* so = reinterpret_cast<double*>(smem + size_t(STAGES) * TILE * (sizeof(A) + sizeof(B)));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + size_t(STAGES) * TILE * (sizeof(A) + sizeof(B)) + (BULK_STORE ? 2 * TILE * sizeof(double) : 0));
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const size_t my_tiles = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    auto issue = [&](size_t k) {  // elected thread: load tile k of this CTA into stage k % STAGES
        const int s = int(k % STAGES);
        const size_t t = blockIdx.x + k * gridDim.x;
        mbar_expect_tx(&full[s], TILE * (sizeof(A) + sizeof(B)));
        bulk_load(sa + size_t(s) * TILE, a + t * TILE, TILE * sizeof(A), &full[s]);
        bulk_load(sb + size_t(s) * TILE, b + t * TILE, TILE * sizeof(B), &full[s]);
    };
    if (tid == 0)
        for (size_t k = 0; k < STAGES && k < my_tiles; ++k) issue(k);
    for (size_t k = 0; k < my_tiles; ++k) {
        const int s = int(k % STAGES);
        const size_t t = blockIdx.x + k * gridDim.x;
        mbar_wait(&full[s], uint32_t((k / STAGES) & 1));
        const A* ta = sa + size_t(s) * TILE;
        const B* tb = sb + size_t(s) * TILE;
        if constexpr (BULK_STORE) {
            double* to = so + size_t(k & 1) * TILE;
            if (k >= 2) {  // the bulk store that last read this output stage must have drained
                if (tid == 0) bulk_store_wait_read<1>();
                __syncthreads();
            }
#pragma unroll
            for (int u = 0; u < ITERS; ++u) {
                const int c = (u * THREADS + tid) * V;
                const Vec<A, V> va = *reinterpret_cast<const Vec<A, V>*>(ta + c);
                const Vec<B, V> vb = *reinterpret_cast<const Vec<B, V>*>(tb + c);
                Vec<double, V> vo;
#pragma unroll
                for (int j = 0; j < V; ++j) vo.v[j] = f(va.v[j], vb.v[j]);
                *reinterpret_cast<Vec<double, V>*>(to + c) = vo;
            }
            fence_async_smem();
            __syncthreads();
            if (tid == 0) {
                bulk_store(o + t * TILE, to, TILE * sizeof(double));
                if (k + STAGES < my_tiles) issue(k + STAGES);
            }
        } else {
#pragma unroll
            for (int u = 0; u < ITERS; ++u) {
                const int c = (u * THREADS + tid) * V;
                const Vec<A, V> va = *reinterpret_cast<const Vec<A, V>*>(ta + c);
                const Vec<B, V> vb = *reinterpret_cast<const Vec<B, V>*>(tb + c);
                Vec<double, V> vo;
#pragma unroll
                for (int j = 0; j < V; ++j) vo.v[j] = f(va.v[j], vb.v[j]);
                st_stream<double, V>(o + t * TILE + c, vo);
            }
            __syncthreads();  // everyone is done reading stage s
            if (tid == 0 && k + STAGES < my_tiles) issue(k + STAGES);
        }
    }
    if constexpr (BULK_STORE) {
        if (tid == 0) bulk_store_wait_read<0>();
    }
}

static cudaEvent_t g_e0, g_e1;
template <class Fn> static float time_ms(Fn&& launch, int iters = 16) {
    for (int i = 0; i < 3; ++i) launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int i = 0; i < iters; ++i) {
        CK(cudaEventRecord(g_e0));
        launch();
        CK(cudaEventRecord(g_e1));
        CK(cudaEventSynchronize(g_e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, g_e0, g_e1));
        best = ms < best ? ms : best;
    }
    CK(cudaGetLastError());
    return best;
}

__global__ void diff_kernel(const double* x, const double* y, size_t n, unsigned long long* bad) {
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x)
        if (__double_as_longlong(x[i]) != __double_as_longlong(y[i])) atomicAdd(bad, 1ull);
}

template <class F, int TILE, int STAGES, bool BULK>
static void run(const char* op, const typename F::A* a, const typename F::B* b, double* o, double* ref, size_t n, double bpc, int sms,
                unsigned long long* bad) {
    using A = typename F::A; using B = typename F::B;
    constexpr int THREADS = 256;
    const size_t smem = size_t(STAGES) * TILE * (sizeof(A) + sizeof(B)) + (BULK ? 2 * TILE * sizeof(double) : 0) + STAGES * 8 + 128;
    auto kern = map2_tma_kernel<F, TILE, STAGES, THREADS, BULK>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const size_t n_tiles = n / TILE;
    for (int cm : {1, 2, 3, 4, 6, 8}) {
        int occ = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, smem));
        if (cm > occ) continue;
        const int grid = cm * sms;
        float ms = time_ms([&] { kern<<<grid, THREADS, smem>>>(a, b, o, n_tiles, F{}); });
        *bad = 0;
        diff_kernel<<<1024, 256>>>(o, ref, n_tiles * TILE, bad);
        CK(cudaDeviceSynchronize());
        printf("%s,tile=%d,stages=%d,bulk_store=%d,ctas_per_sm=%d(occ %d),smem=%zu,ms=%.4f,GBps=%.1f,mismatches=%llu\n", op, TILE, STAGES, (int)BULK, cm, occ,
               smem, ms, bpc * n_tiles * TILE / (ms * 1e-3) / 1e9, *bad);
    }
}

template <class F> static void all(const char* op, const typename F::A* a, const typename F::B* b, double* o, double* ref, size_t n, double bpc, int sms,
                                   unsigned long long* bad) {
    constexpr size_t TILE0 = size_t(256) * 4 * 4;
    const int grid = int(n / TILE0);
    float ms = time_ms([&] { map2_kernel<F, 32, 4, 256><<<grid, 256>>>(a, b, ref, n, F{}, nullptr, nullptr, nullptr, MaskCount{nullptr, nullptr, 0}); });
    printf("%s,direct LDG.256/STG.256 (library kernel),ms=%.4f,GBps=%.1f\n", op, ms, bpc * n / (ms * 1e-3) / 1e9);
    run<F, 4096, 4, false>(op, a, b, o, ref, n, bpc, sms, bad);
    run<F, 4096, 8, false>(op, a, b, o, ref, n, bpc, sms, bad);
    run<F, 8192, 4, false>(op, a, b, o, ref, n, bpc, sms, bad);
    run<F, 2048, 8, false>(op, a, b, o, ref, n, bpc, sms, bad);
    run<F, 4096, 4, true>(op, a, b, o, ref, n, bpc, sms, bad);
    run<F, 2048, 4, true>(op, a, b, o, ref, n, bpc, sms, bad);
    run<F, 2048, 8, true>(op, a, b, o, ref, n, bpc, sms, bad);
}

int main(int argc, char** argv) {
    size_t n = size_t(1) << 28;
    if (argc > 1) n = strtoull(argv[1], nullptr, 0);
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    CK(cudaEventCreate(&g_e0));
    CK(cudaEventCreate(&g_e1));
    void *in0, *in1;
    double *out, *ref;
    CK(cudaMalloc(&in0, n * 2));
    CK(cudaMalloc(&in1, n * 2));
    CK(cudaMalloc(&out, n * 8));
    CK(cudaMalloc(&ref, n * 8));
    CK(cudaMemset(in0, 0x3C, n * 2));
    CK(cudaMemset(in1, 0x41, n * 2));
    unsigned long long* bad;
    CK(cudaMallocManaged(&bad, 8));
    printf("# %s, %d SMs, n=%zu cells\n", prop.name, prop.multiProcessorCount, n);
    all<BinaryF<uint8_t, uint16_t, OP_DIV>>("div_u8_u16", (const uint8_t*)in0, (const uint16_t*)in1, out, ref, n, 11, prop.multiProcessorCount, bad);
    all<BinaryF<int16_t, int16_t, OP_SUB>>("sub_i16_i16", (const int16_t*)in0, (const int16_t*)in1, out, ref, n, 12, prop.multiProcessorCount, bad);
    all<NormDiffF<uint16_t, uint16_t>>("normdiff_u16", (const uint16_t*)in0, (const uint16_t*)in1, out, ref, n, 12, prop.multiProcessorCount, bad);
    return 0;
}
