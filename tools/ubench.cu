// Kernel-geometry sweep for the streaming kernels (development tool, not part of the library).
// For representative ops of the five BASELINE configs it times every (VB, UNROLL, THREADS, grid cap)
// combination with CUDA events on buffers far larger than L2 and prints algorithmic GB/s, so the
// library defaults can be chosen from measurements. Also prints what the GPU's raw f64 ops do with
// NaNs (the library rewrites NaN results to the x86 rule regardless).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -prec-div=true -std=c++17 \
//        -I erased_cells_b200/csrc tools/ubench.cu -o tools/ubench
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ec_map.cuh"
#include "ec_mask.cuh"
#include "ec_reduce.cuh"

using namespace ec;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

static int g_sms = 148;
static size_t g_arena_in = size_t(1) << 30, g_arena_out = size_t(2) << 30;  // bytes
static unsigned g_it = 0;
// rotate through the arena so that small inputs are not L2-resident between launches
template <class T> static T* rot(T* base, size_t n, size_t arena_bytes) {
    size_t slots = arena_bytes / (n * sizeof(T));
    if (slots < 1) slots = 1;
    if (slots > 64) slots = 64;
    return base + (g_it % slots) * n;
}
static cudaEvent_t g_e0, g_e1;
static void* g_flush = nullptr;
static size_t g_flush_bytes = 0;
static ReduceScratch g_sc;

template <class Fn> static float time_ms(Fn&& launch, int iters = 16) {
    for (int i = 0; i < 3; ++i) launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f, sum = 0;
    for (int i = 0; i < iters; ++i) {
        CK(cudaEventRecord(g_e0));
        launch();
        CK(cudaEventRecord(g_e1));
        CK(cudaEventSynchronize(g_e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, g_e0, g_e1));
        best = ms < best ? ms : best;
        sum += ms;
    }
    CK(cudaGetLastError());
    printf("%.4f,%.4f,", best, sum / iters);
    return best;
}
static int grid_of(size_t n, size_t tile, int cap) {
    size_t full = n / tile;
    if (full == 0) full = 1;
    if (cap > 0 && full > (size_t)cap) full = cap;
    return (int)full;
}
static const int kCaps[] = {0, 2, 4, 8, 16};  // x SM count; 0 = one tile per CTA

static void report(const char* op, int vb, int unroll, int threads, int capmul, double bytes, float ms) {
    printf("%s,vb=%d,unroll=%d,threads=%d,cap=%d,GBps=%.1f\n", op, vb, unroll, threads, capmul, bytes / (ms * 1e-3) / 1e9);
}

template <class F, int VB, int UNROLL, int THREADS>
static void run_map1(const char* op, const typename F::A* a, typename F::O* o, size_t n, F f, double bpc) {
    constexpr int V = VB / cmax<sizeof(typename F::A), sizeof(typename F::O)>();
    constexpr size_t TILE = size_t(THREADS) * V * UNROLL;
    for (int cm : kCaps) {
        const int grid = grid_of(n, TILE, cm * g_sms);
        float ms = time_ms([&] { ++g_it; map1_kernel<F, VB, UNROLL, THREADS><<<grid, THREADS>>>(rot(a, n, g_arena_in), rot(o, n, g_arena_out), n, f); });
        report(op, VB, UNROLL, THREADS, cm, bpc * n, ms);
    }
}
template <class F, int VB, int UNROLL, int THREADS>
static void run_map2(const char* op, const typename F::A* a, const typename F::B* b, double* o, size_t n, F f, double bpc,
                     const uint32_t* lm = nullptr, const uint32_t* rm = nullptr, uint32_t* om = nullptr) {
    constexpr int V = VB / cmax<cmax<sizeof(typename F::A), sizeof(typename F::B)>(), 8>();
    constexpr size_t TILE = size_t(THREADS) * V * UNROLL;
    for (int cm : kCaps) {
        const int grid = grid_of(n, TILE, cm * g_sms);
        float ms = time_ms([&] {
            ++g_it;
            if (lm) map2_kernel<F, VB, UNROLL, THREADS, true><<<grid, THREADS>>>(rot(a, n, g_arena_in), rot(b, n, g_arena_in), rot(o, n, g_arena_out), n, f, lm, rm, om, MaskCount{nullptr, nullptr, 0});
            else map2_kernel<F, VB, UNROLL, THREADS, false><<<grid, THREADS>>>(rot(a, n, g_arena_in), rot(b, n, g_arena_in), rot(o, n, g_arena_out), n, f, lm, rm, om, MaskCount{nullptr, nullptr, 0});
        });
        report(op, VB, UNROLL, THREADS, cm, bpc * n, ms);
    }
}
template <class T, bool MASKED, int VB, int UNROLL, int THREADS>
static void run_minmax(const char* op, const T* a, const uint32_t* m, size_t n, double bpc) {
    constexpr int V = VB / sizeof(T);
    constexpr size_t TILE = size_t(THREADS) * V * UNROLL;
    const okey_t<T> smin = to_key<T>(std::numeric_limits<T>::max()), smax = to_key<T>(std::numeric_limits<T>::lowest());
    for (int cm : {2, 4, 8, 16, 32}) {
        const int grid = grid_of(n, TILE, cm * g_sms);
        float ms = time_ms([&] { ++g_it; min_max_kernel<T, MASKED, VB, UNROLL, THREADS><<<grid, THREADS>>>(rot(a, n, g_arena_in), m, n, smin, smax, g_sc); });
        report(op, VB, UNROLL, THREADS, cm, bpc * n, ms);
    }
}
template <class U, int VB, int UNROLL, int THREADS>
static void run_maskbuild(const char* op, const U* a, size_t n, uint32_t* out, double bpc) {
    constexpr int V0 = VB / sizeof(U);
    constexpr int V = V0 > 32 ? 32 : V0;
    constexpr size_t TILE = size_t(THREADS) * V * UNROLL;
    for (int cm : kCaps) {
        const int grid = grid_of(n, TILE, cm * g_sms);
        float ms = time_ms([&] { ++g_it; mask_build_kernel<U, false, VB, UNROLL, THREADS><<<grid, THREADS>>>(rot(a, n, g_arena_in), n, U(0x8000), out, MaskCount{nullptr, nullptr, 0}); });
        report(op, VB, UNROLL, THREADS, cm, bpc * n, ms);
    }
}

__global__ void nan_probe(uint64_t* out) {
    const double z = __longlong_as_double(0), inf = __longlong_as_double(0x7FF0000000000000ll);
    const double snan = __longlong_as_double(0x7FF0000000000123ll), qnan = __longlong_as_double(0xFFF8000000000456ull);
    double r[8] = {__ddiv_rn(z, z), __dsub_rn(inf, inf), __dmul_rn(inf, z), __dadd_rn(snan, qnan), __dadd_rn(qnan, snan),
                   __dadd_rn(1.0, snan), __dmul_rn(snan, 2.0), (double)__uint_as_float(0xFF800001u)};
    for (int i = 0; i < 8; ++i) out[i] = (uint64_t)__double_as_longlong(r[i]);
}

template <int VB, int UNROLL, int THREADS> static void sweep(size_t n, void* in0, void* in1, void* outp, uint32_t* m0, uint32_t* m1, uint32_t* m2) {
    double* o = static_cast<double*>(outp);
    run_map1<CastF<double, double>, VB, UNROLL, THREADS>("copy_f64", (const double*)in0, o, n / 2, CastF<double, double>{}, 16);
    run_map2<BinaryF<uint8_t, uint16_t, OP_DIV>, VB, UNROLL, THREADS>("div_u8_u16", (const uint8_t*)in0, (const uint16_t*)in1, o, n, {}, 11);
    run_map2<BinaryF<int16_t, int16_t, OP_SUB>, VB, UNROLL, THREADS>("sub_i16_i16", (const int16_t*)in0, (const int16_t*)in1, o, n, {}, 12);
    run_map2<BinaryF<int16_t, int16_t, OP_SUB>, VB, UNROLL, THREADS>("masked_sub_i16_i16", (const int16_t*)in0, (const int16_t*)in1, o, n, {}, 12.375, m0, m1, m2);
    run_map2<BinaryF<double, double, OP_DIV>, VB, UNROLL, THREADS>("div_f64_f64", (const double*)in0, (const double*)in1, o, n / 2, {}, 24);
    run_map2<NormDiffF<uint16_t, uint16_t>, VB, UNROLL, THREADS>("normdiff_u16", (const uint16_t*)in0, (const uint16_t*)in1, o, n, {}, 12);
    run_map1<ScalarF<double>, VB, UNROLL, THREADS>("mul_f64_scalar", (const double*)in0, o, n / 2, ScalarF<double>{OP_MUL, 0.5}, 16);
    run_map1<CastF<uint8_t, uint16_t>, VB, UNROLL, THREADS>("cast_u8_u16", (const uint8_t*)in0, (uint16_t*)outp, n, {}, 3);
    run_map1<CastF<uint8_t, double>, VB, UNROLL, THREADS>("cast_u8_f64", (const uint8_t*)in0, o, n, {}, 9);
    run_map1<CastF<uint64_t, double>, VB, UNROLL, THREADS>("cast_u64_f64", (const uint64_t*)in0, o, n / 2, {}, 16);
    run_map1<CastF<float, double>, VB, UNROLL, THREADS>("cast_f32_f64", (const float*)in0, o, n, {}, 12);
    run_minmax<float, false, VB, UNROLL, THREADS>("minmax_f32", (const float*)in0, nullptr, n, 4);
    run_minmax<uint8_t, false, VB, UNROLL, THREADS>("minmax_u8", (const uint8_t*)in0, nullptr, n * 4, 1);
    run_minmax<int16_t, false, VB, UNROLL, THREADS>("minmax_i16", (const int16_t*)in0, nullptr, n * 2, 2);
    run_minmax<double, true, VB, UNROLL, THREADS>("masked_minmax_f64", (const double*)in0, m0, n / 2, 8.125);
    run_minmax<double, false, VB, UNROLL, THREADS>("minmax_f64", (const double*)in0, nullptr, n / 2, 8);
    run_maskbuild<uint16_t, VB, UNROLL, THREADS>("from_nodata_i16", (const uint16_t*)in0, n * 2, m2, 2.125);
}

int main(int argc, char** argv) {
    size_t n = size_t(1) << 28;  // cells for <= 4-byte types (f32: 1 GiB); 8-byte ops use n/2
    if (argc > 1) n = strtoull(argv[1], nullptr, 0);
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    g_sms = prop.multiProcessorCount;
    printf("# device %s, %d SMs, L2 %.0f MB, n=%zu, hint=%s\n", prop.name, g_sms, prop.l2CacheSize / 1048576.0, n,
#ifdef EC_HINT_PLAIN
           "plain(.nc / .cs)"
#else
           "no_allocate"
#endif
    );
    CK(cudaEventCreate(&g_e0));
    CK(cudaEventCreate(&g_e1));
    void *in0, *in1, *out;
    const size_t n0 = size_t(1) << 28;
    if (n > n0) { fprintf(stderr, "n too large\n"); return 1; }
    CK(cudaMalloc(&in0, g_arena_in));
    CK(cudaMalloc(&in1, g_arena_in));
    CK(cudaMalloc(&out, g_arena_out));
    uint32_t *m0, *m1, *m2;
    CK(cudaMalloc(&m0, n0 / 8 * 2 + 64));
    CK(cudaMalloc(&m1, n0 / 8 * 2 + 64));
    CK(cudaMalloc(&m2, n0 / 8 * 2 + 64));
    CK(cudaMemset(in0, 0x3C, g_arena_in));   // finite, non-zero patterns for every type
    CK(cudaMemset(in1, 0x41, g_arena_in));
    CK(cudaMemset(m0, 0xA5, n0 / 8 * 2));
    CK(cudaMemset(m1, 0xFF, n0 / 8 * 2));
    void* scratch;
    CK(cudaMalloc(&scratch, (2 * 8192 + 8) * 8));
    CK(cudaMemset(scratch, 0, (2 * 8192 + 8) * 8));
    g_sc.partials = (uint64_t*)scratch;
    g_sc.result = g_sc.partials + 2 * 8192;
    g_sc.ticket = (unsigned int*)(g_sc.result + 4);

    uint64_t* probe;
    CK(cudaMallocManaged(&probe, 64));
    nan_probe<<<1, 1>>>(probe);
    CK(cudaDeviceSynchronize());
    printf("# raw GPU NaN results: 0/0=%016llx inf-inf=%016llx inf*0=%016llx snan+qnan=%016llx qnan+snan=%016llx 1+snan=%016llx snan*2=%016llx cvt(f32 snan)=%016llx\n",
           (unsigned long long)probe[0], (unsigned long long)probe[1], (unsigned long long)probe[2], (unsigned long long)probe[3],
           (unsigned long long)probe[4], (unsigned long long)probe[5], (unsigned long long)probe[6], (unsigned long long)probe[7]);

    {   // reference point: the runtime's own device-to-device copy
        printf("best_ms,avg_ms,");
        float ms = time_ms([&] { ++g_it; CK(cudaMemcpyAsync(rot((char*)out, n * 4, g_arena_out), rot((char*)in0, n * 4, g_arena_in), n * 4, cudaMemcpyDeviceToDevice)); });
        printf("memcpy_d2d,GBps=%.1f\n", 2.0 * n * 4 / (ms * 1e-3) / 1e9);
    }
    sweep<32, 1, 256>(n, in0, in1, out, m0, m1, m2);
    sweep<32, 2, 256>(n, in0, in1, out, m0, m1, m2);
    sweep<32, 4, 256>(n, in0, in1, out, m0, m1, m2);
    sweep<32, 8, 256>(n, in0, in1, out, m0, m1, m2);
    sweep<16, 2, 256>(n, in0, in1, out, m0, m1, m2);
    sweep<16, 4, 256>(n, in0, in1, out, m0, m1, m2);
    sweep<16, 8, 256>(n, in0, in1, out, m0, m1, m2);
    sweep<32, 2, 512>(n, in0, in1, out, m0, m1, m2);
    sweep<32, 4, 512>(n, in0, in1, out, m0, m1, m2);
    sweep<32, 2, 128>(n, in0, in1, out, m0, m1, m2);
    sweep<32, 4, 128>(n, in0, in1, out, m0, m1, m2);
    sweep<16, 4, 512>(n, in0, in1, out, m0, m1, m2);
    return 0;
}
