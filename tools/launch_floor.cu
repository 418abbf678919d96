// What does the platform charge for "launch a kernel and see its one-word answer on the host"? The floor under every small
// reduction (min_max / counts of a small raster, the per-strip part of a sharded call).
//   raw        <<<1, 32>>> kernel that writes a tag into mapped pinned memory, host polls          (launch + pickup + publish)
//   raw 592    same with a 592 x 256 grid whose last CTA (atomic ticket) publishes                  (+ grid fan-out / ticket)
//   graph      the 1 x 32 kernel as a one-node CUDA graph, parameters updated before each launch
//   library    ec_buf_min_max on a 4096-cell u8 buffer and on 2^24 cells                            (the product path)
// Prints best / median microseconds of 2000 calls each.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -Iinclude tools/launch_floor.cu -o tools/bin/launch_floor -Lerased_cells_b200/lib -lerased_cells_b200 -Xlinker -rpath,$PWD/erased_cells_b200/lib
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>

#include "erased_cells_b200.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); std::exit(1); } } while (0)
#define EK(x) do { ec_status s_ = (x); if (s_ != EC_OK) { std::fprintf(stderr, "%s: %s\n", #x, ec_last_error()); std::exit(1); } } while (0)

__global__ void publish1(volatile unsigned long long* host, unsigned long long tag) {
    if (threadIdx.x == 0) *host = tag;
}
__global__ void publish_grid(volatile unsigned long long* host, unsigned long long tag, unsigned int* ticket) {
    __shared__ bool last;
    if (threadIdx.x == 0) {
        __threadfence();
        last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && threadIdx.x == 0) { *ticket = 0; *host = tag; }
}
static double now_us() {
    return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
template <class F> static void report(const char* what, F call, int n = 2000) {
    std::vector<double> t(n);
    for (int i = 0; i < 200; ++i) call(i);
    for (int i = 0; i < n; ++i) {
        const double t0 = now_us();
        call(1000 + i);
        t[i] = now_us() - t0;
    }
    std::sort(t.begin(), t.end());
    std::printf("%-44s best %6.2f us   median %6.2f us   p90 %6.2f us\n", what, t[0], t[n / 2], t[n * 9 / 10]);
}
int main() {
    EK(ec_init(0));
    cudaStream_t st;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    volatile unsigned long long* host;
    unsigned long long* dev_view;
    CK(cudaHostAlloc((void**)&host, 64, cudaHostAllocMapped));
    CK(cudaHostGetDevicePointer((void**)&dev_view, (void*)host, 0));
    unsigned int* ticket;
    CK(cudaMalloc(&ticket, 4));
    CK(cudaMemset(ticket, 0, 4));
    *host = 0;
    report("raw <<<1,32>>> + poll", [&](int i) {
        const unsigned long long tag = 0x100000000ull + i;
        publish1<<<1, 32, 0, st>>>(dev_view, tag);
        while (*host != tag) {}
    });
    report("raw <<<592,256>>> ticket + poll", [&](int i) {
        const unsigned long long tag = 0x200000000ull + i;
        publish_grid<<<592, 256, 0, st>>>(dev_view, tag, ticket);
        while (*host != tag) {}
    });
    {   // one-node graph, parameters replaced before every launch
        cudaGraph_t g;
        cudaGraphNode_t node;
        CK(cudaGraphCreate(&g, 0));
        unsigned long long tag = 0;
        void* args[2] = {&dev_view, &tag};
        cudaKernelNodeParams p = {};
        p.func = (void*)publish1;
        p.gridDim = dim3(1); p.blockDim = dim3(32); p.kernelParams = args;
        CK(cudaGraphAddKernelNode(&node, g, nullptr, 0, &p));
        cudaGraphExec_t ex;
        CK(cudaGraphInstantiate(&ex, g, 0));
        report("graph (1 node, params updated) + poll", [&](int i) {
            tag = 0x300000000ull + i;
            CK(cudaGraphExecKernelNodeSetParams(ex, node, &p));
            CK(cudaGraphLaunch(ex, st));
            while (*host != tag) {}
        });
        report("graph (1 node, same params) launch + sync", [&](int) {
            CK(cudaGraphLaunch(ex, st));
            CK(cudaStreamSynchronize(st));
        });
    }
    report("raw <<<1,32>>> + cudaStreamSynchronize", [&](int i) {
        publish1<<<1, 32, 0, st>>>(dev_view, 0x400000000ull + i);
        CK(cudaStreamSynchronize(st));
    });
    ec_set_min_max_cache(0);
    for (size_t n : {size_t(4096), size_t(1) << 24}) {
        std::vector<unsigned char> h(n, 7);
        ec_buf* b;
        EK(ec_buf_from_host(EC_UINT8, h.data(), n, &b));
        EK(ec_synchronize());
        ec_value mn, mx;
        char what[64];
        std::snprintf(what, sizeof what, "library: ec_buf_min_max, u8 x %zu", n);
        report(what, [&](int) { EK(ec_buf_min_max(b, nullptr, &mn, &mx)); });
        ec_buf_free(b);
    }
    {
        ec_mask* m;
        EK(ec_mask_fill(size_t(1) << 24, 1, &m));
        ec_mask* inv;
        EK(ec_mask_not(m, &inv));
        size_t d, nd;
        report("library: ec_mask_counts (counted by its producer)", [&](int) { EK(ec_mask_counts(inv, &d, &nd)); });
        ec_mask_free(m); ec_mask_free(inv);
    }
    return 0;
}
