// Generic fused evaluation of a pending op chain (lazy mode): an accumulator machine over f64
// whose program is uniform across the grid. Shapes with a dedicated kernel ((X-Y)/(X+Y), (X op Y) op s) never get
// here; this is the catch-all for longer band-math expressions, e.g. EVI = 2.5*(nir-red)/(nir+6*red-7.5*blue+1):
// 8 ops, 3 inputs -> one pass at 6+8 bytes per cell instead of ~130 bytes per cell op by op.
// Every op is the same IEEE f64 op with the same x86 NaN rule as the stand-alone kernels, applied in the same order,
// so results are bit-identical to eager evaluation. Interpretation costs instructions, not bytes: measured against
// the unfused chain, not against the roofline.
#pragma once
#include "ec_common.cuh"

namespace ec {

constexpr int kVmInputs = 4, kVmTemps = 3, kVmConsts = 8, kVmCode = 32;
enum : uint8_t { VM_LOAD = 0, VM_OP = 1, VM_OPR = 2, VM_STORE = 3 };  // acc = src | acc = acc op src | acc = src op acc | tmp[src-4] = acc
// operand ids: 0..3 inputs, 4..6 temporaries, 8..15 constants
struct VmInstr {
    uint8_t kind, op, src, pad;
};
struct VmProgram {
    const void* in[kVmInputs];
    double consts[kVmConsts];
    VmInstr code[kVmCode];
    uint8_t ct[kVmInputs];
    int n_in, n_code;
};

template <int V> __device__ __forceinline__ void vm_load(const void* p, int ct, size_t i, double (&o)[V]) {
    switch (ct) {
#define X(id, T)                                                                                  \
    case id: {                                                                                    \
        const Vec<T, V> v = ld_stream<T, V>(static_cast<const T*>(p) + i);                        \
        _Pragma("unroll") for (int j = 0; j < V; ++j) o[j] = as_f64(v.v[j]);                      \
    } break;
        EC_WITH_CT(X)
#undef X
    }
}

template <int V> __device__ __forceinline__ void vm_run(const VmProgram& p, size_t i, double* __restrict__ out) {
    double x[kVmInputs][V], tmp[kVmTemps][V], acc[V], s[V];
#pragma unroll
    for (int k = 0; k < kVmInputs; ++k)
        if (k < p.n_in) vm_load<V>(p.in[k], p.ct[k], i, x[k]);
#pragma unroll
    for (int j = 0; j < V; ++j) acc[j] = 0.0;
    for (int pc = 0; pc < p.n_code; ++pc) {
        const VmInstr in = p.code[pc];
        if (in.kind == VM_STORE) {
            switch (in.src) {
                case 4: _Pragma("unroll") for (int j = 0; j < V; ++j) tmp[0][j] = acc[j]; break;
                case 5: _Pragma("unroll") for (int j = 0; j < V; ++j) tmp[1][j] = acc[j]; break;
                default: _Pragma("unroll") for (int j = 0; j < V; ++j) tmp[2][j] = acc[j]; break;
            }
            continue;
        }
        switch (in.src) {
            case 0: _Pragma("unroll") for (int j = 0; j < V; ++j) s[j] = x[0][j]; break;
            case 1: _Pragma("unroll") for (int j = 0; j < V; ++j) s[j] = x[1][j]; break;
            case 2: _Pragma("unroll") for (int j = 0; j < V; ++j) s[j] = x[2][j]; break;
            case 3: _Pragma("unroll") for (int j = 0; j < V; ++j) s[j] = x[3][j]; break;
            case 4: _Pragma("unroll") for (int j = 0; j < V; ++j) s[j] = tmp[0][j]; break;
            case 5: _Pragma("unroll") for (int j = 0; j < V; ++j) s[j] = tmp[1][j]; break;
            case 6: _Pragma("unroll") for (int j = 0; j < V; ++j) s[j] = tmp[2][j]; break;
            default: {
                const double c = p.consts[in.src & 7];
                _Pragma("unroll") for (int j = 0; j < V; ++j) s[j] = c;
            } break;
        }
        if (in.kind == VM_LOAD) {
#pragma unroll
            for (int j = 0; j < V; ++j) acc[j] = s[j];
        } else if (in.kind == VM_OP) {
#pragma unroll
            for (int j = 0; j < V; ++j) acc[j] = f64_op_rt<true, true>(in.op, acc[j], s[j]);
        } else {
#pragma unroll
            for (int j = 0; j < V; ++j) acc[j] = f64_op_rt<true, true>(in.op, s[j], acc[j]);
        }
    }
    if constexpr (V == 1) {
        out[i] = acc[0];
    } else {
        Vec<double, V> vo;
#pragma unroll
        for (int j = 0; j < V; ++j) vo.v[j] = acc[j];
        st_stream<double, V>(out + i, vo);
    }
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS) vm_kernel(const __grid_constant__ VmProgram p, double* __restrict__ out, size_t n) {
    constexpr int V = 4;
    constexpr size_t TILE = size_t(THREADS) * V;
    const size_t full = n / TILE;
    for (size_t t = blockIdx.x; t < full; t += gridDim.x) vm_run<V>(p, t * TILE + size_t(threadIdx.x) * V, out);
    if (blockIdx.x == full % gridDim.x)
        for (size_t i = full * TILE + threadIdx.x; i < n; i += THREADS) vm_run<1>(p, i, out);
}

}  // namespace ec
