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
// acc = src | acc = acc op src | acc = src op acc | tmp[op] = acc (a STORE keeps its temporary in `op`, src = 7 = no operand)
enum : uint8_t { VM_LOAD = 0, VM_OP = 1, VM_OPR = 2, VM_STORE = 3 };
// operand ids: 0..3 inputs, 4..6 temporaries, 7 none, 8..15 constants
// One 32-bit word per instruction, opcode = kind * 4 + op in bits 0..7, operand id in bits 8..15: kept as plain words
// so the decode stays in the uniform datapath (the program is the same for every thread).
__host__ __device__ constexpr uint32_t vm_word(int kind, int op, int src) { return uint32_t(kind * 4 + op) | (uint32_t(src) << 8); }
struct VmProgram {
    const void* in[kVmInputs];
    double consts[kVmConsts];
    uint32_t code[kVmCode];
    uint32_t ct[kVmInputs];
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

// One arithmetic VM instruction on V cells: the raw IEEE op for all cells, then ONE test whether any result is NaN
// before the (rare) x86 NaN rewrite — a branch per instruction instead of one per cell. `get(j)` names the operand's
// registers directly (no staging copy).
static __device__ __noinline__ double vm_nan_fix(double r, double a, double b) { return r != r ? x86_nan_result(a, b) : r; }
template <int OP, bool REV, int V, class Get> __device__ __forceinline__ void vm_arith(double (&acc)[V], Get get) {
    double r[V];
    bool any_nan = false;
#pragma unroll
    for (int j = 0; j < V; ++j) {
        const double a = REV ? get(j) : acc[j], b = REV ? acc[j] : get(j);
        if constexpr (OP == OP_ADD) r[j] = __dadd_rn(a, b);
        else if constexpr (OP == OP_SUB) r[j] = __dsub_rn(a, b);
        else if constexpr (OP == OP_MUL) r[j] = __dmul_rn(a, b);
        else r[j] = __ddiv_rn(a, b);
        any_nan |= r[j] != r[j];
    }
    if (__builtin_expect(any_nan, 0)) {  // rare: keep the rewrite out of line so the common path stays V ops + V tests
#pragma unroll
        for (int j = 0; j < V; ++j) r[j] = vm_nan_fix(r[j], REV ? get(j) : acc[j], REV ? acc[j] : get(j));
    }
#pragma unroll
    for (int j = 0; j < V; ++j) acc[j] = r[j];
}
template <int V, class Get> __device__ __forceinline__ void vm_apply(uint32_t opcode, double (&acc)[V], Get get) {
    switch (opcode) {
        case VM_LOAD * 4: _Pragma("unroll") for (int j = 0; j < V; ++j) acc[j] = get(j); break;
        case VM_OP * 4 + OP_ADD: vm_arith<OP_ADD, false, V>(acc, get); break;
        case VM_OP * 4 + OP_SUB: vm_arith<OP_SUB, false, V>(acc, get); break;
        case VM_OP * 4 + OP_MUL: vm_arith<OP_MUL, false, V>(acc, get); break;
        case VM_OP * 4 + OP_DIV: vm_arith<OP_DIV, false, V>(acc, get); break;
        case VM_OPR * 4 + OP_ADD: vm_arith<OP_ADD, true, V>(acc, get); break;
        case VM_OPR * 4 + OP_SUB: vm_arith<OP_SUB, true, V>(acc, get); break;
        case VM_OPR * 4 + OP_MUL: vm_arith<OP_MUL, true, V>(acc, get); break;
        default: vm_arith<OP_DIV, true, V>(acc, get); break;
    }
}

template <int V> __device__ __forceinline__ void vm_run(const VmProgram& p, size_t i, double* __restrict__ out) {
    double x[kVmInputs][V], tmp[kVmTemps][V], acc[V];
#pragma unroll
    for (int k = 0; k < kVmInputs; ++k)
        if (k < p.n_in) vm_load<V>(p.in[k], static_cast<int>(p.ct[k]), i, x[k]);
#pragma unroll
    for (int j = 0; j < V; ++j) acc[j] = 0.0;
    for (int pc = 0; pc < p.n_code; ++pc) {
        const uint32_t w = p.code[pc];  // uniform: the program sits in the kernel's constant bank
        const uint32_t opcode = w & 0xFFu, src = (w >> 8) & 0xFFu;
        if (opcode >= VM_STORE * 4) {  // tmp[op] = acc
            const uint32_t t = opcode - VM_STORE * 4;
            if (t == 0) { _Pragma("unroll") for (int j = 0; j < V; ++j) tmp[0][j] = acc[j]; }
            else if (t == 1) { _Pragma("unroll") for (int j = 0; j < V; ++j) tmp[1][j] = acc[j]; }
            else { _Pragma("unroll") for (int j = 0; j < V; ++j) tmp[2][j] = acc[j]; }
            continue;
        }
        switch (src) {  // dense cases; each arm names its operand registers at compile time
            case 0: vm_apply<V>(opcode, acc, [&](int j) { return x[0][j]; }); break;
            case 1: vm_apply<V>(opcode, acc, [&](int j) { return x[1][j]; }); break;
            case 2: vm_apply<V>(opcode, acc, [&](int j) { return x[2][j]; }); break;
            case 3: vm_apply<V>(opcode, acc, [&](int j) { return x[3][j]; }); break;
            case 4: vm_apply<V>(opcode, acc, [&](int j) { return tmp[0][j]; }); break;
            case 5: vm_apply<V>(opcode, acc, [&](int j) { return tmp[1][j]; }); break;
            case 6: vm_apply<V>(opcode, acc, [&](int j) { return tmp[2][j]; }); break;
            default: {
                const double c = p.consts[src & 7u];
                vm_apply<V>(opcode, acc, [&](int) { return c; });
            } break;
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

#ifndef EC_VM_V
#define EC_VM_V 4
#endif
template <int THREADS>
__global__ void __launch_bounds__(THREADS, 512 / THREADS) vm_kernel(const __grid_constant__ VmProgram p, double* __restrict__ out, size_t n) {
    constexpr int V = EC_VM_V;
    constexpr size_t TILE = size_t(THREADS) * V;
    const size_t full = n / TILE;
    for (size_t t = blockIdx.x; t < full; t += gridDim.x) vm_run<V>(p, t * TILE + size_t(threadIdx.x) * V, out);
    if (blockIdx.x == full % gridDim.x)
        for (size_t i = full * TILE + threadIdx.x; i < n; i += THREADS) vm_run<1>(p, i, out);
}

}  // namespace ec
