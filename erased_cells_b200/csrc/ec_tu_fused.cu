// Compile-time specialisations of the fused `(l op1 r) op2 scalar` kernel for the operand pairs band math actually
// uses — both rasters of the same cell type (10 pairs) and the README's u8/u16 — for all 16 (op1, op2). Other pairs
// run the runtime-op functor of ec_tu_binary.cu (same results, ~1.4x the instructions).
#include "ec_internal.hpp"
#include "ec_map.cuh"

#ifndef EC_VB
#define EC_VB 32
#endif
#ifndef EC_UNROLL
#define EC_UNROLL 4
#endif

namespace ec {

template <class L, class R, int OP1, int OP2>
static cudaError_t go(const Launch& Lc, const void* l, const void* r, double s, double* out, size_t n) {
    using F = BinaryScalarT<L, R, OP1, OP2>;
    constexpr int V = EC_VB / cmax<cmax<sizeof(L), sizeof(R)>(), 8>();
    constexpr size_t TILE = size_t(kThreads) * V * EC_UNROLL;
    return launch_k(Lc, map2_kernel<F, EC_VB, EC_UNROLL, kThreads>, grid_for(n, TILE, Lc), kThreads,
                    static_cast<const L*>(l), static_cast<const R*>(r), out, n, F{s}, nullptr, nullptr, nullptr, MaskCount{nullptr, nullptr, 0});
}
template <class L, class R, int OP1>
static cudaError_t by_op2(const Launch& Lc, int op2, const void* l, const void* r, double s, double* out, size_t n) {
    switch (op2) {
        case OP_ADD: return go<L, R, OP1, OP_ADD>(Lc, l, r, s, out, n);
        case OP_SUB: return go<L, R, OP1, OP_SUB>(Lc, l, r, s, out, n);
        case OP_MUL: return go<L, R, OP1, OP_MUL>(Lc, l, r, s, out, n);
        default: return go<L, R, OP1, OP_DIV>(Lc, l, r, s, out, n);
    }
}
template <class L, class R>
static cudaError_t by_ops(const Launch& Lc, int op1, int op2, const void* l, const void* r, double s, double* out, size_t n) {
    switch (op1) {
        case OP_ADD: return by_op2<L, R, OP_ADD>(Lc, op2, l, r, s, out, n);
        case OP_SUB: return by_op2<L, R, OP_SUB>(Lc, op2, l, r, s, out, n);
        case OP_MUL: return by_op2<L, R, OP_MUL>(Lc, op2, l, r, s, out, n);
        default: return by_op2<L, R, OP_DIV>(Lc, op2, l, r, s, out, n);
    }
}

// cudaErrorNotSupported: no specialisation for this operand pair (the caller falls back to the runtime-op kernel)
cudaError_t launch_binary_scalar_static(const Launch& Lc, int op1, int lct, const void* l, int rct, const void* r, int op2, double s,
                                        double* out, size_t n) {
    if (lct == CT_U8 && rct == CT_U16) return by_ops<uint8_t, uint16_t>(Lc, op1, op2, l, r, s, out, n);
    if (lct != rct) return cudaErrorNotSupported;
    switch (lct) {
#define X(id, p) case id: return by_ops<p, p>(Lc, op1, op2, l, r, s, out, n);
        EC_WITH_CT(X)
#undef X
    }
    return cudaErrorNotSupported;
}

// ---- `(a op1 s1) op2 s2` for every cell type and all 16 (op1, op2) -----------------------------------------------
template <class L, int OP1, int OP2>
static cudaError_t go_ss(const Launch& Lc, const void* a, double s1, double s2, double* out, size_t n) {
    using F = ScalarScalarT<L, OP1, OP2>;
    constexpr int V = EC_VB / cmax<sizeof(L), 8>();
    constexpr size_t TILE = size_t(kThreads) * V * EC_UNROLL;
    return launch_k(Lc, map1_kernel<F, EC_VB, EC_UNROLL, kThreads>, grid_for(n, TILE, Lc), kThreads, static_cast<const L*>(a), out, n, F{s1, s2});
}
template <class L, int OP1> static cudaError_t ss_op2(const Launch& Lc, int op2, const void* a, double s1, double s2, double* out, size_t n) {
    switch (op2) {
        case OP_ADD: return go_ss<L, OP1, OP_ADD>(Lc, a, s1, s2, out, n);
        case OP_SUB: return go_ss<L, OP1, OP_SUB>(Lc, a, s1, s2, out, n);
        case OP_MUL: return go_ss<L, OP1, OP_MUL>(Lc, a, s1, s2, out, n);
        default: return go_ss<L, OP1, OP_DIV>(Lc, a, s1, s2, out, n);
    }
}
template <class L> static cudaError_t ss_ops(const Launch& Lc, int op1, int op2, const void* a, double s1, double s2, double* out, size_t n) {
    switch (op1) {
        case OP_ADD: return ss_op2<L, OP_ADD>(Lc, op2, a, s1, s2, out, n);
        case OP_SUB: return ss_op2<L, OP_SUB>(Lc, op2, a, s1, s2, out, n);
        case OP_MUL: return ss_op2<L, OP_MUL>(Lc, op2, a, s1, s2, out, n);
        default: return ss_op2<L, OP_DIV>(Lc, op2, a, s1, s2, out, n);
    }
}
cudaError_t launch_scalar_scalar(const Launch& Lc, int op1, int ct, const void* a, double s1, int op2, double s2, double* out, size_t n) {
    switch (ct) {
#define X(id, p) case id: return ss_ops<p>(Lc, op1, op2, a, s1, s2, out, n);
        EC_WITH_CT(X)
#undef X
    }
    return cudaErrorInvalidValue;
}

}  // namespace ec
