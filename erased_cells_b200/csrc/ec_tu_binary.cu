// Two-input map kernels for one left-hand cell type (compiled once per -DEC_LCT=0..9 so the 10x10
// lattice builds in parallel): binary op x4, normalized difference, binary-then-scalar.
#include "ec_internal.hpp"
#include "ec_map.cuh"

#ifndef EC_LCT
#error "compile with -DEC_LCT=<cell type 0..9>"
#endif
#ifndef EC_VB
#define EC_VB 32
#endif
#ifndef EC_UNROLL
#define EC_UNROLL 4
#endif

namespace ec {

using LT = type_of<EC_LCT>::type;

// Loads in flight per operand. 4 everywhere, except where ptxas cannot fit the schedule into the 64 registers of the
// 4-CTAs-per-SM budget: the division of 8-bit by 16-bit cells spills at 4 and not at 8 (ptxas -v), and the spill-free
// form measured 11-14 % faster at 2^24 .. 2^28 cells (tools/map2_variants.cu, profiles/r02_map2_variants.txt).
template <class F> struct map2_unroll { static constexpr int value = EC_UNROLL; };
template <class L, class R> struct map2_unroll<BinaryF<L, R, OP_DIV>> {
    static constexpr int value = (sizeof(L) == 1 && sizeof(R) == 2 && !is_fp<L> && !is_fp<R>) ? 8 : EC_UNROLL;
};
template <class F, bool MASKED = false>
static cudaError_t go2(const Launch& Lc, const typename F::A* l, const typename F::B* r, double* out, size_t n, F f,
                       const uint32_t* lm, const uint32_t* rm, uint32_t* om, const MaskCount& mc) {
    constexpr int V = EC_VB / cmax<cmax<sizeof(typename F::A), sizeof(typename F::B)>(), sizeof(double)>();
    constexpr int U = MASKED ? EC_UNROLL : map2_unroll<F>::value;
    constexpr size_t TILE = size_t(kThreads) * V * U;
    return launch_k(Lc, map2_kernel<F, EC_VB, U, kThreads, MASKED>, grid_for(n, TILE, Lc), kThreads, l, r, out, n, f, lm, rm, om, mc);
}
static const MaskCount kNoCount{nullptr, nullptr, 0};

template <class R>
static cudaError_t binary_r(const Launch& Lc, int op, const LT* l, const R* r, double* out, size_t n,
                            const uint32_t* lm, const uint32_t* rm, uint32_t* om, const MaskCount& mc) {
    if (lm != nullptr) {  // MaskedCellBuffer op: the flavour that carries the mask AND and its count along
        switch (op) {
            case OP_ADD: return go2<BinaryF<LT, R, OP_ADD>, true>(Lc, l, r, out, n, {}, lm, rm, om, mc);
            case OP_SUB: return go2<BinaryF<LT, R, OP_SUB>, true>(Lc, l, r, out, n, {}, lm, rm, om, mc);
            case OP_MUL: return go2<BinaryF<LT, R, OP_MUL>, true>(Lc, l, r, out, n, {}, lm, rm, om, mc);
            case OP_DIV: return go2<BinaryF<LT, R, OP_DIV>, true>(Lc, l, r, out, n, {}, lm, rm, om, mc);
        }
        return cudaErrorInvalidValue;
    }
    switch (op) {
        case OP_ADD: return go2(Lc, l, r, out, n, BinaryF<LT, R, OP_ADD>{}, lm, rm, om, mc);
        case OP_SUB: return go2(Lc, l, r, out, n, BinaryF<LT, R, OP_SUB>{}, lm, rm, om, mc);
        case OP_MUL: return go2(Lc, l, r, out, n, BinaryF<LT, R, OP_MUL>{}, lm, rm, om, mc);
        case OP_DIV: return go2(Lc, l, r, out, n, BinaryF<LT, R, OP_DIV>{}, lm, rm, om, mc);
    }
    return cudaErrorInvalidValue;
}

#define EC_CAT_(a, b) a##b
#define EC_CAT(a, b) EC_CAT_(a, b)

cudaError_t EC_CAT(launch_binary_l, EC_LCT)(const Launch& Lc, int op, const void* l, int rct, const void* r, double* out,
                                            size_t n, const uint32_t* lm, const uint32_t* rm, uint32_t* om, const MaskCount& mc) {
    switch (rct) {
#define X(id, p) case id: return binary_r<p>(Lc, op, static_cast<const LT*>(l), static_cast<const p*>(r), out, n, lm, rm, om, mc);
        EC_WITH_CT(X)
#undef X
    }
    return cudaErrorInvalidValue;
}

cudaError_t EC_CAT(launch_normdiff_l, EC_LCT)(const Launch& Lc, const void* l, int rct, const void* r, double* out, size_t n) {
    switch (rct) {
#define X(id, p) case id: return go2(Lc, static_cast<const LT*>(l), static_cast<const p*>(r), out, n, NormDiffF<LT, p>{}, nullptr, nullptr, nullptr, kNoCount);
        EC_WITH_CT(X)
#undef X
    }
    return cudaErrorInvalidValue;
}

cudaError_t EC_CAT(launch_binary_scalar_l, EC_LCT)(const Launch& Lc, int op1, const void* l, int rct, const void* r, int op2,
                                                   double s, double* out, size_t n) {
    switch (rct) {
#define X(id, p) case id: return go2(Lc, static_cast<const LT*>(l), static_cast<const p*>(r), out, n, BinaryScalarF<LT, p>{op1, op2, s}, nullptr, nullptr, nullptr, kNoCount);
        EC_WITH_CT(X)
#undef X
    }
    return cudaErrorInvalidValue;
}

}  // namespace ec
