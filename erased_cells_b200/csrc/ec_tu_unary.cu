// One-input map kernels: buffer (op) scalar, neg, the 31 legal casts, fill, convert+NoData fill.
#include "ec_internal.hpp"
#include "ec_map.cuh"

#ifndef EC_VB
#define EC_VB 32
#endif
#ifndef EC_UNROLL
#define EC_UNROLL 4
#endif

namespace ec {

template <class F>
static cudaError_t go1(const Launch& Lc, const typename F::A* a, typename F::O* out, size_t n, F f) {
    constexpr int V = EC_VB / cmax<sizeof(typename F::A), sizeof(typename F::O)>();
    constexpr size_t TILE = size_t(kThreads) * V * EC_UNROLL;
    return launch_k(Lc, map1_kernel<F, EC_VB, EC_UNROLL, kThreads>, grid_for(n, TILE, Lc), kThreads, a, out, n, f);
}

cudaError_t launch_scalar(const Launch& Lc, int op, int lct, const void* l, double s, double* out, size_t n) {
    switch (lct) {
#define X(id, p) case id: return go1(Lc, static_cast<const p*>(l), out, n, ScalarF<p>{op, s});
        EC_WITH_CT(X)
#undef X
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_neg(const Launch& Lc, int ct, const void* a, void* out, size_t n) {
    switch (ct) {
#define X(id, p) case id: return go1(Lc, static_cast<const p*>(a), static_cast<typename neg_out<p>::type*>(out), n, NegF<p>{});
        EC_WITH_CT(X)
#undef X
    }
    return cudaErrorInvalidValue;
}

// the legal widenings of src/ctype.rs:129-131; identity pairs are D2D copies
#define EC_LEGAL_CASTS(X)                                                                                   \
    X(uint8_t, uint16_t) X(uint8_t, uint32_t) X(uint8_t, uint64_t) X(uint8_t, int16_t) X(uint8_t, int32_t)   \
    X(uint8_t, int64_t) X(uint8_t, float) X(uint8_t, double)                                                 \
    X(uint16_t, uint32_t) X(uint16_t, uint64_t) X(uint16_t, int32_t) X(uint16_t, int64_t) X(uint16_t, float) \
    X(uint16_t, double)                                                                                      \
    X(uint32_t, uint64_t) X(uint32_t, int64_t) X(uint32_t, double)                                           \
    X(uint64_t, double)                                                                                      \
    X(int8_t, int16_t) X(int8_t, int32_t) X(int8_t, int64_t) X(int8_t, float) X(int8_t, double)              \
    X(int16_t, int32_t) X(int16_t, int64_t) X(int16_t, float) X(int16_t, double)                             \
    X(int32_t, int64_t) X(int32_t, double)                                                                   \
    X(int64_t, double)                                                                                       \
    X(float, double)

cudaError_t launch_convert(const Launch& Lc, int sct, const void* a, int dct, void* out, size_t n) {
#define X(S, D) \
    if (sct == ct_of<S>::value && dct == ct_of<D>::value) return go1(Lc, static_cast<const S*>(a), static_cast<D*>(out), n, CastF<S, D>{});
    EC_LEGAL_CASTS(X)
#undef X
    return cudaErrorInvalidValue;
}

// Clone / identity convert (src/buffer.rs:50, :151-153): the same streaming kernel with an identity
// functor on the cell's unsigned carrier (measured faster than the runtime's D2D memcpy on B200).
cudaError_t launch_copy(const Launch& Lc, int cell_bytes, const void* a, void* out, size_t n) {
    switch (cell_bytes) {
        case 1: return go1(Lc, static_cast<const uint8_t*>(a), static_cast<uint8_t*>(out), n, CastF<uint8_t, uint8_t>{});
        case 2: return go1(Lc, static_cast<const uint16_t*>(a), static_cast<uint16_t*>(out), n, CastF<uint16_t, uint16_t>{});
        case 4: return go1(Lc, static_cast<const uint32_t*>(a), static_cast<uint32_t*>(out), n, CastF<uint32_t, uint32_t>{});
        default: return go1(Lc, static_cast<const uint64_t*>(a), static_cast<uint64_t*>(out), n, CastF<uint64_t, uint64_t>{});
    }
}

template <class U> static cudaError_t fill_u(const Launch& Lc, void* out, size_t n, uint64_t bits) {
    constexpr size_t TILE = size_t(kThreads) * (EC_VB / sizeof(U));
    return launch_k(Lc, fill_kernel<U, EC_VB, kThreads>, grid_for(n, TILE, Lc), kThreads, static_cast<U*>(out), n, static_cast<U>(bits));
}
cudaError_t launch_fill(const Launch& Lc, int ct, void* out, size_t n, uint64_t bits) {
    static const int sz[CT_COUNT] = {1, 2, 4, 8, 1, 2, 4, 8, 4, 8};
    switch (sz[ct]) {
        case 1: return fill_u<uint8_t>(Lc, out, n, bits);
        case 2: return fill_u<uint16_t>(Lc, out, n, bits);
        case 4: return fill_u<uint32_t>(Lc, out, n, bits);
        default: return fill_u<uint64_t>(Lc, out, n, bits);
    }
}

template <class S, class D>
static cudaError_t fill_nd(const Launch& Lc, const void* a, const uint32_t* m, void* out, size_t n, uint64_t nd_bits) {
    constexpr int V0 = EC_VB / cmax<sizeof(S), sizeof(D)>();
    constexpr int V = V0 > 32 ? 32 : V0;
    constexpr size_t TILE = size_t(kThreads) * V * EC_UNROLL;
    const D nd = from_bits<D>(static_cast<bits_t<D>>(nd_bits));
    return launch_k(Lc, fill_nodata_kernel<S, D, EC_VB, EC_UNROLL, kThreads>, grid_for(n, TILE, Lc), kThreads,
                    static_cast<const S*>(a), m, static_cast<D*>(out), n, nd);
}
cudaError_t launch_fill_nodata(const Launch& Lc, int sct, const void* a, const uint32_t* m, int dct, void* out, size_t n,
                               uint64_t nodata_bits) {
#define X(S, D) \
    if (sct == ct_of<S>::value && dct == ct_of<D>::value) return fill_nd<S, D>(Lc, a, m, out, n, nodata_bits);
    EC_LEGAL_CASTS(X)
#undef X
#define X(id, p) \
    if (sct == id && dct == id) return fill_nd<p, p>(Lc, a, m, out, n, nodata_bits);
    EC_WITH_CT(X)
#undef X
    return cudaErrorInvalidValue;
}

}  // namespace ec
