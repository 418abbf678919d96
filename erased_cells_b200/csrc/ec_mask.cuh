// Validity masks as packed bit words in HBM: bit (i % 32) of little-endian word (i / 32) is
// `mask[i]` of the reference's Vec<bool> (src/masked/mask.rs:12); bits past `len` are zero.
// 1/8 byte per cell instead of the reference's 1 byte.
#pragma once
#include "ec_common.cuh"

namespace ec {

// Combine the V validity bits each lane computed for its V consecutive cells into 32-bit words
// (V | 32). Lanes whose (lane % (32/V)) == 0 end up holding a complete word.
template <int V> __device__ __forceinline__ uint32_t assemble_word(uint32_t bits, int lane) {
    if constexpr (V >= 32) return bits;
    constexpr int TPW = 32 / V;  // threads per word
    uint32_t w = bits << (V * (lane % TPW));
#pragma unroll
    for (int o = 1; o < TPW; o <<= 1) w |= __shfl_xor_sync(0xFFFFFFFFu, w, o);
    return w;
}

// from_vec_with_nodata (src/masked/masked_buffer.rs:62-71): mask[i] = cell[i] != sentinel, where
// CellValue equality is total-order equality == bitwise equality for same-typed cells
// (src/value.rs:248-271). T is the unsigned carrier of the cell width.
// The ragged tail runs one cell per lane and builds its words with __ballot_sync.
template <class U, bool PACK_BOOLS, int VB, int UNROLL, int THREADS>
__global__ void __launch_bounds__(THREADS) mask_build_kernel(const U* __restrict__ a, size_t n, U sentinel,
                                                             uint32_t* __restrict__ out, MaskCount mc) {
    constexpr int V0 = VB / sizeof(U);
    constexpr int V = V0 > 32 ? 32 : V0;
    constexpr int TPW = 32 / V;
    constexpr size_t TILE = size_t(THREADS) * V * UNROLL;
    const int lane = threadIdx.x & 31;
    const size_t full = n / TILE;
    unsigned int ones = 0;
    overlap_prologue();
    for (size_t t = blockIdx.x; t < full; t += gridDim.x) {
        const size_t base = t * TILE + size_t(threadIdx.x) * V;
        Vec<U, V> va[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) va[u] = ld_stream<U, V>(a + base + size_t(u) * THREADS * V);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            uint32_t bits = 0;
            if constexpr (sizeof(U) == 1 && V == 32) {
                // 8-bit cells, four at a time: "byte != sentinel" as a SWAR test (high bit of each byte set iff the xor is
                // non-zero), the four flags gathered into a nibble with one multiply — 2 instructions per cell instead of 3
                const uint32_t s4 = PACK_BOOLS ? 0u : static_cast<uint32_t>(sentinel) * 0x01010101u;
                uint32_t words[8];
                memcpy(words, &va[u], 32);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const uint32_t d = words[q] ^ s4;
                    const uint32_t nz = (((d & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | d) & 0x80808080u;
                    bits |= ((((nz >> 7) * 0x01020408u) >> 24) & 0xFu) << (4 * q);
                }
            } else {
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    const bool valid = PACK_BOOLS ? (va[u].v[j] != U(0)) : (va[u].v[j] != sentinel);
                    bits |= static_cast<uint32_t>(valid) << j;
                }
            }
            ones += __popc(bits);
            const uint32_t w = assemble_word<V>(bits, lane);
            if (lane % TPW == 0) out[(base + size_t(u) * THREADS * V) / 32] = w;
        }
    }
    if (blockIdx.x == full % gridDim.x) {
        const size_t start = full * TILE;  // multiple of 32
        for (size_t i0 = start + size_t(threadIdx.x & ~31); i0 < n; i0 += THREADS) {
            const size_t i = i0 + lane;
            bool valid = false;
            if (i < n) valid = PACK_BOOLS ? (a[i] != U(0)) : (a[i] != sentinel);
            const uint32_t w = __ballot_sync(0xFFFFFFFFu, valid);
            if (lane == 0) out[i0 / 32] = w;
            ones += valid ? 1u : 0u;
        }
    }
    if (mc.acc != nullptr) {
        const unsigned long long c = block_count<THREADS>(ones);
        if (threadIdx.x == 0) publish_count(mc, c);
    }
}

// Mask -> Vec<bool> bytes (IntoIterator, src/masked/mask.rs:171-177)
template <int THREADS>
__global__ void __launch_bounds__(THREADS) mask_unpack_kernel(const uint32_t* __restrict__ m, size_t n, uint8_t* __restrict__ out) {
    // each thread expands 16 mask bits into 16 bool bytes (one 16-byte store)
    const size_t groups = n / 16;
    for (size_t g = blockIdx.x * size_t(THREADS) + threadIdx.x; g < groups; g += size_t(gridDim.x) * THREADS) {
        const uint32_t bits = (__ldg(m + g / 2) >> ((g & 1) * 16)) & 0xFFFFu;
        Vec<uint32_t, 4> o;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t nib = (bits >> (4 * q)) & 0xFu;
            o.v[q] = (nib * 0x00204081u) & 0x01010101u;  // spread 4 bits to 4 bytes
        }
        st_stream<uint32_t, 4>(reinterpret_cast<uint32_t*>(out + g * 16), o);
    }
    if (blockIdx.x == 0)
        for (size_t i = groups * 16 + threadIdx.x; i < n; i += THREADS) out[i] = (m[i / 32] >> (i % 32)) & 1u;
}

// !, &, | on words (src/masked/mask.rs:103-164). `n` = cells of the result (min of the operands for
// the binary ops); the last word is trimmed so bits past n stay zero.
enum : int { MOP_NOT = 0, MOP_AND = 1, MOP_OR = 2 };
__device__ __forceinline__ uint32_t mask_word_op(int mop, uint32_t x, uint32_t y, size_t w, size_t words, size_t n) {
    uint32_t o = mop == MOP_NOT ? ~x : (mop == MOP_AND ? (x & y) : (x | y));
    if (w == words - 1 && (n % 32) != 0) o &= (1u << (n % 32)) - 1u;
    return o;
}
template <int THREADS>
__global__ void __launch_bounds__(THREADS) mask_bitop_kernel(int mop, const uint32_t* __restrict__ l,
                                                             const uint32_t* __restrict__ r, size_t n,
                                                             uint32_t* __restrict__ out, MaskCount mc) {
    const size_t words = (n + 31) / 32;
    const size_t groups = words / 4;
    unsigned int ones = 0;
    overlap_prologue();
    for (size_t g = blockIdx.x * size_t(THREADS) + threadIdx.x; g < groups; g += size_t(gridDim.x) * THREADS) {
        Vec<uint32_t, 4> x = ld_stream<uint32_t, 4>(l + 4 * g), y = x, o;
        if (mop != MOP_NOT) y = ld_stream<uint32_t, 4>(r + 4 * g);
#pragma unroll
        for (int q = 0; q < 4; ++q) { o.v[q] = mask_word_op(mop, x.v[q], y.v[q], 4 * g + q, words, n); ones += __popc(o.v[q]); }
        st_stream<uint32_t, 4>(out + 4 * g, o);
    }
    if (blockIdx.x == 0 && threadIdx.x < words % 4) {
        const size_t w = groups * 4 + threadIdx.x;
        const uint32_t x = mask_word_op(mop, l[w], mop == MOP_NOT ? 0u : r[w], w, words, n);
        out[w] = x;
        ones += __popc(x);
    }
    if (mc.acc != nullptr) {
        const unsigned long long c = block_count<THREADS>(ones);
        if (threadIdx.x == 0) publish_count(mc, c);
    }
}

// Mask::fill (src/masked/mask.rs:21-23)
template <int THREADS>
__global__ void __launch_bounds__(THREADS) mask_fill_kernel(uint32_t* __restrict__ out, size_t n, uint32_t word) {
    const size_t words = (n + 31) / 32;
    overlap_prologue();
    for (size_t w = blockIdx.x * size_t(THREADS) + threadIdx.x; w < words; w += size_t(gridDim.x) * THREADS) {
        uint32_t v = word;
        if (w == words - 1 && (n % 32) != 0) v &= (1u << (n % 32)) - 1u;
        out[w] = v;
    }
}

}  // namespace ec
