// ORACLE — TEST INFRASTRUCTURE ONLY. NOT PART OF THE PRODUCT PATH.
//
// CPU restatement of the per-cell compute path of s22s/erased-cells v0.1.1, written from the
// reference's Rust sources (cited as `src/...:line`, relative to /root/reference). Only `tests/`,
// `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may load this.
//
// Parity pinning: the reference cannot be compiled here (no rustc/cargo in the image), so this
// restatement is pinned against every known-answer test the reference holds for the path
// (tests/test_oracle_reference_kat.py ports them with file:line) and against the Landsat TIFF
// fixtures (tests/golden/). Cases no reference test exercises (u64/i64 -> f64 rounding, NaN payloads,
// signed MIN negation) are "parity unpinned" and follow Rust's documented `as` / IEEE semantics on
// x86-64, see DESIGN.md.
//
// Third-party arithmetic: num-traits 0.2.17 (Cargo.lock:106-113) is not vendored under
// /root/reference. Its `ToPrimitive` semantics are restated in `prim_to_*` below from the published
// algorithm of that release (cast.rs): int->int is range checked, int->float is `as` (RNE),
// float->float is `as`, float->int is truncation iff the value lies in the open interval
// (MIN-1, MAX+1), NaN -> None.
//
// Two flavours of every buffer op live here:
//   * faithful_*  — per-cell tagged dispatch through CellValue with the reference's 16-byte
//                   intermediate Vec<CellValue> and second collection pass (src/buffer.rs:229-250).
//                   This is what the CPU baseline times.
//   * tight_*     — the same arithmetic as typed loops, for big parity sweeps.
// tests/test_oracle_self.py asserts faithful == tight bit-for-bit.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <optional>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

namespace eco {

// ---------------------------------------------------------------------------------------------
// CellType — src/lib.rs:85-101 (with_ct! order fixes the discriminants), src/ctype.rs:11-20
// ---------------------------------------------------------------------------------------------
enum CellType : uint8_t {
    UInt8 = 0, UInt16 = 1, UInt32 = 2, UInt64 = 3,
    Int8 = 4, Int16 = 5, Int32 = 6, Int64 = 7,
    Float32 = 8, Float64 = 9,
};
constexpr int kNumCellTypes = 10;

template <class T> struct ct_of;
#define ECO_WITH_CT(X) \
    X(UInt8, uint8_t) X(UInt16, uint16_t) X(UInt32, uint32_t) X(UInt64, uint64_t) \
    X(Int8, int8_t) X(Int16, int16_t) X(Int32, int32_t) X(Int64, int64_t)         \
    X(Float32, float) X(Float64, double)
#define X(id, p) template <> struct ct_of<p> { static constexpr CellType value = id; };
ECO_WITH_CT(X)
#undef X

// src/ctype.rs:55-68
inline bool is_integral(CellType ct) { return ct != Float32 && ct != Float64; }
// src/ctype.rs:71-84
inline bool is_signed(CellType ct) { return ct >= Int8; }
// src/ctype.rs:87-96
inline size_t size_of(CellType ct) {
    switch (ct) {
#define X(id, p) case id: return sizeof(p);
        ECO_WITH_CT(X)
#undef X
    }
    return 0;
}

// src/ctype.rs:99-126
inline CellType union_(CellType self, CellType other) {
    size_t min_bytes;
    const bool si = is_integral(self), oi = is_integral(other);
    if (si && !oi) {
        min_bytes = std::max(size_of(other), 2 * size_of(self));
    } else if (!si && oi) {
        min_bytes = std::max(size_of(self), 2 * size_of(other));
    } else {
        const bool ss = is_signed(self), os = is_signed(other);
        if (ss && !os) min_bytes = std::max(size_of(self), 2 * size_of(other));
        else if (!ss && os) min_bytes = std::max(size_of(other), 2 * size_of(self));
        else min_bytes = std::max(size_of(self), size_of(other));
    }
    const bool sgn = is_signed(self) || is_signed(other);
    const bool integral = si && oi;
    if (min_bytes == 1 && !sgn && integral) return UInt8;
    if (min_bytes == 1 && sgn && integral) return Int8;
    if (min_bytes == 2 && !sgn && integral) return UInt16;
    if (min_bytes == 2 && sgn && integral) return Int16;
    if (min_bytes == 4 && !sgn && integral) return UInt32;
    if (min_bytes == 4 && sgn && integral) return Int32;
    if (min_bytes == 4 && !integral) return Float32;
    if (min_bytes == 8 && !sgn && integral) return UInt64;
    if (min_bytes == 8 && sgn && integral) return Int64;
    return Float64;
}
// src/ctype.rs:129-131
inline bool can_fit_into(CellType self, CellType other) { return union_(self, other) == other; }

inline const char* name(CellType ct) {
    switch (ct) {
#define X(id, p) case id: return #id;
        ECO_WITH_CT(X)
#undef X
    }
    return "?";
}
// src/ctype.rs:29-43 (FromStr); returns false for the ParseError case
inline bool from_str(const std::string& s, CellType* out) {
#define X(id, p) if (s == #id) { *out = id; return true; }
    ECO_WITH_CT(X)
#undef X
    return false;
}

// ---------------------------------------------------------------------------------------------
// CellValue — src/value.rs:12-20. A tagged 16-byte scalar.
// ---------------------------------------------------------------------------------------------
struct CellValue {
    CellType ct;
    union {
        uint8_t u8; uint16_t u16; uint32_t u32; uint64_t u64;
        int8_t i8; int16_t i16; int32_t i32; int64_t i64;
        float f32; double f64;
        uint64_t bits;
    };
};
static_assert(sizeof(CellValue) == 16, "CellValue mirrors the 16-byte Rust enum");

// src/value.rs:24-33 / src/encoding.rs:33 (into_cell_value)
template <class T> inline CellValue make(T x) {
    CellValue v;
    v.ct = ct_of<T>::value;
    v.bits = 0;
    std::memcpy(&v.bits, &x, sizeof(T));
    return v;
}
template <class T> inline T raw(const CellValue& v) {
    T x;
    std::memcpy(&x, &v.bits, sizeof(T));
    return x;
}

// src/ctype.rs:158-179, :134-155
inline CellValue min_value(CellType ct) {
    switch (ct) {
#define X(id, p) case id: return make<p>(std::numeric_limits<p>::lowest());
        ECO_WITH_CT(X)
#undef X
    }
    return make<uint8_t>(0);
}
inline CellValue max_value(CellType ct) {
    switch (ct) {
#define X(id, p) case id: return make<p>(std::numeric_limits<p>::max());
        ECO_WITH_CT(X)
#undef X
    }
    return make<uint8_t>(0);
}
inline CellValue zero(CellType ct) {
    switch (ct) {
#define X(id, p) case id: return make<p>(p(0));
        ECO_WITH_CT(X)
#undef X
    }
    return make<uint8_t>(0);
}
inline CellValue one(CellType ct) {
    switch (ct) {
#define X(id, p) case id: return make<p>(p(1));
        ECO_WITH_CT(X)
#undef X
    }
    return make<uint8_t>(1);
}

// ---------------------------------------------------------------------------------------------
// num-traits 0.2.17 ToPrimitive on primitives (restated; see header comment).
// ---------------------------------------------------------------------------------------------
template <class S> inline std::optional<int64_t> prim_to_i64(S v) {
    if constexpr (std::is_floating_point_v<S>) {
        // size_of::<f>() <= size_of::<i64>() branch of float_to_int: MIN inclusive, < 2^63
        const S lo = static_cast<S>(std::numeric_limits<int64_t>::min());
        const S hi = static_cast<S>(std::numeric_limits<int64_t>::max());  // rounds to 2^63
        if (v >= lo && v < hi) return static_cast<int64_t>(v);
        return std::nullopt;
    } else if constexpr (std::is_signed_v<S>) {
        return static_cast<int64_t>(v);
    } else {
        if (static_cast<uint64_t>(v) <= static_cast<uint64_t>(std::numeric_limits<int64_t>::max()))
            return static_cast<int64_t>(v);
        return std::nullopt;
    }
}
template <class S> inline std::optional<uint64_t> prim_to_u64(S v) {
    if constexpr (std::is_floating_point_v<S>) {
        const S hi = static_cast<S>(std::numeric_limits<uint64_t>::max());  // rounds to 2^64
        if (v > S(-1.0) && v < hi) return static_cast<uint64_t>(v);
        return std::nullopt;
    } else if constexpr (std::is_signed_v<S>) {
        if (v >= 0) return static_cast<uint64_t>(v);
        return std::nullopt;
    } else {
        return static_cast<uint64_t>(v);
    }
}
template <class S> inline std::optional<double> prim_to_f64(S v) {
    return static_cast<double>(v);  // `as f64`: RNE for u64/i64, exact otherwise
}
// narrowing int -> int of the default ToPrimitive chain: range check then `as`
template <class D> inline std::optional<D> i64_to(int64_t v) {
    if (v >= static_cast<int64_t>(std::numeric_limits<D>::min()) &&
        v <= static_cast<int64_t>(std::numeric_limits<D>::max()))
        return static_cast<D>(v);
    return std::nullopt;
}
template <class D> inline std::optional<D> u64_to(uint64_t v) {
    if (v <= static_cast<uint64_t>(std::numeric_limits<D>::max())) return static_cast<D>(v);
    return std::nullopt;
}

// ToPrimitive for CellValue — src/value.rs:118-157: only to_i64/to_u64/to_f64 are overridden,
// every other `to_<p>` goes through the trait's default chain.
inline std::optional<int64_t> to_i64(const CellValue& v) {
    switch (v.ct) {
#define X(id, p) case id: return prim_to_i64<p>(raw<p>(v));
        ECO_WITH_CT(X)
#undef X
    }
    return std::nullopt;
}
inline std::optional<uint64_t> to_u64(const CellValue& v) {
    switch (v.ct) {
#define X(id, p) case id: return prim_to_u64<p>(raw<p>(v));
        ECO_WITH_CT(X)
#undef X
    }
    return std::nullopt;
}
inline std::optional<double> to_f64(const CellValue& v) {
    switch (v.ct) {
#define X(id, p) case id: return prim_to_f64<p>(raw<p>(v));
        ECO_WITH_CT(X)
#undef X
    }
    return std::nullopt;
}
// default-chain `to_<p>` on a CellValue, returning the result wrapped as a CellValue of type D
template <class D> inline std::optional<CellValue> to_prim(const CellValue& v) {
    if constexpr (std::is_same_v<D, int64_t>) {
        auto r = to_i64(v); if (!r) return std::nullopt; return make<int64_t>(*r);
    } else if constexpr (std::is_same_v<D, uint64_t>) {
        auto r = to_u64(v); if (!r) return std::nullopt; return make<uint64_t>(*r);
    } else if constexpr (std::is_same_v<D, double>) {
        auto r = to_f64(v); if (!r) return std::nullopt; return make<double>(*r);
    } else if constexpr (std::is_same_v<D, float>) {
        // default to_f32: self.to_f64().and_then(f64::to_f32) ; f64::to_f32 is `as f32`
        auto r = to_f64(v); if (!r) return std::nullopt; return make<float>(static_cast<float>(*r));
    } else if constexpr (std::is_signed_v<D>) {
        auto r = to_i64(v); if (!r) return std::nullopt;
        auto n = i64_to<D>(*r); if (!n) return std::nullopt; return make<D>(*n);
    } else {
        auto r = to_u64(v); if (!r) return std::nullopt;
        auto n = u64_to<D>(*r); if (!n) return std::nullopt; return make<D>(*n);
    }
}

// ---------------------------------------------------------------------------------------------
// Errors — src/error.rs:12-27 (only NarrowingError is reachable from the hot path)
// ---------------------------------------------------------------------------------------------
enum Status : int { Ok = 0, NarrowingError = 1, OutOfBounds = 2, LengthMismatch = 3 };

// src/value.rs:74-98
inline Status convert(const CellValue& self, CellType cell_type, CellValue* out) {
    if (!can_fit_into(self.ct, cell_type)) return NarrowingError;
    if (cell_type == self.ct) { *out = self; return Ok; }
    std::optional<CellValue> r;
    switch (cell_type) {
#define X(id, p) case id: r = to_prim<p>(self); break;
        ECO_WITH_CT(X)
#undef X
    }
    if (!r) return NarrowingError;
    *out = *r;
    return Ok;
}

// src/value.rs:103-107
inline std::pair<CellValue, CellValue> unify(const CellValue& a, const CellValue& b) {
    const CellType dest = union_(a.ct, b.ct);
    CellValue l, r;
    convert(a, dest, &l);  // `unwrap` in the reference: union guarantees success
    convert(b, dest, &r);
    return {l, r};
}

enum Op : int { Add = 0, Sub = 1, Mul = 2, Div = 3 };

// The four scalar SSE2 instructions rustc emits for `f64 op f64` on x86-64, with the operand order
// pinned (destination = lhs) so NaN payload propagation does not depend on what this compiler
// chooses to commute: a NaN lhs wins, else a NaN rhs, else the x86 default NaN 0xFFF8000000000000.
inline double f64_op(Op op, double a, double b) {
#if defined(__x86_64__) && defined(__SSE2__)
    switch (op) {
        case Add: __asm__("addsd %1, %0" : "+x"(a) : "x"(b)); break;
        case Sub: __asm__("subsd %1, %0" : "+x"(a) : "x"(b)); break;
        case Mul: __asm__("mulsd %1, %0" : "+x"(a) : "x"(b)); break;
        case Div: __asm__("divsd %1, %0" : "+x"(a) : "x"(b)); break;
    }
    return a;
#else
#error "the oracle pins x86-64 SSE2 NaN semantics (the platform the reference is built for)"
#endif
}

// src/value.rs:199-222 (cv_bin_op!): unify, both to f64, f64 op, result is always Float64.
// The oracle is compiled with -ffp-contract=off -msse2 so these are addsd/subsd/mulsd/divsd,
// i.e. what rustc emits on the reference's platform (NaN sign/payload included).
inline CellValue binary(Op op, const CellValue& lhs, const CellValue& rhs) {
    auto [l, r] = unify(lhs, rhs);
    const double a = *to_f64(l), b = *to_f64(r);
    return make<double>(f64_op(op, a, b));
}

template <class F> inline F flip_sign(F v) {
    using U = std::conditional_t<sizeof(F) == 4, uint32_t, uint64_t>;
    U b; std::memcpy(&b, &v, sizeof(F));
    b ^= U(1) << (sizeof(F) * 8 - 1);
    std::memcpy(&v, &b, sizeof(F));
    return v;
}
template <class I> inline I wrapping_neg(I v) {
    using U = std::make_unsigned_t<I>;
    return static_cast<I>(U(0) - static_cast<U>(v));
}
// src/value.rs:224-240. Signed MIN: the reference panics in debug and wraps in release builds;
// the restatement (and the CUDA path) take the release behaviour (wrapping).
inline CellValue neg(const CellValue& v) {
    switch (v.ct) {
        case UInt8: return make<int16_t>(static_cast<int16_t>(-static_cast<int16_t>(v.u8)));
        case UInt16: return make<int32_t>(-static_cast<int32_t>(v.u16));
        case UInt32: return make<double>(flip_sign(static_cast<double>(v.u32)));
        case UInt64: return make<double>(flip_sign(static_cast<double>(v.u64)));
        case Int8: return make<int8_t>(wrapping_neg(v.i8));
        case Int16: return make<int16_t>(wrapping_neg(v.i16));
        case Int32: return make<int32_t>(wrapping_neg(v.i32));
        case Int64: return make<int64_t>(wrapping_neg(v.i64));
        case Float32: return make<float>(flip_sign(v.f32));
        case Float64: return make<double>(flip_sign(v.f64));
    }
    return v;
}

// f32/f64::total_cmp key (Rust core): bits ^ (((bits >> (w-1)) as unsigned) >> 1), compared signed
inline int32_t total_key(float f) {
    int32_t b; std::memcpy(&b, &f, 4);
    b ^= static_cast<int32_t>(static_cast<uint32_t>(b >> 31) >> 1);
    return b;
}
inline int64_t total_key(double f) {
    int64_t b; std::memcpy(&b, &f, 8);
    b ^= static_cast<int64_t>(static_cast<uint64_t>(b >> 63) >> 1);
    return b;
}
template <class T> inline int cmp3(T a, T b) { return a < b ? -1 : (a > b ? 1 : 0); }

// src/value.rs:248-265
inline int cmp(const CellValue& a, const CellValue& b) {
    auto [l, r] = unify(a, b);
    switch (l.ct) {
        case UInt8: return cmp3(l.u8, r.u8);
        case UInt16: return cmp3(l.u16, r.u16);
        case UInt32: return cmp3(l.u32, r.u32);
        case UInt64: return cmp3(l.u64, r.u64);
        case Int8: return cmp3(l.i8, r.i8);
        case Int16: return cmp3(l.i16, r.i16);
        case Int32: return cmp3(l.i32, r.i32);
        case Int64: return cmp3(l.i64, r.i64);
        case Float32: return cmp3(total_key(l.f32), total_key(r.f32));
        case Float64: return cmp3(total_key(l.f64), total_key(r.f64));
    }
    return 0;
}
inline bool eq(const CellValue& a, const CellValue& b) { return cmp(a, b) == 0; }  // src/value.rs:267-271

// src/value.rs:51-67 — get::<T>() as a CellValue of exactly type `want`
inline Status get_as(const CellValue& v, CellType want, CellValue* out) { return convert(v, want, out); }

// ---------------------------------------------------------------------------------------------
// CellBuffer — src/buffer.rs:12-55. Typed storage behind a tag.
// ---------------------------------------------------------------------------------------------
struct CellBuffer {
    CellType ct = UInt8;
    size_t len = 0;
    std::vector<uint8_t> bytes;

    static CellBuffer from_raw(CellType ct, const void* data, size_t n) {  // src/buffer.rs:64-66
        CellBuffer b; b.ct = ct; b.len = n; b.bytes.resize(n * size_of(ct));
        if (n) std::memcpy(b.bytes.data(), data, b.bytes.size());
        return b;
    }
    static CellBuffer with_defaults(size_t n, CellType ct) {  // src/buffer.rs:68-77
        CellBuffer b; b.ct = ct; b.len = n; b.bytes.assign(n * size_of(ct), 0);
        return b;
    }
    // src/buffer.rs:125-134 (panics when idx >= len in the reference; callers check)
    CellValue get(size_t idx) const {
        CellValue v; v.ct = ct; v.bits = 0;
        switch (ct) {
#define X(id, p) case id: std::memcpy(&v.bits, bytes.data() + idx * sizeof(p), sizeof(p)); break;
            ECO_WITH_CT(X)
#undef X
        }
        return v;
    }
    // src/buffer.rs:136-148
    Status put(size_t idx, const CellValue& value) {
        CellValue c;
        if (Status s = convert(value, ct, &c); s != Ok) return s;
        if (idx >= len) return OutOfBounds;
        std::memcpy(bytes.data() + idx * size_of(ct), &c.bits, size_of(ct));
        return Ok;
    }
};

// src/buffer.rs:79-88
inline CellBuffer fill(size_t n, const CellValue& value) {
    CellBuffer b = CellBuffer::with_defaults(n, value.ct);
    const size_t sz = size_of(value.ct);
    for (size_t i = 0; i < n; ++i) std::memcpy(b.bytes.data() + i * sz, &value.bits, sz);
    return b;
}

// FromIterator<CellValue> — src/buffer.rs:229-250: type of the first element; empty => UInt8([]).
inline CellBuffer collect(const std::vector<CellValue>& values) {
    if (values.empty()) return CellBuffer::with_defaults(0, UInt8);
    const CellType ct = values[0].ct;
    CellBuffer b = CellBuffer::with_defaults(values.size(), ct);
    const size_t sz = size_of(ct);
    for (size_t i = 0; i < values.size(); ++i) {
        CellValue c;
        get_as(values[i], ct, &c);  // v.get().unwrap()
        std::memcpy(b.bytes.data() + i * sz, &c.bits, sz);
    }
    return b;
}

// cb_bin_op! &A op &B — src/buffer.rs:324-329: zip (=> min length), per-cell CellValue op, collect
inline CellBuffer faithful_binary(Op op, const CellBuffer& l, const CellBuffer& r) {
    const size_t n = std::min(l.len, r.len);
    std::vector<CellValue> tmp;
    tmp.reserve(n);
    for (size_t i = 0; i < n; ++i) tmp.push_back(binary(op, l.get(i), r.get(i)));
    return collect(tmp);
}
// cb_bin_op! A op scalar — src/buffer.rs:346-352
inline CellBuffer faithful_scalar(Op op, const CellBuffer& l, const CellValue& r) {
    std::vector<CellValue> tmp;
    tmp.reserve(l.len);
    for (size_t i = 0; i < l.len; ++i) tmp.push_back(binary(op, l.get(i), r));
    return collect(tmp);
}
// Neg — src/buffer.rs:360-371
inline CellBuffer faithful_neg(const CellBuffer& b) {
    std::vector<CellValue> tmp;
    tmp.reserve(b.len);
    for (size_t i = 0; i < b.len; ++i) tmp.push_back(neg(b.get(i)));
    return collect(tmp);
}
// convert — src/buffer.rs:150-167
inline Status faithful_convert(const CellBuffer& b, CellType cell_type, CellBuffer* out) {
    if (cell_type == b.ct) { *out = b; return Ok; }
    if (!can_fit_into(b.ct, cell_type)) return NarrowingError;
    std::vector<CellValue> tmp;
    tmp.reserve(b.len);
    for (size_t i = 0; i < b.len; ++i) {
        CellValue c;
        convert(b.get(i), cell_type, &c);
        tmp.push_back(c);
    }
    *out = collect(tmp);
    return Ok;
}
// Extend<C> — src/buffer.rs:205-221: every appended cell goes through `c.into_cell_value().to_<p>().unwrap()`,
// i.e. the VALUE-checked ToPrimitive chain (not the type-checked convert); `None` panics in the reference,
// reported here as NarrowingError.
inline Status faithful_checked_cast(const CellBuffer& src, CellType dst, CellBuffer* out) {
    CellBuffer r = CellBuffer::with_defaults(src.len, dst);
    const size_t sz = size_of(dst);
    for (size_t i = 0; i < src.len; ++i) {
        const CellValue v = src.get(i);
        std::optional<CellValue> c;
        switch (dst) {
#define X(id, p) case id: c = to_prim<p>(v); break;
            ECO_WITH_CT(X)
#undef X
        }
        if (!c) return NarrowingError;
        std::memcpy(r.bytes.data() + i * sz, &c->bits, sz);
    }
    *out = std::move(r);
    return Ok;
}

// min_max — src/buffer.rs:169-173 (+ masked: src/masked/masked_buffer.rs:208-217); mask may be null
inline std::pair<CellValue, CellValue> faithful_min_max(const CellBuffer& b, const uint8_t* mask) {
    CellValue amin = max_value(b.ct), amax = min_value(b.ct);
    for (size_t i = 0; i < b.len; ++i) {
        if (mask && !mask[i]) continue;
        const CellValue v = b.get(i);
        if (cmp(v, amin) < 0) amin = v;   // Ord::min: returns `other` only when strictly smaller
        if (cmp(v, amax) >= 0) amax = v;  // Ord::max(self, other): returns `other` when self <= other
    }
    return {amin, amax};
}
// Ord for CellBuffer — src/buffer.rs:390-436: cell type first, then lexicographic (total order), then len
inline int faithful_buffer_cmp(const CellBuffer& l, const CellBuffer& r) {
    if (l.ct != r.ct) return l.ct < r.ct ? -1 : 1;
    const size_t n = std::min(l.len, r.len);
    for (size_t i = 0; i < n; ++i) {
        const int c = cmp(l.get(i), r.get(i));
        if (c) return c;
    }
    return cmp3(l.len, r.len);
}

// ---------------------------------------------------------------------------------------------
// Mask (Vec<bool>, 1 byte per cell) — src/masked/mask.rs:12-164
// ---------------------------------------------------------------------------------------------
using Mask = std::vector<uint8_t>;
inline Mask mask_and(const Mask& l, const Mask& r) {  // :129-140 (zip => min length)
    Mask o(std::min(l.size(), r.size()));
    for (size_t i = 0; i < o.size(); ++i) o[i] = l[i] & r[i];
    return o;
}
inline Mask mask_or(const Mask& l, const Mask& r) {  // :153-164
    Mask o(std::min(l.size(), r.size()));
    for (size_t i = 0; i < o.size(); ++i) o[i] = l[i] | r[i];
    return o;
}
inline Mask mask_not(const Mask& m) {  // :111-116
    Mask o(m.size());
    for (size_t i = 0; i < o.size(); ++i) o[i] = !m[i];
    return o;
}
inline std::pair<size_t, size_t> mask_counts(const Mask& m) {  // :72-80 -> (data, nodata)
    size_t d = 0, nd = 0;
    for (uint8_t b : m) { if (b) ++d; else ++nd; }
    return {d, nd};
}
inline bool mask_all(const Mask& m, bool value) {  // :67-69
    for (uint8_t b : m) if ((b != 0) != value) return false;
    return true;
}

// ---------------------------------------------------------------------------------------------
// NoData<T> — src/masked/nodata.rs:9-68
// ---------------------------------------------------------------------------------------------
enum NoDataKind : int { NoDataNone = 0, NoDataDefault = 1, NoDataValue = 2 };
struct NoData {
    NoDataKind kind;
    CellType ct;      // the static T of NoData<T>
    CellValue value;  // meaningful for NoDataValue
};
// :23-40 — Default sentinel: T::MIN for integers (0 for unsigned), NAN (positive quiet) for floats
inline std::optional<CellValue> nodata_value(const NoData& nd) {
    switch (nd.kind) {
        case NoDataNone: return std::nullopt;
        case NoDataValue: return nd.value;
        case NoDataDefault:
            if (nd.ct == Float32) return make<float>(std::numeric_limits<float>::quiet_NaN());
            if (nd.ct == Float64) return make<double>(std::numeric_limits<double>::quiet_NaN());
            return min_value(nd.ct);
    }
    return std::nullopt;
}
// :42-49 — total-order equality through CellValue ==
inline bool nodata_is(const NoData& nd, const CellValue& v) {
    auto s = nodata_value(nd);
    return s ? eq(*s, v) : false;
}

// from_vec_with_nodata — src/masked/masked_buffer.rs:62-71
inline Mask faithful_mask_from_nodata(const CellBuffer& b, const NoData& nd) {
    Mask m(b.len, 1);
    for (size_t i = 0; i < b.len; ++i) m[i] = !nodata_is(nd, b.get(i));
    return m;
}
// to_vec_with_nodata — src/masked/masked_buffer.rs:137-152: convert to T, then select
inline Status faithful_fill_nodata(const CellBuffer& b, const Mask& mask, const NoData& nd, CellBuffer* out) {
    CellBuffer conv;
    if (Status s = faithful_convert(b, nd.ct, &conv); s != Ok) return s;
    // to_vec::<T>() of an empty buffer: convert() yields UInt8([]) but danger::cast then asserts the
    // tag; an empty Vec<T> comes back when T == UInt8, otherwise the reference panics. The
    // restatement returns an empty buffer of type T.
    conv.ct = nd.ct;
    auto s = nodata_value(nd);
    if (s) {
        const size_t sz = size_of(nd.ct);
        const size_t n = std::min(conv.len, mask.size());  // zip
        conv.len = n; conv.bytes.resize(n * sz);
        for (size_t i = 0; i < n; ++i)
            if (!mask[i]) std::memcpy(conv.bytes.data() + i * sz, &s->bits, sz);
    }
    *out = std::move(conv);
    return Ok;
}

// ---------------------------------------------------------------------------------------------
// tight_* — the same arithmetic as typed loops (fast comparator for large sweeps)
// ---------------------------------------------------------------------------------------------
template <class L, class R> inline void tight_binary_t(Op op, const L* l, const R* r, double* o, size_t n) {
    for (size_t i = 0; i < n; ++i) o[i] = f64_op(op, static_cast<double>(l[i]), static_cast<double>(r[i]));
}
template <class S, class D> inline void tight_convert_t(const S* s, D* d, size_t n) {
    for (size_t i = 0; i < n; ++i) d[i] = static_cast<D>(s[i]);
}

}  // namespace eco
