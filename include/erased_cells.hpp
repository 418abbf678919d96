// erased_cells.hpp — C++17 host mirror of the erased-cells crate API over the B200 C ABI.
//
// The reference is a Rust crate; this image has no Rust toolchain, so the host side above the C ABI
// (include/erased_cells_b200.h) is written in C++ with the reference's names, argument meaning and
// error behaviour: CellType (src/ctype.rs), CellEncoding (src/encoding.rs), CellValue (src/value.rs),
// CellBuffer + BufferOps (src/buffer.rs, src/lib.rs:104-163), Mask (src/masked/mask.rs), NoData
// (src/masked/nodata.rs), MaskedCellBuffer (src/masked/masked_buffer.rs). The Rust shim a maintainer
// would add is in INTEGRATION.md / rust/; it is the same calls.
//
// Rust -> C++ mapping: Result<T, Error> -> exceptions (NarrowingError, ...); panics -> std::out_of_range /
// std::logic_error; Clone -> copy constructor (deep device copy); Drop -> destructor; &A op &B and
// A op B -> the same operator on const references (inputs are never mutated).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <optional>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "erased_cells_b200.h"

namespace erased_cells {

// ---- errors — src/error.rs -------------------------------------------------------------------
struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};
enum class CellType : uint8_t { UInt8, UInt16, UInt32, UInt64, Int8, Int16, Int32, Int64, Float32, Float64 };
struct NarrowingError : Error {
    CellType src, dst;
    NarrowingError(CellType s, CellType d, const std::string& m) : Error(m), src(s), dst(d) {}
};
struct ParseError : Error {
    using Error::Error;
};

namespace detail {
inline void check(ec_status s) {
    if (s == EC_OK) return;
    const std::string msg = ec_last_error();
    switch (s) {
        case EC_NARROWING: {
            uint8_t a, b;
            ec_last_narrowing(&a, &b);
            throw NarrowingError(CellType(a), CellType(b), msg);
        }
        case EC_OOB: throw std::out_of_range(msg);            // index panics, src/lib.rs:136-147
        case EC_LEN_MISMATCH: throw std::logic_error(msg);    // assert_eq!, src/masked/masked_buffer.rs:49-53
        case EC_PARSE: throw ParseError(msg);
        default: throw Error("erased_cells_b200: " + msg);
    }
}
}  // namespace detail

// ---- one process, several GPUs (extension; the crate is single-threaded CPU code) -------------------------------------
// Call before the first buffer is made (or set EC_DEVICES=0,1,...): from then on buffers / masks of at least
// set_shard_min_cells() cells live as row strips, one per GPU, behind the very same CellBuffer / Mask objects.
inline int init_devices(const std::vector<int>& devices) {
    detail::check(ec_init_devices(devices.data(), static_cast<int>(devices.size())));
    return ec_device_count();
}
inline size_t set_shard_min_cells(size_t cells) { return ec_set_shard_min_cells(cells); }
// host threads that move pageable memory (std::vector) to / from pinned staging under the DMA; 0 = the driver's own path
inline int set_host_copy_threads(int threads) { return ec_set_host_copy_threads(threads); }

// ---- CellEncoding — src/encoding.rs:9-40 ----------------------------------------------------------
template <class T> struct CellEncoding;  // only the ten primitives are CellEncoding
#define EC_HPP_WITH_CT(X) \
    X(UInt8, uint8_t) X(UInt16, uint16_t) X(UInt32, uint32_t) X(UInt64, uint64_t) \
    X(Int8, int8_t) X(Int16, int16_t) X(Int32, int32_t) X(Int64, int64_t) X(Float32, float) X(Float64, double)
#define X(id, p) \
    template <> struct CellEncoding<p> { static constexpr CellType cell_type() { return CellType::id; } };
EC_HPP_WITH_CT(X)
#undef X

class CellValue;

// ---- CellType methods — src/ctype.rs ------------------------------------------------------------------
inline std::vector<CellType> cell_types() {  // CellType::iter
    std::vector<CellType> v;
    for (int i = 0; i < 10; ++i) v.push_back(CellType(i));
    return v;
}
inline bool is_integral(CellType ct) { return ec_ctype_is_integral(uint8_t(ct)); }
inline bool is_signed(CellType ct) { return ec_ctype_is_signed(uint8_t(ct)); }
inline size_t size_of(CellType ct) { return ec_ctype_size_of(uint8_t(ct)); }
inline CellType union_(CellType a, CellType b) { return CellType(ec_ctype_union(uint8_t(a), uint8_t(b))); }
inline bool can_fit_into(CellType a, CellType b) { return ec_ctype_can_fit_into(uint8_t(a), uint8_t(b)); }
inline std::string to_string(CellType ct) { return ec_ctype_name(uint8_t(ct)); }
inline CellType cell_type_from_str(const std::string& s) {
    uint8_t ct;
    detail::check(ec_ctype_from_name(s.c_str(), &ct));
    return CellType(ct);
}

// ---- CellValue — src/value.rs ---------------------------------------------------------------------------
class CellValue {
public:
    ec_value v{};
    CellValue() { v.ct = EC_UINT8; }
    explicit CellValue(const ec_value& raw) : v(raw) {}
    template <class T, class = decltype(CellEncoding<T>::cell_type())> CellValue(T x) {  // From<T: CellEncoding>
        std::memset(&v, 0, sizeof v);
        v.ct = uint8_t(CellEncoding<T>::cell_type());
        std::memcpy(&v.bits, &x, sizeof(T));
    }
    template <class T> static CellValue new_(T x) { return CellValue(x); }
    CellType cell_type() const { return CellType(v.ct); }
    CellValue convert(CellType ct) const {
        CellValue o;
        detail::check(ec_value_convert(&v, uint8_t(ct), &o.v));
        return o;
    }
    template <class T> T get() const {  // src/value.rs:51-67
        const CellValue c = convert(CellEncoding<T>::cell_type());
        T x;
        std::memcpy(&x, &c.v.bits, sizeof(T));
        return x;
    }
    std::pair<CellValue, CellValue> unify(const CellValue& o) const {
        const CellType d = union_(cell_type(), o.cell_type());
        return {convert(d), o.convert(d)};
    }
    std::optional<double> to_f64() const { double o; int some; detail::check(ec_value_to_f64(&v, &o, &some)); return some ? std::optional<double>(o) : std::nullopt; }
    std::optional<int64_t> to_i64() const { int64_t o; int some; detail::check(ec_value_to_i64(&v, &o, &some)); return some ? std::optional<int64_t>(o) : std::nullopt; }
    std::optional<uint64_t> to_u64() const { uint64_t o; int some; detail::check(ec_value_to_u64(&v, &o, &some)); return some ? std::optional<uint64_t>(o) : std::nullopt; }
    int cmp(const CellValue& o) const { int r; detail::check(ec_value_cmp(&v, &o.v, &r)); return r; }
    CellValue operator-() const { CellValue o; detail::check(ec_value_neg(&v, &o.v)); return o; }
    static CellValue binary(int op, const CellValue& l, const CellValue& r) {
        CellValue o;
        detail::check(ec_value_binary(op, &l.v, &r.v, &o.v));
        return o;
    }
};
inline CellValue operator+(const CellValue& l, const CellValue& r) { return CellValue::binary(EC_ADD, l, r); }
inline CellValue operator-(const CellValue& l, const CellValue& r) { return CellValue::binary(EC_SUB, l, r); }
inline CellValue operator*(const CellValue& l, const CellValue& r) { return CellValue::binary(EC_MUL, l, r); }
inline CellValue operator/(const CellValue& l, const CellValue& r) { return CellValue::binary(EC_DIV, l, r); }
inline bool operator==(const CellValue& l, const CellValue& r) { return l.cmp(r) == 0; }
inline bool operator!=(const CellValue& l, const CellValue& r) { return l.cmp(r) != 0; }
inline bool operator<(const CellValue& l, const CellValue& r) { return l.cmp(r) < 0; }
inline bool operator>(const CellValue& l, const CellValue& r) { return l.cmp(r) > 0; }

inline CellValue zero(CellType ct) { CellValue o; detail::check(ec_ctype_zero(uint8_t(ct), &o.v)); return o; }
inline CellValue one(CellType ct) { CellValue o; detail::check(ec_ctype_one(uint8_t(ct), &o.v)); return o; }
inline CellValue min_value(CellType ct) { CellValue o; detail::check(ec_ctype_min_value(uint8_t(ct), &o.v)); return o; }
inline CellValue max_value(CellType ct) { CellValue o; detail::check(ec_ctype_max_value(uint8_t(ct), &o.v)); return o; }

// Defer buffer arithmetic on this thread so op chains fuse into single passes over HBM (bit-identical results).
inline void set_lazy(bool on) { detail::check(ec_set_lazy(on ? 1 : 0)); }
struct LazyScope {  // RAII: `{ LazyScope lazy; auto ndvi = (nir - red) / (nir + red); ... }`
    int prev;
    // mode 1: the precompiled fused shapes; 3: also kernels specialised at run time for longer chains (needs libnvrtc)
    explicit LazyScope(int mode = 1) : prev(ec_get_lazy()) { detail::check(ec_set_lazy(mode)); }
    ~LazyScope() { ec_set_lazy(prev); }
    LazyScope(const LazyScope&) = delete;
    LazyScope& operator=(const LazyScope&) = delete;
};

// Launch overlap (process-wide, on by default): a kernel's CTAs are scheduled while the previous op's grid drains
// (programmatic dependent launch); results are identical either way. Returns the previous setting.
inline bool set_launch_overlap(bool on) { return ec_set_launch_overlap(on ? 1 : 0) == 1; }

// ---- CellBuffer — src/buffer.rs; BufferOps — src/lib.rs:104-163 -----------------------------------------------
// count / min / max / mean / population stddev of the valid cells — an extension (ec_statistics): the reference has no
// statistics beyond min_max and Mask::counts. Order-independent definition, see DESIGN.md §4.6.
struct Statistics {
    uint64_t count;
    CellValue min, max;
    double mean, stddev;
};

class Mask;
class CellBuffer {
    ec_buf* h_ = nullptr;
    static CellBuffer own(ec_buf* h) { CellBuffer b; b.h_ = h; return b; }
    friend class MaskedCellBuffer;
    friend class Mask;
    template <class T> friend class Ingest;

public:
    CellBuffer() = default;
    CellBuffer(const CellBuffer& o) { if (o.h_) detail::check(ec_buf_clone(o.h_, &h_)); }  // #[derive(Clone)]
    CellBuffer(CellBuffer&& o) noexcept : h_(o.h_) { o.h_ = nullptr; }
    CellBuffer& operator=(CellBuffer o) noexcept { std::swap(h_, o.h_); return *this; }
    ~CellBuffer() { ec_buf_free(h_); }
    template <class T> CellBuffer(const std::vector<T>& data) { detail::check(ec_buf_from_host(uint8_t(CellEncoding<T>::cell_type()), data.data(), data.size(), &h_)); detail::check(ec_synchronize()); }
    const ec_buf* handle() const { return h_; }

    template <class T> static CellBuffer from_vec(const std::vector<T>& data) { return CellBuffer(data); }
    static CellBuffer with_defaults(size_t len, CellType ct) { ec_buf* h; detail::check(ec_buf_with_defaults(len, uint8_t(ct), &h)); return own(h); }
    static CellBuffer fill(size_t len, const CellValue& value) { ec_buf* h; detail::check(ec_buf_fill(len, &value.v, &h)); return own(h); }
    template <class T, class F> static CellBuffer fill_via(size_t len, F f) {
        std::vector<T> v(len);
        for (size_t i = 0; i < len; ++i) v[i] = f(i);
        return from_vec(v);
    }
    // FromIterator<CellValue> (src/buffer.rs:229-250)
    static CellBuffer from_values(const std::vector<CellValue>& values) {
        if (values.empty()) return with_defaults(0, CellType::UInt8);
        CellBuffer b = with_defaults(values.size(), values[0].cell_type());
        for (size_t i = 0; i < values.size(); ++i) b.put(i, values[i]);
        return b;
    }
    size_t len() const { return ec_buf_len(h_); }
    bool is_empty() const { return len() == 0; }
    CellType cell_type() const { return CellType(ec_buf_ctype(h_)); }
    // row strips this buffer is kept as (0: one GPU) and where strip g lives
    int shard_count() const { return ec_buf_shard_count(h_); }
    ec_shard_info shard(int g) const { ec_shard_info i; detail::check(ec_buf_shard(h_, g, &i, nullptr)); return i; }
    // cells [offset, offset + len) as a buffer sharing this allocation (a row strip; offset on a 32-byte boundary)
    CellBuffer view(size_t offset, size_t len) const { ec_buf* h; detail::check(ec_buf_view(h_, offset, len, &h)); return own(h); }
    CellValue get(size_t index) const { CellValue o; detail::check(ec_buf_get(h_, index, &o.v)); return o; }
    void put(size_t index, const CellValue& value) { detail::check(ec_buf_put(h_, index, &value.v)); }
    template <class T> void extend(const std::vector<T>& more) { detail::check(ec_buf_extend_host(h_, uint8_t(CellEncoding<T>::cell_type()), more.data(), more.size())); }
    CellBuffer convert(CellType ct) const { ec_buf* h; detail::check(ec_buf_convert(h_, uint8_t(ct), &h)); return own(h); }
    std::pair<CellValue, CellValue> min_max() const {
        CellValue a, b;
        detail::check(ec_buf_min_max(h_, nullptr, &a.v, &b.v));
        return {a, b};
    }
    Statistics statistics() const {  // extension: the reference stops at min_max
        ec_statistics s;
        detail::check(ec_buf_statistics(h_, nullptr, &s));
        return Statistics{s.count, CellValue(s.min), CellValue(s.max), s.mean, s.stddev};
    }
    template <class T> std::vector<T> to_vec() const {  // src/buffer.rs:175-185
        const CellBuffer r = convert(CellEncoding<T>::cell_type());
        std::vector<T> out(r.len());
        if (r.len()) detail::check(ec_buf_to_host(r.h_, out.data(), out.size() * sizeof(T)));
        return out;
    }
    int cmp(const CellBuffer& o) const { int r; detail::check(ec_buf_cmp(h_, o.h_, &r)); return r; }
    static CellBuffer binary(int op, const CellBuffer& l, const CellBuffer& r) { ec_buf* h; detail::check(ec_buf_binary(op, l.h_, r.h_, &h)); return own(h); }
    static CellBuffer scalar(int op, const CellBuffer& l, const CellValue& r) { ec_buf* h; detail::check(ec_buf_scalar(op, l.h_, &r.v, &h)); return own(h); }
    CellBuffer operator-() const { ec_buf* h; detail::check(ec_buf_neg(h_, &h)); return own(h); }
    CellBuffer normalized_difference(const CellBuffer& o) const { ec_buf* h; detail::check(ec_buf_normalized_difference(h_, o.h_, &h)); return own(h); }
    CellBuffer binary_scalar(int op1, const CellBuffer& o, int op2, const CellValue& s) const {
        ec_buf* h;
        detail::check(ec_buf_binary_scalar(op1, h_, o.h_, op2, &s.v, &h));
        return own(h);
    }
    std::string debug() const {  // src/buffer.rs:188-203 + Elided: at most ten cells leave the device
        const size_t n = len();
        std::string s = to_string(cell_type()) + "CellBuffer(";
        auto item = [&](size_t i) { char b[64]; snprintf(b, sizeof b, "%g", get(i).to_f64().value_or(0.0)); return std::string(b); };
        if (n > 10) {
            for (size_t i = 0; i < 5; ++i) s += item(i) + (i < 4 ? ", " : "");
            s += ", ... ";
            for (size_t i = n - 5; i < n; ++i) s += item(i) + (i + 1 < n ? ", " : "");
        } else {
            for (size_t i = 0; i < n; ++i) s += item(i) + (i + 1 < n ? ", " : "");
        }
        return s + ")";
    }
};
#define EC_HPP_BUF_OP(sym, code)                                                                                         \
    inline CellBuffer operator sym(const CellBuffer& l, const CellBuffer& r) { return CellBuffer::binary(code, l, r); }  \
    inline CellBuffer operator sym(const CellBuffer& l, const CellValue& r) { return CellBuffer::scalar(code, l, r); }   \
    template <class T, class = decltype(CellEncoding<T>::cell_type())>                                                   \
    inline CellBuffer operator sym(const CellBuffer& l, T r) { return CellBuffer::scalar(code, l, CellValue(r)); }
EC_HPP_BUF_OP(+, EC_ADD) EC_HPP_BUF_OP(-, EC_SUB) EC_HPP_BUF_OP(*, EC_MUL) EC_HPP_BUF_OP(/, EC_DIV)
#undef EC_HPP_BUF_OP
inline bool operator==(const CellBuffer& l, const CellBuffer& r) { return l.cmp(r) == 0; }
inline bool operator!=(const CellBuffer& l, const CellBuffer& r) { return l.cmp(r) != 0; }
inline bool operator<(const CellBuffer& l, const CellBuffer& r) { return l.cmp(r) < 0; }
inline bool operator>(const CellBuffer& l, const CellBuffer& r) { return l.cmp(r) > 0; }

// ---- Mask — src/masked/mask.rs -------------------------------------------------------------------------------------
class Mask {
    ec_mask* h_ = nullptr;
    struct Adopt {};
    Mask(ec_mask* h, Adopt) : h_(h) {}
    static Mask own(ec_mask* h) { return Mask(h, Adopt{}); }
    friend class MaskedCellBuffer;
    template <class T> friend class Ingest;

public:
    Mask() { detail::check(ec_mask_fill(0, 1, &h_)); }  // Default
    Mask(const Mask& o) { detail::check(ec_mask_clone(o.h_, &h_)); }
    Mask(Mask&& o) noexcept : h_(o.h_) { o.h_ = nullptr; }
    Mask& operator=(Mask o) noexcept { std::swap(h_, o.h_); return *this; }
    ~Mask() { ec_mask_free(h_); }
    explicit Mask(const std::vector<bool>& values) {  // Mask::new
        std::vector<uint8_t> b(values.begin(), values.end());
        detail::check(ec_mask_from_bools(b.data(), b.size(), &h_));
        detail::check(ec_synchronize());
    }
    const ec_mask* handle() const { return h_; }
    static Mask fill(size_t len, bool value) { ec_mask* h; detail::check(ec_mask_fill(len, value, &h)); return own(h); }
    template <class F> static Mask fill_via(size_t len, F f) {
        std::vector<bool> v(len);
        for (size_t i = 0; i < len; ++i) v[i] = f(i);
        return Mask(v);
    }
    size_t len() const { return ec_mask_len(h_); }
    bool is_empty() const { return len() == 0; }
    void put(size_t index, bool value) { detail::check(ec_mask_put(h_, index, value)); }
    bool get(size_t index) const { int o; detail::check(ec_mask_get(h_, index, &o)); return o != 0; }
    bool operator[](size_t index) const { return get(index); }
    bool all(bool value) const { int o; detail::check(ec_mask_all(h_, value, &o)); return o != 0; }
    std::pair<size_t, size_t> counts() const { size_t d, nd; detail::check(ec_mask_counts(h_, &d, &nd)); return {d, nd}; }
    void extend(const std::vector<bool>& more) {
        std::vector<uint8_t> b(more.begin(), more.end());
        detail::check(ec_mask_extend_host(h_, b.data(), b.size()));
    }
    std::vector<bool> to_vec() const {
        std::vector<uint8_t> b(len());
        detail::check(ec_mask_to_bools(h_, b.data(), b.size()));
        return std::vector<bool>(b.begin(), b.end());
    }
    // validity bits of cells [offset, offset + len) as a mask of its own (a row strip; offset on a 128-cell boundary)
    Mask slice(size_t offset, size_t len) const { ec_mask* h; detail::check(ec_mask_slice(h_, offset, len, &h)); return own(h); }
    Mask operator!() const { ec_mask* h; detail::check(ec_mask_not(h_, &h)); return own(h); }
    Mask operator&(const Mask& o) const { ec_mask* h; detail::check(ec_mask_and(h_, o.h_, &h)); return own(h); }
    Mask operator|(const Mask& o) const { ec_mask* h; detail::check(ec_mask_or(h_, o.h_, &h)); return own(h); }
    int cmp(const Mask& o) const { int r; detail::check(ec_mask_cmp(h_, o.h_, &r)); return r; }
    bool operator==(const Mask& o) const { return cmp(o) == 0; }
    bool operator!=(const Mask& o) const { return cmp(o) != 0; }
};

// ---- NoData<T> — src/masked/nodata.rs ------------------------------------------------------------------------------
template <class T> struct NoData {
    enum Kind { None = EC_NODATA_NONE, Default = EC_NODATA_DEFAULT, Value = EC_NODATA_VALUE } kind = Default;
    T v{};
    NoData() = default;
    NoData(Kind k) : kind(k) {}
    static NoData new_(T value) { NoData n; n.kind = Value; n.v = value; return n; }
    std::optional<T> value() const {
        const CellValue cv(v);
        ec_value out;
        int has;
        detail::check(ec_nodata_value(kind, uint8_t(CellEncoding<T>::cell_type()), &cv.v, &out, &has));
        if (!has) return std::nullopt;
        T x;
        std::memcpy(&x, &out.bits, sizeof(T));
        return x;
    }
    bool is(const CellValue& value) const {  // :42-49
        const auto s = this->value();
        return s ? CellValue(*s) == value : false;
    }
};

// ---- MaskedCellBuffer — src/masked/masked_buffer.rs --------------------------------------------------------------------
class MaskedCellBuffer {
    CellBuffer buf_;
    Mask mask_;

public:
    MaskedCellBuffer(CellBuffer buffer, Mask mask) : buf_(std::move(buffer)), mask_(std::move(mask)) {  // new (:48-55)
        if (buf_.len() != mask_.len()) throw std::logic_error("Mask and buffer must have the same length.");
    }
    MaskedCellBuffer(CellBuffer buffer) : buf_(std::move(buffer)), mask_(Mask::fill(buf_.len(), true)) {}  // From<CellBuffer>
    template <class T> static MaskedCellBuffer from_vec(const std::vector<T>& data) { return MaskedCellBuffer(CellBuffer::from_vec(data)); }
    template <class T> static MaskedCellBuffer from_vec_with_nodata(const std::vector<T>& data, NoData<T> nodata) {  // :62-71
        CellBuffer b = CellBuffer::from_vec(data);
        const CellValue cv(nodata.v);
        ec_mask* m;
        detail::check(ec_mask_from_nodata(b.h_, nodata.kind, &cv.v, &m));
        return MaskedCellBuffer(std::move(b), Mask::own(m));
    }
    static MaskedCellBuffer with_defaults(size_t len, CellType ct) { return MaskedCellBuffer(CellBuffer::with_defaults(len, ct)); }
    static MaskedCellBuffer fill(size_t len, const CellValue& v) { return MaskedCellBuffer(CellBuffer::fill(len, v)); }
    template <class T, class F> static MaskedCellBuffer fill_via(size_t len, F f) { return MaskedCellBuffer(CellBuffer::fill_via<T>(len, f)); }
    template <class T, class F> static MaskedCellBuffer fill_with_mask_via(size_t len, F mv) {  // :73-79
        std::vector<T> d(len);
        std::vector<bool> m(len);
        for (size_t i = 0; i < len; ++i) { auto p = mv(i); d[i] = p.first; m[i] = p.second; }
        return MaskedCellBuffer(CellBuffer::from_vec(d), Mask(m));
    }
    // a row strip of a resident masked raster (offset on a 128-cell boundary, ec_row_strip)
    MaskedCellBuffer view(size_t offset, size_t len) const { return MaskedCellBuffer(buf_.view(offset, len), mask_.slice(offset, len)); }
    const CellBuffer& buffer() const { return buf_; }
    CellBuffer& buffer_mut() { return buf_; }
    const Mask& mask() const { return mask_; }
    Mask& mask_mut() { return mask_; }
    size_t len() const { return buf_.len(); }
    CellType cell_type() const { return buf_.cell_type(); }
    CellValue get(size_t i) const { return buf_.get(i); }
    void put(size_t i, const CellValue& v) { buf_.put(i, v); }
    std::optional<CellValue> get_masked(size_t i) const { return mask_.get(i) ? std::optional<CellValue>(buf_.get(i)) : std::nullopt; }
    std::pair<CellValue, bool> get_with_mask(size_t i) const { return {buf_.get(i), mask_.get(i)}; }
    void put_with_mask(size_t i, const CellValue& v, bool m) { buf_.put(i, v); mask_.put(i, m); }
    template <class T> void extend(const std::vector<std::pair<T, bool>>& more) {
        std::vector<T> d;
        std::vector<bool> m;
        for (auto& p : more) { d.push_back(p.first); m.push_back(p.second); }
        buf_.extend(d);
        mask_.extend(m);
    }
    std::pair<size_t, size_t> counts() const { return mask_.counts(); }
    MaskedCellBuffer convert(CellType ct) const { return MaskedCellBuffer(buf_.convert(ct), mask_); }
    std::pair<CellValue, CellValue> min_max() const {  // :208-217
        CellValue a, b;
        detail::check(ec_buf_min_max(buf_.h_, mask_.h_, &a.v, &b.v));
        return {a, b};
    }
    Statistics statistics() const {  // of the valid cells (extension)
        ec_statistics s;
        detail::check(ec_buf_statistics(buf_.h_, mask_.h_, &s));
        return Statistics{s.count, CellValue(s.min), CellValue(s.max), s.mean, s.stddev};
    }
    template <class T> std::vector<T> to_vec() const { return buf_.to_vec<T>(); }
    template <class T> std::vector<T> to_vec_with_nodata(NoData<T> no_data) const {  // :137-152
        const CellValue cv(no_data.v);
        ec_buf* h;
        detail::check(ec_buf_fill_nodata(buf_.h_, mask_.h_, uint8_t(CellEncoding<T>::cell_type()), no_data.kind, &cv.v, &h));
        CellBuffer filled = CellBuffer::own(h);
        std::vector<T> out(filled.len());
        if (filled.len()) detail::check(ec_buf_to_host(filled.h_, out.data(), out.size() * sizeof(T)));
        return out;
    }
    static MaskedCellBuffer binary(int op, const MaskedCellBuffer& l, const MaskedCellBuffer& r) {  // :326-336
        ec_buf* b;
        ec_mask* m;
        detail::check(ec_masked_binary(op, l.buf_.h_, l.mask_.h_, r.buf_.h_, r.mask_.h_, &b, &m));
        return MaskedCellBuffer(CellBuffer::own(b), Mask::own(m));
    }
    static MaskedCellBuffer scalar(int op, const MaskedCellBuffer& l, const CellValue& r) {  // :353-364
        return MaskedCellBuffer(CellBuffer::scalar(op, l.buf_, r), l.mask_);
    }
    MaskedCellBuffer operator-() const { return MaskedCellBuffer(-buf_, mask_); }
    bool operator==(const MaskedCellBuffer& o) const { return buf_ == o.buf_ && mask_ == o.mask_; }
    bool operator!=(const MaskedCellBuffer& o) const { return !(*this == o); }
};
// ---- chunked ingest (extension): what the GDAL adapter's read_cells / read_cells_masked (src/gdal/rasterband.rs:81-126)
// become when the reader produces the band block by block. The reader fills pinned staging buffers handed out one at
// a time; uploads, NoData compares and the reader overlap.
//     Ingest<uint16_t> in(len, NoData<uint16_t>::new_(0));
//     while (auto chunk = in.next()) { size_t n = read_rows(chunk.data, chunk.capacity); in.submit(n); }
//     MaskedCellBuffer band = in.finish_masked();
template <class T> class Ingest {
    ec_ingest* g_ = nullptr;
    bool masked_ = false;

public:
    struct Chunk {
        T* data = nullptr;
        size_t capacity = 0;  // cells
        explicit operator bool() const { return data != nullptr; }
    };
    explicit Ingest(size_t len, size_t chunk_cells = 0) {  // read_cells: no mask
        detail::check(ec_ingest_begin(uint8_t(CellEncoding<T>::cell_type()), len, EC_NODATA_NONE, nullptr, 0, chunk_cells, &g_));
    }
    Ingest(size_t len, NoData<T> nodata, size_t chunk_cells = 0) : masked_(true) {  // read_cells_masked
        const CellValue cv(nodata.v);
        detail::check(ec_ingest_begin(uint8_t(CellEncoding<T>::cell_type()), len, nodata.kind, &cv.v, 1, chunk_cells, &g_));
    }
    Ingest(const Ingest&) = delete;
    Ingest& operator=(const Ingest&) = delete;
    ~Ingest() { if (g_) ec_ingest_abort(g_); }
    Chunk next() {  // empty once every cell has been submitted
        void* p;
        size_t cap;
        detail::check(ec_ingest_next_buffer(g_, &p, &cap));
        Chunk c;
        c.data = static_cast<T*>(p);
        c.capacity = cap;
        return c;
    }
    void submit(size_t n_cells) { detail::check(ec_ingest_submit(g_, n_cells)); }
    CellBuffer finish() {
        ec_buf* b;
        ec_ingest* g = g_;
        detail::check(ec_ingest_finish(g, &b, nullptr));
        g_ = nullptr;
        return CellBuffer::own(b);
    }
    MaskedCellBuffer finish_masked() {
        if (!masked_) throw std::logic_error("Ingest: begun without a NoData (read_cells); use finish()");
        ec_buf* b;
        ec_mask* m;
        detail::check(ec_ingest_finish(g_, &b, &m));
        g_ = nullptr;
        return MaskedCellBuffer(CellBuffer::own(b), Mask::own(m));
    }
};

#define EC_HPP_MBUF_OP(sym, code)                                                                                                       \
    inline MaskedCellBuffer operator sym(const MaskedCellBuffer& l, const MaskedCellBuffer& r) { return MaskedCellBuffer::binary(code, l, r); } \
    inline MaskedCellBuffer operator sym(const MaskedCellBuffer& l, const CellValue& r) { return MaskedCellBuffer::scalar(code, l, r); }       \
    template <class T, class = decltype(CellEncoding<T>::cell_type())>                                                                  \
    inline MaskedCellBuffer operator sym(const MaskedCellBuffer& l, T r) { return MaskedCellBuffer::scalar(code, l, CellValue(r)); }
EC_HPP_MBUF_OP(+, EC_ADD) EC_HPP_MBUF_OP(-, EC_SUB) EC_HPP_MBUF_OP(*, EC_MUL) EC_HPP_MBUF_OP(/, EC_DIV)
#undef EC_HPP_MBUF_OP

}  // namespace erased_cells
