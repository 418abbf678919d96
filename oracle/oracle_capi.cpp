// ORACLE — TEST INFRASTRUCTURE ONLY (see erased_cells_oracle.hpp). C entry points so that the
// pytest suite, smoke() and bench.py's CPU-baseline legs can drive the restatement through ctypes.
// Buffers cross this boundary as (cell type tag, raw pointer, length); scalars as eco_value, which
// has the same 16-byte layout as the product's ec_value.
#include "erased_cells_oracle.hpp"

#include <chrono>

using namespace eco;

// Seconds spent inside the last faithful_* call made on this thread (input staging and result
// copy-out excluded), so the CPU-baseline legs time the restated algorithm only.
static thread_local double g_last_op_seconds = 0.0;
struct OpTimer {
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    ~OpTimer() { g_last_op_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
};

extern "C" {

struct eco_value { uint8_t ct; uint8_t pad[7]; uint64_t bits; };

static inline CellValue in(const eco_value& v) { CellValue c; c.ct = static_cast<CellType>(v.ct); c.bits = v.bits; return c; }
static inline eco_value out(const CellValue& c) {
    eco_value v; std::memset(&v, 0, sizeof v); v.ct = c.ct;
    std::memcpy(&v.bits, &c.bits, size_of(c.ct));  // keep unused high bytes zero
    return v;
}

int eco_abi_version() { return 1; }
double eco_last_op_seconds() { return g_last_op_seconds; }

// --- CellType ---------------------------------------------------------------------------------
int eco_ctype_union(int a, int b) { return union_(CellType(a), CellType(b)); }
int eco_ctype_can_fit_into(int a, int b) { return can_fit_into(CellType(a), CellType(b)); }
int eco_ctype_size_of(int a) { return (int)size_of(CellType(a)); }
int eco_ctype_is_integral(int a) { return is_integral(CellType(a)); }
int eco_ctype_is_signed(int a) { return is_signed(CellType(a)); }
const char* eco_ctype_name(int a) { return name(CellType(a)); }
int eco_ctype_from_str(const char* s) { CellType ct; return from_str(s, &ct) ? (int)ct : -1; }
void eco_ctype_min_value(int a, eco_value* o) { *o = out(min_value(CellType(a))); }
void eco_ctype_max_value(int a, eco_value* o) { *o = out(max_value(CellType(a))); }
void eco_ctype_zero(int a, eco_value* o) { *o = out(zero(CellType(a))); }
void eco_ctype_one(int a, eco_value* o) { *o = out(one(CellType(a))); }

// --- CellValue --------------------------------------------------------------------------------
int eco_value_convert(const eco_value* v, int ct, eco_value* o) {
    CellValue c;
    Status s = convert(in(*v), CellType(ct), &c);
    if (s == Ok) *o = out(c);
    return s;
}
void eco_value_binary(int op, const eco_value* l, const eco_value* r, eco_value* o) { *o = out(binary(Op(op), in(*l), in(*r))); }
void eco_value_neg(const eco_value* v, eco_value* o) { *o = out(neg(in(*v))); }
int eco_value_to_prim(const eco_value* v, int ct, eco_value* o) {  // default-chain to_<p>(); 0 = None
    std::optional<CellValue> r;
    switch (CellType(ct)) {
#define X(id, p) case id: r = to_prim<p>(in(*v)); break;
        ECO_WITH_CT(X)
#undef X
    }
    if (!r) return 0;
    *o = out(*r);
    return 1;
}
int eco_value_cmp(const eco_value* l, const eco_value* r) { return cmp(in(*l), in(*r)); }
int eco_value_to_f64(const eco_value* v, double* o) { auto r = to_f64(in(*v)); if (!r) return 1; *o = *r; return 0; }
int eco_value_to_i64(const eco_value* v, int64_t* o) { auto r = to_i64(in(*v)); if (!r) return 1; *o = *r; return 0; }
int eco_value_to_u64(const eco_value* v, uint64_t* o) { auto r = to_u64(in(*v)); if (!r) return 1; *o = *r; return 0; }

// --- CellBuffer (faithful flavour) ---------------------------------------------------------------
// Results are written into caller-provided storage `o` sized for the worst case; the result's cell
// type and length come back through out_ct/out_len (empty results are UInt8, src/buffer.rs:234).
static void emit(const CellBuffer& b, int* out_ct, size_t* out_len, void* o) {
    *out_ct = b.ct; *out_len = b.len;
    if (b.len) std::memcpy(o, b.bytes.data(), b.bytes.size());
}
void eco_buf_binary(int op, int lct, const void* l, size_t ln, int rct, const void* r, size_t rn,
                    int* out_ct, size_t* out_len, void* o) {
    const CellBuffer a = CellBuffer::from_raw(CellType(lct), l, ln), b = CellBuffer::from_raw(CellType(rct), r, rn);
    CellBuffer res;
    { OpTimer t; res = faithful_binary(Op(op), a, b); }
    emit(res, out_ct, out_len, o);
}
void eco_buf_scalar(int op, int lct, const void* l, size_t ln, const eco_value* r, int* out_ct, size_t* out_len, void* o) {
    const CellBuffer a = CellBuffer::from_raw(CellType(lct), l, ln);
    CellBuffer res;
    { OpTimer t; res = faithful_scalar(Op(op), a, in(*r)); }
    emit(res, out_ct, out_len, o);
}
void eco_buf_neg(int ct, const void* p, size_t n, int* out_ct, size_t* out_len, void* o) {
    const CellBuffer a = CellBuffer::from_raw(CellType(ct), p, n);
    CellBuffer res;
    { OpTimer t; res = faithful_neg(a); }
    emit(res, out_ct, out_len, o);
}
int eco_buf_convert(int ct, const void* p, size_t n, int dst, int* out_ct, size_t* out_len, void* o) {
    const CellBuffer a = CellBuffer::from_raw(CellType(ct), p, n);
    CellBuffer r;
    Status s;
    { OpTimer t; s = faithful_convert(a, CellType(dst), &r); }
    if (s == Ok) emit(r, out_ct, out_len, o);
    return s;
}
int eco_checked_cast(int ct, const void* p, size_t n, int dst, void* o) {
    CellBuffer r;
    Status s = faithful_checked_cast(CellBuffer::from_raw(CellType(ct), p, n), CellType(dst), &r);
    if (s == Ok && n) std::memcpy(o, r.bytes.data(), r.bytes.size());
    return s;
}
void eco_buf_min_max(int ct, const void* p, size_t n, const uint8_t* mask_or_null, eco_value* mn, eco_value* mx) {
    const CellBuffer buf = CellBuffer::from_raw(CellType(ct), p, n);
    std::pair<CellValue, CellValue> r;
    { OpTimer t; r = faithful_min_max(buf, mask_or_null); }
    *mn = out(r.first); *mx = out(r.second);
}
int eco_buf_cmp(int lct, const void* l, size_t ln, int rct, const void* r, size_t rn) {
    return faithful_buffer_cmp(CellBuffer::from_raw(CellType(lct), l, ln), CellBuffer::from_raw(CellType(rct), r, rn));
}
void eco_buf_fill(size_t n, const eco_value* v, void* o) {
    CellBuffer b = fill(n, in(*v));
    if (n) std::memcpy(o, b.bytes.data(), b.bytes.size());
}
int eco_buf_put(int ct, void* p, size_t n, size_t idx, const eco_value* v) {
    CellBuffer b = CellBuffer::from_raw(CellType(ct), p, n);
    Status s = b.put(idx, in(*v));
    if (s == Ok) std::memcpy(p, b.bytes.data(), b.bytes.size());
    return s;
}

// --- Mask / NoData ------------------------------------------------------------------------------
size_t eco_mask_and(const uint8_t* l, size_t ln, const uint8_t* r, size_t rn, uint8_t* o) {
    Mask m = mask_and(Mask(l, l + ln), Mask(r, r + rn));
    std::memcpy(o, m.data(), m.size()); return m.size();
}
size_t eco_mask_or(const uint8_t* l, size_t ln, const uint8_t* r, size_t rn, uint8_t* o) {
    Mask m = mask_or(Mask(l, l + ln), Mask(r, r + rn));
    std::memcpy(o, m.data(), m.size()); return m.size();
}
void eco_mask_not(const uint8_t* m, size_t n, uint8_t* o) {
    Mask r = mask_not(Mask(m, m + n));
    std::memcpy(o, r.data(), r.size());
}
void eco_mask_counts(const uint8_t* m, size_t n, size_t* data, size_t* nodata) {
    auto [d, nd] = mask_counts(Mask(m, m + n)); *data = d; *nodata = nd;
}
int eco_mask_all(const uint8_t* m, size_t n, int value) { return mask_all(Mask(m, m + n), value != 0); }

static NoData nd_of(int kind, int ct, const eco_value* v) {
    NoData nd; nd.kind = NoDataKind(kind); nd.ct = CellType(ct);
    nd.value = v ? in(*v) : zero(CellType(ct));
    return nd;
}
int eco_nodata_value(int kind, int ct, const eco_value* v, eco_value* o) {
    auto s = nodata_value(nd_of(kind, ct, v));
    if (!s) return 0;
    *o = out(*s); return 1;
}
int eco_nodata_is(int kind, int ct, const eco_value* nd, const eco_value* v) { return nodata_is(nd_of(kind, ct, nd), in(*v)); }
void eco_mask_from_nodata(int ct, const void* p, size_t n, int kind, const eco_value* nd, uint8_t* mask_out) {
    Mask m = faithful_mask_from_nodata(CellBuffer::from_raw(CellType(ct), p, n), nd_of(kind, ct, nd));
    if (n) std::memcpy(mask_out, m.data(), n);
}
int eco_fill_nodata(int ct, const void* p, size_t n, const uint8_t* mask, size_t mask_len, int dst_ct, int kind,
                    const eco_value* nd, size_t* out_len, void* o) {
    CellBuffer r;
    Status s = faithful_fill_nodata(CellBuffer::from_raw(CellType(ct), p, n), Mask(mask, mask + mask_len),
                                    nd_of(kind, dst_ct, nd), &r);
    if (s != Ok) return s;
    *out_len = r.len;
    if (r.len) std::memcpy(o, r.bytes.data(), r.bytes.size());
    return Ok;
}

// --- tight flavour (typed loops, same arithmetic) -------------------------------------------------
}  // extern "C"
template <class L> static void tight_binary_l(Op op, const L* l, int rct, const void* r, double* o, size_t n) {
    switch (CellType(rct)) {
#define X(id, p) case id: tight_binary_t<L, p>(op, l, static_cast<const p*>(r), o, n); break;
        ECO_WITH_CT(X)
#undef X
    }
}
extern "C" {
void eco_tight_binary(int op, int lct, const void* l, int rct, const void* r, size_t n, double* o) {
    switch (CellType(lct)) {
#define X(id, p) case id: tight_binary_l<p>(Op(op), static_cast<const p*>(l), rct, r, o, n); break;
        ECO_WITH_CT(X)
#undef X
    }
}
}  // extern "C"
template <class S> static void tight_convert_s(const S* s, int dct, void* d, size_t n) {
    switch (CellType(dct)) {
#define X(id, p) case id: tight_convert_t<S, p>(s, static_cast<p*>(d), n); break;
        ECO_WITH_CT(X)
#undef X
    }
}
extern "C" {
// Legal widenings only (the caller checks can_fit_into); `as` casts.
int eco_tight_convert(int sct, const void* s, size_t n, int dct, void* d) {
    if (!can_fit_into(CellType(sct), CellType(dct))) return NarrowingError;
    switch (CellType(sct)) {
#define X(id, p) case id: tight_convert_s<p>(static_cast<const p*>(s), dct, d, n); break;
        ECO_WITH_CT(X)
#undef X
    }
    return Ok;
}
}  // extern "C"
template <class T> static void tight_min_max_t(const T* p, size_t n, const uint8_t* mask, eco_value* mn, eco_value* mx) {
    T lo = std::numeric_limits<T>::max(), hi = std::numeric_limits<T>::lowest();
    if constexpr (std::is_floating_point_v<T>) {
        auto klo = total_key(lo), khi = total_key(hi);
        for (size_t i = 0; i < n; ++i) {
            if (mask && !mask[i]) continue;
            auto k = total_key(p[i]);
            if (k < klo) { klo = k; lo = p[i]; }
            if (k >= khi) { khi = k; hi = p[i]; }
        }
    } else {
        for (size_t i = 0; i < n; ++i) {
            if (mask && !mask[i]) continue;
            if (p[i] < lo) lo = p[i];
            if (p[i] > hi) hi = p[i];
        }
    }
    *mn = out(make<T>(lo)); *mx = out(make<T>(hi));
}
extern "C" {
void eco_tight_min_max(int ct, const void* p, size_t n, const uint8_t* mask_or_null, eco_value* mn, eco_value* mx) {
    switch (CellType(ct)) {
#define X(id, t) case id: tight_min_max_t<t>(static_cast<const t*>(p), n, mask_or_null, mn, mx); break;
        ECO_WITH_CT(X)
#undef X
    }
}

}  // extern "C"
