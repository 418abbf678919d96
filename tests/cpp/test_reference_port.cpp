// The reference crate's own unit tests, doc-tests and examples, ported line by line onto the C++ host
// mirror (include/erased_cells.hpp) of its API. Each test names the reference test it follows
// (file:line under /root/reference). Runs on the GPU box: every buffer op goes through the C ABI into
// the CUDA kernels. `tests/test_gpu_cpp_port.py` builds and runs this binary.
#include <cmath>
#include <cstdio>
#include <limits>

#include "erased_cells.hpp"

using namespace erased_cells;

static int g_checks = 0, g_failed = 0;
#define CHECK(cond) do { ++g_checks; if (!(cond)) { ++g_failed; std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); } } while (0)
#define CHECK_THROWS(expr, Ex) do { ++g_checks; bool t_ = false; try { (void)(expr); } catch (const Ex&) { t_ = true; } if (!t_) { ++g_failed; std::fprintf(stderr, "FAILED %s:%d: %s did not throw %s\n", __FILE__, __LINE__, #expr, #Ex); } } while (0)

template <class T> static std::vector<T> iota(size_t n, T start = 0) { std::vector<T> v(n); for (size_t i = 0; i < n; ++i) v[i] = T(start + T(i)); return v; }

// ---- src/ctype.rs tests ------------------------------------------------------------------------------
static void can_union() {  // src/ctype.rs:188-207
    CHECK(union_(CellType::UInt8, CellType::UInt8) == CellType::UInt8);
    CHECK(union_(CellType::UInt16, CellType::UInt16) == CellType::UInt16);
    CHECK(union_(CellType::Float32, CellType::Float32) == CellType::Float32);
    CHECK(union_(CellType::Float64, CellType::Float64) == CellType::Float64);
    CHECK(union_(CellType::Int16, CellType::Float32) == CellType::Float32);
    CHECK(union_(CellType::Float32, CellType::Int16) == CellType::Float32);
    CHECK(union_(CellType::UInt8, CellType::UInt16) == CellType::UInt16);
    CHECK(union_(CellType::Int32, CellType::Float32) == CellType::Float64);
}
static void ctype_misc() {  // src/ctype.rs:209-278
    CHECK(is_integral(CellType::UInt8) && is_integral(CellType::UInt16) && !is_integral(CellType::Float32) && !is_integral(CellType::Float64));
    const size_t sz[10] = {1, 2, 4, 8, 1, 2, 4, 8, 4, 8};
    for (CellType ct : cell_types()) CHECK(size_of(ct) == sz[int(ct)]);
    CHECK(min_value(CellType::Int16) == CellValue(std::numeric_limits<int16_t>::min()));
    CHECK(max_value(CellType::UInt64) == CellValue(std::numeric_limits<uint64_t>::max()));
    CHECK(min_value(CellType::Float32) == CellValue(std::numeric_limits<float>::lowest()));
    for (CellType ct : cell_types()) CHECK(cell_type_from_str(to_string(ct)) == ct);
    CHECK_THROWS(cell_type_from_str("UInt57"), ParseError);
    for (CellType ct : cell_types()) CHECK(one(ct) + zero(ct) == one(ct));
}
// ---- src/value.rs tests ------------------------------------------------------------------------------
static void value_tests() {
    CHECK(CellValue(uint8_t(43)).convert(CellType::Int16) == CellValue(int16_t(43)));  // :313-329
    CHECK(CellValue(uint8_t(43)).convert(CellType::Int16).cell_type() == CellType::Int16);
    CHECK_THROWS(CellValue(3.11111f).convert(CellType::Int32), NarrowingError);
    CHECK(CellValue(uint16_t(33)).convert(CellType::Float32).get<float>() == 33.0f);
    CHECK((-CellValue(uint8_t(1))).cell_type() == CellType::Int16 && (-CellValue(uint8_t(1))).get<int16_t>() == -1);  // :338-346
    CHECK((-CellValue(uint16_t(1))).cell_type() == CellType::Int32);
    CHECK((-CellValue(1.0)).get<double>() == -1.0 && (-CellValue(1.0f)).cell_type() == CellType::Float32);
    const CellValue l(uint8_t(1)), r(uint8_t(2));  // :349-391
    CHECK(l + r == CellValue(3.0)); CHECK(l + CellValue(2) == CellValue(3.0)); CHECK(l - r == CellValue(-1.0));
    CHECK(r - l == CellValue(1.0)); CHECK(l * r == CellValue(2.0)); CHECK(l / r == CellValue(0.5)); CHECK(r / l == CellValue(2.0));
    CHECK((l / r).cell_type() == CellType::Float64);
    CHECK(CellValue(1.0f) + CellValue(2.0f) == CellValue(3.0f));
    for (CellType ct : cell_types()) {  // get (:293-310)
        CHECK(zero(ct).get<double>() == 0.0);
    }
}
// ---- src/buffer.rs tests -----------------------------------------------------------------------------
static void buffer_tests() {
    for (CellType ct : cell_types()) {  // defaults :469-480, put_get :482-494
        CellBuffer cv = CellBuffer::with_defaults(3, ct);
        CHECK(cv.len() == 3 && cv.get(0) == zero(ct));
        CellBuffer f = CellBuffer::fill(3, zero(ct));
        f.put(1, one(ct));
        CHECK(f.get(1) == one(ct).convert(ct));
    }
    {  // extend :496-505
        CellBuffer buf = CellBuffer::fill(3, CellValue(uint8_t(0)));
        CHECK(!buf.is_empty() && buf.cell_type() == CellType::UInt8);
        buf.extend(std::vector<uint8_t>{1});
        CHECK(buf.cell_type() == CellType::UInt8 && buf.get(0) == CellValue(0) && buf.get(3) == CellValue(1));
    }
    {  // to_vec :507-519
        std::vector<float> v(3, 0.f);
        CHECK(CellBuffer::from_vec(v).to_vec<float>() == v);
        CHECK_THROWS(CellBuffer::from_vec(v).to_vec<int32_t>(), NarrowingError);
    }
    {  // min_max :515-526
        auto mm = CellBuffer::from_vec(std::vector<double>{-1.0, 3.0, 2000.0, -5555.5}).min_max();
        CHECK(mm.first == CellValue(-5555.5) && mm.second == CellValue(2000.0));
        auto m8 = CellBuffer::from_vec(std::vector<uint8_t>{1, 3, 200, 0}).min_max();
        CHECK(m8.first == CellValue(uint8_t(0)) && m8.second == CellValue(uint8_t(200)) && m8.first.cell_type() == CellType::UInt8);
    }
    {  // from_others :528-556
        CellBuffer b = CellBuffer::from_values({CellValue(uint16_t(3)), CellValue(uint16_t(4)), CellValue(uint16_t(5))});
        CHECK(b.cell_type() == CellType::UInt16 && b.len() == 3 && b.get(2) == CellValue(uint16_t(5)));
        CellBuffer f = CellBuffer::from_vec(std::vector<float>{33.3f, 44.4f, 55.5f});
        CHECK(f.cell_type() == CellType::Float32 && f.get(2) == CellValue(55.5f));
    }
    CHECK(CellBuffer::fill(5, CellValue(37)).debug().rfind("Int32CellBuffer", 0) == 0);  // debug :558-564
    CHECK(CellBuffer::fill(15, CellValue(37)).debug().find("...") != std::string::npos);
    CHECK(CellBuffer::from_vec(std::vector<uint16_t>{0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11}).debug() == "UInt16CellBuffer(0, 1, 2, 3, 4, ... 7, 8, 9, 10, 11)");
    CHECK(CellBuffer::fill(3, CellValue(uint8_t(1))).debug() == "UInt8CellBuffer(1, 1, 1)");  // Elided, src/lib.rs:198-206
    for (CellType ct : cell_types()) {  // convert :566-578
        CellBuffer buf = CellBuffer::with_defaults(3, ct);
        for (CellType target : cell_types()) {
            if (can_fit_into(ct, target)) CHECK(buf.convert(target).cell_type() == target);
            else CHECK_THROWS(buf.convert(target), NarrowingError);
        }
    }
    for (CellType ct : cell_types()) {  // unary :580-592
        CHECK((-CellBuffer::fill(3, one(ct))).get(0) == -one(ct));
    }
    for (CellType lct : cell_types()) {  // binary :595-614 — all 100 pairs
        const CellValue lv = one(lct);
        for (CellType rct : cell_types()) {
            const CellBuffer lhs = CellBuffer::fill(3, lv);
            const CellValue rv = one(rct) + one(rct);
            const CellBuffer rhs = CellBuffer::fill(3, rv);
            CHECK((lhs + rhs).get(0) == lv + rv); CHECK((rhs + lhs).get(1) == rv + lv);
            CHECK((lhs - rhs).get(2) == lv - rv); CHECK((rhs - lhs).get(0) == rv - lv);
            CHECK((lhs * rhs).get(1) == lv * rv); CHECK((rhs * lhs).get(2) == rv * lv);
            CHECK((lhs / rhs).get(0) == lv / rv); CHECK((rhs / lhs).get(1) == rv / lv);
        }
    }
    {  // scalar :617-621
        CellBuffer buf = CellBuffer::fill_via<uint8_t>(9, [](size_t i) { return uint8_t(i + 1); });
        CHECK(buf * 2.0 == CellBuffer::fill_via<double>(9, [](size_t i) { return (double(i) + 1.0) * 2.0; }));
    }
    {  // equal :623-636, cmp :638-672
        const double nan = std::numeric_limits<double>::quiet_NaN();
        CellBuffer buf = CellBuffer::fill_via<double>(9, [&](size_t i) { return i % 2 == 0 ? nan : double(i); });
        CHECK(buf == buf);
        CHECK(CellBuffer::with_defaults(4, CellType::UInt8) == CellBuffer::with_defaults(4, CellType::UInt8));
        CHECK(CellBuffer::with_defaults(4, CellType::UInt8) != CellBuffer::with_defaults(5, CellType::UInt8));
        CHECK(CellBuffer(std::vector<int32_t>{1, 2, 3}) < CellBuffer(std::vector<int32_t>{2, 3, 4}));
        CHECK(CellBuffer(std::vector<int32_t>{1, 2, 3}) < CellBuffer(std::vector<int32_t>{2, 3}));
        CHECK(CellBuffer(std::vector<double>{nan, 2.0, 3.0}) < CellBuffer(std::vector<double>{nan, 2.0, 4.0}));
        CHECK(CellBuffer::with_defaults(4, CellType::UInt8) < CellBuffer::with_defaults(4, CellType::Float32));
        CHECK(CellBuffer::with_defaults(5, CellType::Float64) > CellBuffer::with_defaults(4, CellType::Float64));
    }
    {  // README.md:22-33 / examples/quick.rs
        CellBuffer result = CellBuffer(std::vector<uint8_t>{1, 2, 3}) / CellBuffer(std::vector<uint16_t>{2, 4, 6}) * 0.5;
        CHECK(result == CellBuffer(std::vector<double>{0.25, 0.25, 0.25}));
    }
    {  // examples/buffer.rs
        CellBuffer buf1 = CellBuffer::fill_via<uint8_t>(9, [](size_t i) { return uint8_t(i); });
        auto mm = buf1.min_max();
        CHECK(mm.first == CellValue(uint8_t(0)) && mm.second == CellValue(uint8_t(8)));
        CHECK((mm.second - mm.first + CellValue(1)) / CellValue(2) == CellValue(4.5));
        CellBuffer buf2 = CellBuffer::fill_via<float>(9, [](size_t i) { return 8.0f - float(i); });
        CellBuffer diff = buf2 - buf1;
        auto dm = diff.min_max();
        CHECK(dm.first == CellValue(-8) && dm.second == CellValue(8));
    }
}
// ---- src/masked tests --------------------------------------------------------------------------------
static uint8_t filler(size_t i) { return uint8_t(i); }
static bool masker(size_t i) { return i % 2 == 0; }
static std::pair<uint8_t, bool> filler_masker(size_t i) { return {filler(i), masker(i)}; }

static void mask_tests() {  // src/masked/mask.rs:183-242
    CHECK(Mask::fill(3, true).counts() == std::make_pair(size_t(3), size_t(0)));
    CHECK(Mask::fill(3, false).counts() == std::make_pair(size_t(0), size_t(3)));
    CHECK(Mask::fill_via(3, masker).counts() == std::make_pair(size_t(2), size_t(1)));
    Mask m = Mask::fill(3, true);
    m.put(1, false); m.put(0, false);
    CHECK(m == Mask({false, false, true}));
    CHECK((!Mask::fill(4, true)) == Mask::fill(4, false));
    CHECK((!Mask({true, false, true, false})) == Mask({false, true, false, true}));
    Mask alt = Mask::fill_via(4, masker);
    CHECK(!alt.all(true) && !alt.all(false) && Mask::fill(4, true).all(true) && !Mask::fill(4, true).all(false));
    Mask l = Mask::fill_via(4, masker), r = Mask::fill_via(4, [](size_t i) { return i % 2 != 0; });
    CHECK((l & r).all(false)); CHECK((l | r).all(true));
}
static void nodata_tests() {  // src/masked/nodata.rs:74-95
    CHECK(!NoData<int16_t>(NoData<int16_t>::None).value().has_value());
    CHECK(NoData<uint8_t>().value().value() == 0);
    CHECK(std::isnan(NoData<float>().value().value()));
    CHECK(NoData<uint16_t>::new_(6).value().value() == 6);
    CHECK(NoData<double>().is(CellValue(std::numeric_limits<double>::quiet_NaN())));
}
static void masked_tests() {
    {  // ctor :400-410
        CHECK(MaskedCellBuffer::fill_via<uint8_t>(3, filler) == MaskedCellBuffer(CellBuffer::fill_via<uint8_t>(3, filler), Mask::fill(3, true)));
        CHECK(MaskedCellBuffer::from_vec(std::vector<double>(4, 0.0)).mask().counts() == std::make_pair(size_t(4), size_t(0)));
        CHECK_THROWS(MaskedCellBuffer(CellBuffer::with_defaults(4, CellType::UInt8), Mask::fill(5, true)), std::logic_error);
    }
    {  // vec_with_nodata :412-425
        const double nan = std::numeric_limits<double>::quiet_NaN();
        std::vector<double> v{1.0, nan, 3.0, nan};
        CHECK(MaskedCellBuffer::from_vec_with_nodata(v, NoData<double>()) == MaskedCellBuffer(CellBuffer(v), Mask({true, false, true, false})));
        CHECK(MaskedCellBuffer::from_vec_with_nodata(v, NoData<double>::new_(3.0)) == MaskedCellBuffer(CellBuffer(v), Mask({true, true, false, true})));
    }
    {  // get_masked :427-440
        MaskedCellBuffer buf = MaskedCellBuffer::fill_with_mask_via<uint8_t>(9, filler_masker);
        CHECK(buf.get(4) == CellValue(4) && buf.get_masked(4).value() == CellValue(4) && !buf.get_masked(5).has_value());
        buf.put(5, CellValue(uint8_t(4)));
        CHECK(!buf.get_masked(5).has_value());
        buf.mask_mut().put(5, true);
        CHECK(buf.get_masked(5).value() == CellValue(4));
        buf.put_with_mask(5, CellValue(uint8_t(99)), false);
        CHECK(!buf.get_masked(5).has_value());
    }
    {  // convert :442-447
        auto r = MaskedCellBuffer::fill_with_mask_via<uint8_t>(4, filler_masker).convert(CellType::Float64);
        CHECK(r.to_vec<double>() == (std::vector<double>{0.0, 1.0, 2.0, 3.0}));
    }
    {  // unary :464-479
        MaskedCellBuffer mbuf = MaskedCellBuffer::fill_with_mask_via<uint8_t>(9, filler_masker);
        const int16_t M = std::numeric_limits<int16_t>::min();
        CHECK((-mbuf).to_vec_with_nodata(NoData<int16_t>()) == (std::vector<int16_t>{0, M, -2, M, -4, M, -6, M, -8}));
    }
    {  // min_max :481-485
        auto mm = MaskedCellBuffer::fill_with_mask_via<uint8_t>(9, [](size_t i) { return std::make_pair(filler(i), i != 0 && i != 8); }).min_max();
        CHECK(mm.first == CellValue(uint8_t(1)) && mm.second == CellValue(uint8_t(7)));
    }
    {  // scalar :487-509
        MaskedCellBuffer all = MaskedCellBuffer::fill_with_mask_via<uint8_t>(9, [](size_t i) { return std::make_pair(filler(i), true); });
        CellBuffer expected = CellBuffer::fill_via<uint8_t>(9, filler) * 2.0;
        CHECK(all * 2.0 == MaskedCellBuffer(expected));
        MaskedCellBuffer r = MaskedCellBuffer::fill_with_mask_via<uint8_t>(9, filler_masker) * 2.0;
        CHECK(r != MaskedCellBuffer(expected));
        const double F = std::numeric_limits<double>::lowest();
        CHECK(r.to_vec_with_nodata(NoData<double>::new_(F)) == (std::vector<double>{0.0, F, 4.0, F, 8.0, F, 12.0, F, 16.0}));
    }
    {  // binary :511-531
        MaskedCellBuffer lhs(CellBuffer::fill(9, CellValue(1.0)), Mask::fill_via(9, masker));
        MaskedCellBuffer rhs(CellBuffer::fill(9, CellValue(2.0)), Mask::fill(9, true));
        CHECK((lhs + rhs).get_masked(0).value() == CellValue(3.0) && !(lhs + rhs).get_masked(1).has_value());
        CHECK((lhs - rhs).get_masked(2).value() == CellValue(-1.0) && !(lhs - rhs).get_masked(3).has_value());
        CHECK((lhs * rhs).get_masked(4).value() == CellValue(2.0) && !(lhs * rhs).get_masked(5).has_value());
        CHECK((lhs / rhs).get_masked(6).value() == CellValue(0.5) && !(lhs / rhs).get_masked(7).has_value());
    }
    {  // doc-test :15-38 / examples/masked.rs
        MaskedCellBuffer buf = MaskedCellBuffer::fill_with_mask_via<double>(4, [](size_t i) { return std::make_pair(double(i), i % 2 == 0); });
        CHECK(buf.mask() == Mask({true, false, true, false}) && buf.counts() == std::make_pair(size_t(2), size_t(2)));
        MaskedCellBuffer r = (buf + MaskedCellBuffer::from_vec(std::vector<double>(4, 1.0))) * 2.0;
        CHECK(r == MaskedCellBuffer(CellBuffer(std::vector<double>{2.0, 4.0, 6.0, 8.0}), Mask({true, false, true, false})));
    }
}

// not in the reference: the same operator chains deferred and fused must give the same buffers
static void lazy_tests() {
    std::vector<uint16_t> nir(5000), red(5000);
    for (size_t i = 0; i < nir.size(); ++i) { nir[i] = uint16_t(5000 + (i * 7919) % 30000); red[i] = uint16_t(i % 17 == 0 ? nir[i] : 5000 + (i * 104729) % 30000); }
    nir[3] = red[3] = 0;  // 0/0
    const CellBuffer n = CellBuffer::from_vec(nir), r = CellBuffer::from_vec(red);
    const CellBuffer eager = (n - r) / (n + r);
    const CellBuffer eager_chain = n / r * 0.5;
    const uint64_t before = ec_kernel_launches();
    {
        LazyScope lazy;
        CellBuffer ndvi = (n - r) / (n + r);
        CellBuffer chain = n / r * 0.5;
        CHECK(ec_kernel_launches() == before);            // nothing ran yet
        CHECK(ndvi == eager);                             // one fused kernel + the comparison
        CHECK(chain == eager_chain);
        CHECK(ndvi == n.normalized_difference(r));
        CHECK(chain == n.binary_scalar(EC_DIV, r, EC_MUL, CellValue(0.5)));
    }
    CHECK(ec_get_lazy() == 0);
}

// The statistics extension has no reference test to port: small cases whose exact answers are representable.
static void statistics_tests() {
    const Statistics s = CellBuffer::from_vec(std::vector<uint8_t>{2, 4, 4, 4, 5, 5, 7, 9}).statistics();
    CHECK(s.count == 8 && s.min == CellValue(uint8_t(2)) && s.max == CellValue(uint8_t(9)) && s.mean == 5.0 && s.stddev == 2.0);
    const Statistics f = CellBuffer::from_vec(std::vector<float>{1.5f, -0.5f, 2.5f, 0.5f}).statistics();
    CHECK(f.count == 4 && f.mean == 1.0 && f.stddev == std::sqrt(1.25));
    MaskedCellBuffer m = MaskedCellBuffer::from_vec_with_nodata(std::vector<int16_t>{-32768, 10, 20, -32768, 30}, NoData<int16_t>());
    const Statistics ms = m.statistics();
    CHECK(ms.count == 3 && ms.min == CellValue(int16_t(10)) && ms.max == CellValue(int16_t(30)) && ms.mean == 20.0);
    CHECK(ms.stddev == std::sqrt(200.0 / 3.0));
    const Statistics e = CellBuffer::with_defaults(0, CellType::Float64).statistics();
    CHECK(e.count == 0 && std::isnan(e.mean) && std::isnan(e.stddev));
    const Statistics inf = CellBuffer::from_vec(std::vector<double>{1.0, HUGE_VAL}).statistics();
    CHECK(inf.count == 2 && inf.mean == HUGE_VAL && std::isnan(inf.stddev));
    const Statistics big = CellBuffer::from_vec(std::vector<uint32_t>{4000000000u, 4000000002u}).statistics();
    CHECK(big.mean == 4000000001.0 && big.stddev == 1.0);
}

// Chunked ingest (the GDAL adapter's read_cells / read_cells_masked, src/gdal/rasterband.rs:81-126, fed block by block) and the
// staged copy of a pageable std::vector: both must give what the one-shot constructors give.
static void ingest_tests() {
    const size_t n = 300 * 1000 + 37;
    std::vector<uint16_t> band(n);
    uint32_t x = 7;
    for (auto& v : band) { x = x * 1664525u + 1013904223u; v = (x >> 28) == 0 ? 0 : uint16_t(x >> 12); }
    Ingest<uint16_t> in(n, NoData<uint16_t>::new_(0), 4096);
    size_t pos = 0;
    while (auto chunk = in.next()) {
        const size_t k = std::min(chunk.capacity, n - pos);
        std::copy(band.begin() + pos, band.begin() + pos + k, chunk.data);
        in.submit(k);
        pos += k;
    }
    CHECK(pos == n);
    const MaskedCellBuffer chunked = in.finish_masked();
    const MaskedCellBuffer one_shot = MaskedCellBuffer::from_vec_with_nodata(band, NoData<uint16_t>::new_(0));
    CHECK(chunked == one_shot && chunked.counts() == one_shot.counts() && chunked.counts().second > 0);
    Ingest<uint16_t> plain(n);
    pos = 0;
    while (auto chunk = plain.next()) {
        const size_t k = std::min(chunk.capacity, n - pos);
        std::copy(band.begin() + pos, band.begin() + pos + k, chunk.data);
        plain.submit(k);
        pos += k;
    }
    CHECK(plain.finish() == CellBuffer::from_vec(band));
    { Ingest<uint16_t> dropped(n); (void)dropped.next(); }  // abandoned half way: the destructor aborts it

    std::vector<double> big((size_t(3) << 20) + 5);  // 24 MiB of pageable memory: the staged copy
    for (size_t i = 0; i < big.size(); ++i) big[i] = double(i) * 0.25 - 1000.0;
    for (int threads : {0, 1, 6}) {
        const int prev = set_host_copy_threads(threads);
        const CellBuffer b = CellBuffer::from_vec(big);
        const std::vector<double> back = b.to_vec<double>();
        CHECK(back == big);
        set_host_copy_threads(prev);
    }
}

int main() {
    try {
        ingest_tests();
        can_union(); ctype_misc(); value_tests(); buffer_tests(); mask_tests(); nodata_tests(); masked_tests(); lazy_tests();
        statistics_tests();
    } catch (const std::exception& e) {
        std::fprintf(stderr, "exception: %s\n", e.what());
        return 2;
    }
    std::printf("%d checks, %d failed, %llu kernel launches\n", g_checks, g_failed, (unsigned long long)ec_kernel_launches());
    return g_failed ? 1 : 0;
}
