"""Pins the CPU oracle against every known-answer test the reference holds for the hot path.

Each test names the reference test it ports (file:line under /root/reference). The oracle is the
checker for the CUDA path, so it has to reproduce these first. CPU only.
"""
import math

import numpy as np
import pytest

CT = range(10)


# --- src/ctype.rs:188-207 can_union ---------------------------------------------------------------
def test_can_union(orc):
    o = orc
    assert o.union(o.UInt8, o.UInt8) == o.UInt8
    assert o.union(o.UInt16, o.UInt16) == o.UInt16
    assert o.union(o.Float32, o.Float32) == o.Float32
    assert o.union(o.Float64, o.Float64) == o.Float64
    assert o.union(o.Int16, o.Float32) == o.Float32
    assert o.union(o.Float32, o.Int16) == o.Float32
    assert o.union(o.UInt8, o.UInt16) == o.UInt16
    assert o.union(o.Int32, o.Float32) == o.Float64


# the full 10x10 lattice as derived from src/ctype.rs:99-126
UNION_TABLE = """
u8  u16 u32 u64 i16 i16 i32 i64 f32 f64
u16 u16 u32 u64 i32 i32 i32 i64 f32 f64
u32 u32 u32 u64 i64 i64 i64 i64 f64 f64
u64 u64 u64 u64 f64 f64 f64 f64 f64 f64
i16 i32 i64 f64 i8  i16 i32 i64 f32 f64
i16 i32 i64 f64 i16 i16 i32 i64 f32 f64
i32 i32 i64 f64 i32 i32 i32 i64 f64 f64
i64 i64 i64 f64 i64 i64 i64 i64 f64 f64
f32 f32 f64 f64 f32 f32 f64 f64 f32 f64
f64 f64 f64 f64 f64 f64 f64 f64 f64 f64
"""
SHORT = ["u8", "u16", "u32", "u64", "i8", "i16", "i32", "i64", "f32", "f64"]


def test_union_table_and_legal_pairs(orc):
    rows = [r.split() for r in UNION_TABLE.strip().splitlines()]
    legal = 0
    for a in CT:
        for b in CT:
            assert orc.union(a, b) == SHORT.index(rows[a][b]), (a, b)
            assert orc.union(a, b) == orc.union(b, a)
            legal += orc.can_fit_into(a, b)
    assert legal == 41


# --- src/ctype.rs:209-215, 217-228 ------------------------------------------------------------------
def test_is_integral_and_size(orc):
    assert orc.is_integral(orc.UInt8) and orc.is_integral(orc.UInt16)
    assert not orc.is_integral(orc.Float32) and not orc.is_integral(orc.Float64)
    assert [orc.size_of(c) for c in CT] == [1, 2, 4, 8, 1, 2, 4, 8, 4, 8]


# --- src/ctype.rs:231-243 has_min_max ---------------------------------------------------------------
def test_has_min_max(orc):
    for ct in CT:
        dt = orc.DTYPES[ct]
        info = np.iinfo(dt) if dt.kind in "iu" else np.finfo(dt)
        assert orc.min_value(ct).numpy() == info.min and orc.min_value(ct).ct == ct
        assert orc.max_value(ct).numpy() == info.max and orc.max_value(ct).ct == ct


# --- src/ctype.rs:245-264 can_string ----------------------------------------------------------------
def test_can_string(orc):
    for ct in CT:
        assert orc.name(ct) == orc.NAMES[ct]
        assert orc.from_str(orc.name(ct)) == ct
    assert orc.from_str("UInt57") == -1


# --- src/ctype.rs:266-278 zero_one ------------------------------------------------------------------
def test_zero_one(orc):
    for ct in CT:
        one, zero = orc.one(ct), orc.zero(ct)
        assert orc.value_cmp(orc.value_binary(orc.ADD, one, zero), one) == 0


# --- src/value.rs:293-310 get -----------------------------------------------------------------------
def test_value_get(orc):
    for ct in CT:
        v = orc.zero(ct)
        assert orc.value_convert(v, ct).key() == v.key()
        r2 = orc.value_convert(v, orc.Float64)
        assert r2.ct == orc.Float64 and r2.numpy() == 0.0


# --- src/value.rs:313-329 convert -------------------------------------------------------------------
def test_value_convert(orc):
    r = orc.value_convert(orc.value(orc.UInt8, 43), orc.Int16)
    assert r.ct == orc.Int16 and r.numpy() == 43
    with pytest.raises(orc.NarrowingError):
        orc.value_convert(orc.value(orc.Float32, 3.11111), orc.Int32)
    r = orc.value_convert(orc.value(orc.Float32, 3.11111), orc.Float32)
    assert r.ct == orc.Float32 and r.numpy() == np.float32(3.11111)
    r = orc.value_convert(orc.value(orc.UInt16, 33), orc.Float32)
    assert r.ct == orc.Float32 and r.numpy() == np.float32(33.0)


# --- src/value.rs:338-346 unary ---------------------------------------------------------------------
def test_value_unary(orc):
    def chk(ct, x, ect, ex):
        r = orc.value_neg(orc.value(ct, x))
        assert (r.ct, r.numpy()) == (ect, ex)

    chk(orc.UInt8, 1, orc.Int16, -1)
    chk(orc.UInt16, 1, orc.Int32, -1)
    chk(orc.Int8, 1, orc.Int8, -1)
    chk(orc.Int16, 1, orc.Int16, -1)
    chk(orc.Float64, 1.0, orc.Float64, -1.0)
    chk(orc.Float32, 1.0, orc.Float32, -1.0)


# --- src/value.rs:349-391 binops --------------------------------------------------------------------
@pytest.mark.parametrize("ct", [0, 1, 8, 9])
def test_value_binops(orc, ct):
    l, r = orc.value(ct, 1), orc.value(ct, 2)
    f64 = lambda x: orc.value(orc.Float64, x)
    exp = [
        (orc.ADD, l, r, 3.0), (orc.SUB, l, r, -1.0), (orc.SUB, r, l, 1.0), (orc.MUL, l, r, 2.0),
        (orc.MUL, r, l, 2.0), (orc.DIV, l, r, 0.5), (orc.DIV, r, l, 2.0),
    ]
    for op, a, b, e in exp:
        got = orc.value_binary(op, a, b)
        assert got.ct == orc.Float64  # src/value.rs:196 — every op promotes to f64
        assert orc.value_cmp(got, f64(e)) == 0
        assert orc.value_cmp(got, orc.value(orc.Float32, e)) == 0  # the f32 rows compare against Float32(..)
    # `l + 2` / `l - 2`: an i32 literal on the right (src/value.rs:353,355)
    assert orc.value_cmp(orc.value_binary(orc.ADD, l, orc.value(orc.Int32, 2)), f64(3.0)) == 0
    assert orc.value_cmp(orc.value_binary(orc.SUB, l, orc.value(orc.Int32, 2)), f64(-1.0)) == 0


# --- src/buffer.rs:469-480 defaults, :482-494 put_get ----------------------------------------------
def test_buffer_defaults_put_get(orc):
    for ct in CT:
        buf = np.zeros(3, dtype=orc.DTYPES[ct])
        assert orc.put(buf, 1, orc.one(ct)) == orc.OK
        assert buf[1] == 1 and buf[0] == 0
        # a narrowing put is rejected (src/buffer.rs:137)
        if ct != orc.Float64:
            assert orc.put(buf, 1, orc.value(orc.Float64, 1.0)) == orc.NARROWING


# --- src/buffer.rs:515-526 min_max ------------------------------------------------------------------
def test_buffer_min_max(orc):
    mn, mx = orc.min_max(np.array([-1.0, 3.0, 2000.0, -5555.5]))
    assert (mn.ct, mn.numpy(), mx.ct, mx.numpy()) == (orc.Float64, -5555.5, orc.Float64, 2000.0)
    mn, mx = orc.min_max(np.array([1, 3, 200, 0], dtype=np.uint8))
    assert (mn.ct, mn.numpy(), mx.ct, mx.numpy()) == (orc.UInt8, 0, orc.UInt8, 200)


# --- src/buffer.rs:22-48 doc-test / examples/buffer.rs ---------------------------------------------
def test_buffer_example(orc):
    buf1 = np.arange(9, dtype=np.uint8)
    mn, mx = orc.min_max(buf1)
    assert (mn.key(), mx.key()) == (orc.value(orc.UInt8, 0).key(), orc.value(orc.UInt8, 8).key())
    # ((max - min + 1) / 2) == 4.5
    t = orc.value_binary(orc.SUB, mx, mn)
    t = orc.value_binary(orc.ADD, t, orc.value(orc.Int32, 1))
    t = orc.value_binary(orc.DIV, t, orc.value(orc.Int32, 2))
    assert t.ct == orc.Float64 and t.numpy() == 4.5
    buf2 = (8 - np.arange(9)).astype(np.float32)
    mn2, mx2 = orc.min_max(buf2)
    assert (mn2.ct, mn2.numpy(), mx2.numpy()) == (orc.Float32, 0.0, 8.0)
    diff = orc.binary(orc.SUB, buf2, buf1)
    assert diff.dtype == np.float64
    dmn, dmx = orc.min_max(diff)
    assert orc.value_cmp(dmn, orc.value(orc.Int32, -8)) == 0 and orc.value_cmp(dmx, orc.value(orc.Int32, 8)) == 0


# --- src/buffer.rs:566-578 convert ------------------------------------------------------------------
def test_buffer_convert_legality(orc):
    for ct in CT:
        buf = np.zeros(3, dtype=orc.DTYPES[ct])
        for target in CT:
            if orc.can_fit_into(ct, target):
                r = orc.convert(buf, target)
                assert r.dtype == orc.DTYPES[target] and len(r) == 3
            else:
                with pytest.raises(orc.NarrowingError) as e:
                    orc.convert(buf, target)
                assert (e.value.src, e.value.dst) == (ct, target)


# --- src/buffer.rs:580-592 unary --------------------------------------------------------------------
def test_buffer_unary(orc):
    for ct in CT:
        one = orc.one(ct)
        buf = orc.neg(orc.fill(3, one))
        exp = orc.value_neg(one)
        assert buf.dtype == orc.DTYPES[exp.ct] and buf[0] == exp.numpy()


# --- src/buffer.rs:595-614 binary: all 100 type pairs x 4 ops, both operand orders ------------------
def test_buffer_binary_all_pairs(orc):
    for lct in CT:
        lv = orc.one(lct)
        for rct in CT:
            rv = orc.value_binary(orc.ADD, orc.one(rct), orc.one(rct))  # Float64(2.0)
            lhs, rhs = orc.fill(3, lv), orc.fill(3, rv)
            for op in orc.OPS:
                a = orc.binary(op, lhs, rhs)
                b = orc.binary(op, rhs, lhs)
                assert a.dtype == np.float64 and b.dtype == np.float64
                assert a[0] == orc.value_binary(op, lv, rv).numpy()
                assert b[1] == orc.value_binary(op, rv, lv).numpy()
    # the plain numbers behind it
    assert orc.binary(orc.DIV, np.ones(3, np.uint8), np.full(3, 2.0))[2] == 0.5


# --- src/buffer.rs:617-621 scalar -------------------------------------------------------------------
def test_buffer_scalar(orc):
    buf = (np.arange(9) + 1).astype(np.uint8)
    r = orc.scalar(orc.MUL, buf, orc.value(orc.Float64, 2.0))
    exp = (np.arange(9, dtype=np.float64) + 1.0) * 2.0
    assert r.dtype == np.float64  # equality in the reference includes the cell type
    assert orc.buffer_cmp(r, exp) == 0


# --- README.md:22-33 / examples/quick.rs ------------------------------------------------------------
def test_quick_example(orc):
    r = orc.binary(orc.DIV, np.array([1, 2, 3], np.uint8), np.array([2, 4, 6], np.uint16))
    r = orc.scalar(orc.MUL, r, orc.value(orc.Float64, 0.5))
    assert r.dtype == np.float64 and orc.buffer_cmp(r, np.array([0.25, 0.25, 0.25])) == 0


# --- src/buffer.rs:623-672 equal / cmp --------------------------------------------------------------
def test_buffer_equal_cmp(orc):
    buf = np.array([np.nan if i % 2 == 0 else float(i) for i in range(9)])
    assert orc.buffer_cmp(buf, buf) == 0
    z = lambda n, ct: np.zeros(n, dtype=orc.DTYPES[ct])
    assert orc.buffer_cmp(z(4, 0), z(4, 0)) == 0
    assert orc.buffer_cmp(z(4, 0), z(5, 0)) != 0
    i32 = lambda *a: np.array(a, dtype=np.int32)
    assert orc.buffer_cmp(i32(1, 2, 3), i32(2, 3, 4)) < 0
    assert orc.buffer_cmp(i32(1, 2, 3), i32(2, 3)) < 0
    assert orc.buffer_cmp(np.array([np.nan, 2.0, 3.0]), np.array([np.nan, 2.0, 4.0])) < 0
    assert orc.buffer_cmp(z(4, orc.UInt8), z(4, orc.Float32)) < 0
    assert orc.buffer_cmp(z(4, orc.Float32), z(4, orc.UInt8)) > 0
    assert orc.buffer_cmp(z(4, 0), z(5, 0)) < 0 and orc.buffer_cmp(z(5, 0), z(4, 0)) > 0
    assert orc.buffer_cmp(z(4, 9), z(5, 9)) < 0 and orc.buffer_cmp(z(5, 9), z(4, 9)) > 0


# --- src/buffer.rs:229-236: an empty result is UInt8([]) --------------------------------------------
def test_empty_result_is_uint8(orc):
    e = np.zeros(0, np.float32)
    assert orc.binary(orc.ADD, e, e).dtype == np.uint8
    assert orc.convert(e, orc.Float64).dtype == np.uint8
    assert orc.convert(e, orc.Float32).dtype == np.float32  # same type => clone (src/buffer.rs:151)


# --- src/masked/mask.rs:183-242 ---------------------------------------------------------------------
def test_mask_ops(orc):
    T, F = True, False
    assert orc.mask_counts([T] * 3) == (3, 0)
    assert orc.mask_counts([F] * 3) == (0, 3)
    assert orc.mask_counts([i % 2 == 0 for i in range(3)]) == (2, 1)
    assert list(orc.mask_not([T] * 4)) == [F] * 4
    assert list(orc.mask_not([T, F, T, F])) == [F, T, F, T]
    alt = [i % 2 == 0 for i in range(4)]
    assert not orc.mask_all(alt, True) and not orc.mask_all(alt, False)
    assert orc.mask_all([T] * 4, True) and not orc.mask_all([T] * 4, False)
    l, r = alt, [i % 2 != 0 for i in range(4)]
    assert orc.mask_all(orc.mask_and(l, r), False)
    assert orc.mask_all(orc.mask_or(l, r), True)
    # zip semantics: the shorter operand decides (src/masked/mask.rs:133-137)
    assert len(orc.mask_and([T] * 5, [T] * 3)) == 3


# --- src/masked/nodata.rs:74-95 ---------------------------------------------------------------------
def test_nodata_defaults(orc):
    assert orc.nodata_value(orc.ND_NONE, orc.Int16) is None
    assert orc.nodata_value(orc.ND_DEFAULT, orc.UInt8).numpy() == 0
    assert math.isnan(orc.nodata_value(orc.ND_DEFAULT, orc.Float32).numpy())
    assert orc.nodata_value(orc.ND_VALUE, orc.UInt16, orc.value(orc.UInt16, 6)).numpy() == 6
    for ct in CT:
        assert orc.nodata_value(orc.ND_DEFAULT, ct) is not None
    assert orc.nodata_is(orc.ND_DEFAULT, orc.Float64, None, orc.value(orc.Float64, float("nan")))
    # integer defaults are T::MIN (src/masked/nodata.rs:27-37)
    assert orc.nodata_value(orc.ND_DEFAULT, orc.Int16).numpy() == -32768


# --- src/masked/masked_buffer.rs:412-425 vec_with_nodata --------------------------------------------
def test_vec_with_nodata(orc):
    v = np.array([1.0, np.nan, 3.0, np.nan])
    assert list(orc.mask_from_nodata(v, orc.ND_DEFAULT)) == [True, False, True, False]
    assert list(orc.mask_from_nodata(v, orc.ND_VALUE, orc.value(orc.Float64, 3.0))) == [True, True, False, True]
    assert list(orc.mask_from_nodata(v, orc.ND_NONE)) == [True] * 4


# --- src/masked/masked_buffer.rs:442-447 convert ----------------------------------------------------
def test_masked_convert(orc):
    r = orc.convert(np.arange(4, dtype=np.uint8), orc.Float64)
    assert list(r) == [0.0, 1.0, 2.0, 3.0]


# --- src/masked/masked_buffer.rs:464-479 unary ------------------------------------------------------
def test_masked_unary(orc):
    buf = np.arange(9, dtype=np.uint8)
    mask = np.arange(9) % 2 == 0
    r = orc.neg(buf)
    assert r.dtype == np.int16
    v = orc.fill_nodata(r, mask, orc.Int16, orc.ND_DEFAULT)
    m = -32768
    assert list(v) == [0, m, -2, m, -4, m, -6, m, -8]


# --- src/masked/masked_buffer.rs:481-485 min_max ----------------------------------------------------
def test_masked_min_max(orc):
    buf = np.arange(9, dtype=np.uint8)
    mask = np.array([i != 0 and i != 8 for i in range(9)])
    mn, mx = orc.min_max(buf, mask)
    assert (mn.key(), mx.key()) == (orc.value(orc.UInt8, 1).key(), orc.value(orc.UInt8, 7).key())


# --- src/masked/masked_buffer.rs:487-509 scalar -----------------------------------------------------
def test_masked_scalar(orc):
    buf = np.arange(9, dtype=np.uint8)
    mask = np.arange(9) % 2 == 0
    r = orc.scalar(orc.MUL, buf, orc.value(orc.Float64, 2.0))
    fmin = np.finfo(np.float64).min
    v = orc.fill_nodata(r, mask, orc.Float64, orc.ND_VALUE, orc.value(orc.Float64, fmin))
    assert list(v) == [0.0, fmin, 4.0, fmin, 8.0, fmin, 12.0, fmin, 16.0]


# --- src/masked/masked_buffer.rs:511-531 binary -----------------------------------------------------
def test_masked_binary(orc):
    lhs, lmask = np.full(9, 1.0), np.arange(9) % 2 == 0
    rhs, rmask = np.full(9, 2.0), np.ones(9, bool)
    pyop = {orc.ADD: 3.0, orc.SUB: -1.0, orc.MUL: 2.0, orc.DIV: 0.5}
    for op, e in pyop.items():
        r = orc.binary(op, lhs, rhs)
        m = orc.mask_and(lmask, rmask)
        assert r[0] == e and m[0] and not m[1] and r[4] == e and not m[5]


# --- src/masked/masked_buffer.rs:15-38 doc-test / examples/masked.rs --------------------------------
def test_masked_example(orc):
    buf, mask = np.arange(4, dtype=np.float64), np.arange(4) % 2 == 0
    assert orc.mask_counts(mask) == (2, 2)
    ones = np.ones(4)
    r = orc.scalar(orc.MUL, orc.binary(orc.ADD, buf, ones), orc.value(orc.Float64, 2.0))
    rm = orc.mask_and(mask, np.ones(4, bool))
    assert list(r) == [2.0, 4.0, 6.0, 8.0] and list(rm) == [True, False, True, False]


# --- src/gdal/rasterband.rs:138-163 read_cells (NDVI on the Landsat fixtures) ------------------------
def test_ndvi_landsat(orc, landsat):
    red, nir = landsat["red"].ravel(), landsat["nir"].ravel()
    ndvi = orc.binary(orc.DIV, orc.binary(orc.SUB, nir, red), orc.binary(orc.ADD, nir, red))
    mn, mx = orc.min_max(ndvi)
    # the reference's own (one-sided) assertions
    assert mn.numpy() - -0.1248899911993 < 1e-8 and mx.numpy() - 0.66998345719859 < 1e-8
    # exact bits (recomputed from the TIFFs)
    assert float(mn.numpy()).hex() == "-0x1.ff8ca5bcc77dcp-4"
    assert float(mx.numpy()).hex() == "0x1.5708125b0ed28p-1"


# --- src/gdal/rasterband.rs:166-191 read_cells_masked ------------------------------------------------
def test_ndvi_landsat_masked(orc, landsat):
    red, nir = landsat["red"].ravel(), landsat["nir_nd"].ravel()
    nd = orc.value(orc.UInt16, int(landsat["gdal_nodata"][0]))
    rmask = orc.mask_from_nodata(red, orc.ND_VALUE, nd)
    nmask = orc.mask_from_nodata(nir, orc.ND_VALUE, nd)
    nir_counts = orc.mask_counts(nmask)
    assert nir_counts == (31430, 4)
    num, nmask2 = orc.binary(orc.SUB, nir, red), orc.mask_and(nmask, rmask)
    den, dmask2 = orc.binary(orc.ADD, nir, red), orc.mask_and(nmask, rmask)
    ndvi, mask = orc.binary(orc.DIV, num, den), orc.mask_and(nmask2, dmask2)
    assert orc.mask_counts(mask) == nir_counts
    mn, mx = orc.min_max(ndvi, mask)
    assert float(mn.numpy()).hex() == "-0x1.ff8ca5bcc77dcp-4"
    assert float(mx.numpy()).hex() == "0x1.5708125b0ed28p-1"
