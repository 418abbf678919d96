"""The reference's own known-answer tests, ported onto the Python mirror of its API so they read like
the originals (file:line under /root/reference) and run on the GPU through the C ABI."""
import numpy as np
import pytest

import erased_cells_b200 as ec
from erased_cells_b200 import CellBuffer, CellType, CellValue, Mask, MaskedCellBuffer, NoData

pytestmark = pytest.mark.gpu
u8 = CellType.UInt8


def filler(i): return i
def masker(i): return i % 2 == 0
def filler_masker(i): return (filler(i), masker(i))


def test_defaults():  # src/buffer.rs:469-480
    for ct in CellType:
        cv = CellBuffer.with_defaults(3, ct)
        assert cv.len() == 3
        assert cv.get(0) == ct.zero()


def test_put_get():  # src/buffer.rs:482-494
    for ct in CellType:
        cv = CellBuffer.fill(3, ct.zero())
        one = ct.one()
        cv.put(1, one)
        assert cv.get(1) == one.convert(ct)


def test_extend():  # src/buffer.rs:496-505
    buf = CellBuffer.fill(3, CellValue(u8, 0))
    assert not buf.is_empty()
    assert buf.cell_type() == u8
    buf.extend(np.array([1], dtype=np.uint8))
    assert buf.cell_type() == u8
    assert buf.get(0) == CellValue.new(0) and buf.get(3) == CellValue.new(1)


def test_to_vec():  # src/buffer.rs:507-519
    for ct in CellType:
        v = np.zeros(3, dtype=ct.dtype)
        assert np.array_equal(CellBuffer.from_vec(v).to_vec(ct), v)


def test_min_max():  # src/buffer.rs:515-526
    mn, mx = CellBuffer.from_vec(np.array([-1.0, 3.0, 2000.0, -5555.5])).min_max()
    assert mn == CellValue(CellType.Float64, -5555.5) and mx == CellValue(CellType.Float64, 2000.0)
    mn, mx = CellBuffer.from_vec(np.array([1, 3, 200, 0], dtype=np.uint8)).min_max()
    assert mn == CellValue(u8, 0) and mx == CellValue(u8, 200)


def test_from_others():  # src/buffer.rs:528-556
    b = CellBuffer.from_iter([CellValue(CellType.UInt16, x) for x in (3, 4, 5)])
    assert b.cell_type() == CellType.UInt16 and b.len() == 3 and b.get(2) == CellValue(CellType.UInt16, 5)
    b = CellBuffer.from_vec(np.array([33.3, 44.4, 55.5], dtype=np.float32))
    assert b.cell_type() == CellType.Float32 and b.len() == 3 and b.get(2) == CellValue(CellType.Float32, 55.5)


def test_debug():  # src/buffer.rs:558-564
    assert repr(CellBuffer.fill(5, CellValue.new(37))).startswith("Int32CellBuffer")
    assert "..." in repr(CellBuffer.fill(15, CellValue.new(37)))
    # Elided (src/lib.rs:198-206): "1, 1, 1" and "0, 0, 0, 0, 0, ... 0, 0, 0, 0, 0"; only the ends leave the device
    assert repr(CellBuffer.fill(3, CellValue(CellType.UInt8, 1))) == "UInt8CellBuffer(1, 1, 1)"
    assert repr(CellBuffer.from_vec(np.arange(30, dtype=np.uint16))) == "UInt16CellBuffer(0, 1, 2, 3, 4, ... 25, 26, 27, 28, 29)"
    assert repr(Mask.new([True, False] * 8)) == "Mask(true, false, true, false, true, ... false, true, false, true, false)"
    assert repr(Mask.new([True, False, True])) == "Mask(true, false, true)"


def test_convert():  # src/buffer.rs:566-578
    for ct in CellType:
        buf = CellBuffer.with_defaults(3, ct)
        for target in (t for t in CellType if ct.can_fit_into(t)):
            assert buf.convert(target).cell_type() == target


def test_unary():  # src/buffer.rs:580-592
    for ct in CellType:
        one = ct.one()
        buf = -CellBuffer.fill(3, one)
        assert buf.get(0) == -one


def test_binary():  # src/buffer.rs:595-614 — all 100 type pairs
    for lhs_ct in CellType:
        lhs_val = lhs_ct.one()
        for rhs_ct in CellType:
            lhs = CellBuffer.fill(3, lhs_val)
            rhs_val = rhs_ct.one() + rhs_ct.one()
            rhs = CellBuffer.fill(3, rhs_val)
            assert (lhs + rhs).get(0) == lhs_val + rhs_val
            assert (rhs + lhs).get(1) == rhs_val + lhs_val
            assert (lhs - rhs).get(2) == lhs_val - rhs_val
            assert (rhs - lhs).get(0) == rhs_val - lhs_val
            assert (lhs * rhs).get(1) == lhs_val * rhs_val
            assert (rhs * lhs).get(2) == rhs_val * lhs_val
            assert (lhs / rhs).get(0) == lhs_val / rhs_val
            assert (rhs / lhs).get(1) == rhs_val / lhs_val


def test_scalar():  # src/buffer.rs:617-621
    buf = CellBuffer.fill_via(9, lambda i: i + 1, u8)
    r = buf * 2.0
    assert r == CellBuffer.fill_via(9, lambda i: (i + 1.0) * 2.0, CellType.Float64)


def test_equal():  # src/buffer.rs:623-636
    buf = CellBuffer.fill_via(9, lambda i: np.nan if i % 2 == 0 else float(i), CellType.Float64)
    assert buf == buf
    assert CellBuffer.with_defaults(4, u8) == CellBuffer.with_defaults(4, u8)
    assert CellBuffer.with_defaults(4, u8) != CellBuffer.with_defaults(5, u8)


def test_cmp():  # src/buffer.rs:638-672
    i32 = lambda *a: CellBuffer.from_vec(np.array(a, dtype=np.int32))
    assert i32(1, 2, 3) < i32(2, 3, 4)
    assert i32(1, 2, 3) < i32(2, 3)
    f = lambda *a: CellBuffer.from_vec(np.array(a, dtype=np.float64))
    assert f(np.nan, 2.0, 3.0) < f(np.nan, 2.0, 4.0)
    assert CellBuffer.with_defaults(4, u8) < CellBuffer.with_defaults(4, CellType.Float32)
    assert CellBuffer.with_defaults(4, CellType.Float32) > CellBuffer.with_defaults(4, u8)
    assert CellBuffer.with_defaults(4, u8) < CellBuffer.with_defaults(5, u8)
    assert CellBuffer.with_defaults(5, CellType.Float64) > CellBuffer.with_defaults(4, CellType.Float64)


def test_quick_example():  # README.md:22-33, examples/quick.rs
    buf1 = CellBuffer.from_vec(np.array([1, 2, 3], dtype=np.uint8))
    buf2 = CellBuffer.from_vec(np.array([2, 4, 6], dtype=np.uint16))
    result = buf1 / buf2 * 0.5
    assert result == CellBuffer.from_vec(np.array([0.25, 0.25, 0.25]))


def test_buffer_example():  # examples/buffer.rs
    buf1 = CellBuffer.fill_via(9, lambda i: i, u8)
    assert buf1.cell_type() == u8 and buf1.get(3) == CellValue(u8, 3)
    mn, mx = buf1.min_max()
    assert (mn, mx) == (CellValue(u8, 0), CellValue(u8, 8))
    assert ((mx - mn + 1) / 2) == CellValue.new(4.5)
    buf2 = CellBuffer.fill_via(9, lambda i: 8.0 - i, CellType.Float32)
    assert buf2.min_max() == (CellValue(CellType.Float32, 0.0), CellValue(CellType.Float32, 8.0))
    diff = buf2 - buf1
    assert diff.min_max() == (CellValue.new(-8), CellValue.new(8))


def test_mask():  # src/masked/mask.rs:183-242
    assert Mask.fill(3, True).counts() == (3, 0) and Mask.fill(3, False).counts() == (0, 3)
    assert Mask.fill_via(3, lambda i: i % 2 == 0).counts() == (2, 1)
    m = Mask.fill(3, True)
    m.put(1, False)
    m[0] = False
    assert m == Mask.new([False, False, True])
    t, f = Mask.fill(4, True), Mask.fill(4, False)
    assert ~t == f
    assert ~Mask.new([True, False, True, False]) == Mask.new([False, True, False, True])
    m = Mask.fill_via(4, lambda i: i % 2 == 0)
    assert not m.all(True) and not m.all(False)
    assert Mask.fill(4, True).all(True) and not Mask.fill(4, True).all(False)
    l, r = Mask.fill_via(4, lambda i: i % 2 == 0), Mask.fill_via(4, lambda i: i % 2 != 0)
    assert (l & r).all(False) and (l | r).all(True)


def test_masked_ctor():  # src/masked/masked_buffer.rs:400-410
    m = MaskedCellBuffer.fill_via(3, filler, u8)
    r = MaskedCellBuffer.new(CellBuffer.fill_via(3, filler, u8), Mask.fill(3, True))
    assert m == r
    assert MaskedCellBuffer.from_vec(np.zeros(4)).mask().counts() == (4, 0)
    assert MaskedCellBuffer.with_defaults(4, CellType.Int16).mask().counts() == (4, 0)


def test_vec_with_nodata():  # src/masked/masked_buffer.rs:412-425
    v = np.array([1.0, np.nan, 3.0, np.nan])
    m = MaskedCellBuffer.from_vec_with_nodata(v, NoData.default(CellType.Float64))
    assert m == MaskedCellBuffer.new(CellBuffer.from_vec(v), Mask.new([True, False, True, False]))
    m = MaskedCellBuffer.from_vec_with_nodata(v, NoData.new(CellType.Float64, 3.0))
    assert m == MaskedCellBuffer.new(CellBuffer.from_vec(v), Mask.new([True, True, False, True]))


def test_get_masked():  # src/masked/masked_buffer.rs:427-440
    buf = MaskedCellBuffer.fill_with_mask_via(9, filler_masker, u8)
    assert buf.get(4) == CellValue.new(4)
    assert buf.get_masked(4) == CellValue.new(4) and buf.get_masked(5) is None
    buf.put(5, CellValue(u8, 4))
    assert buf.get_masked(5) is None
    buf.mask_mut().put(5, True)
    assert buf.get_masked(5) == CellValue.new(4)
    buf.put_with_mask(5, CellValue(u8, 99), False)
    assert buf.get_masked(5) is None


def test_masked_extend_and_from_iter():  # src/masked/masked_buffer.rs:449-462
    buf = MaskedCellBuffer.fill(3, CellValue.new(0))
    buf.extend([(1, False)])
    assert buf.get_masked(0) == CellValue.new(0) and buf.get_masked(3) is None
    buf = MaskedCellBuffer.from_vec(np.arange(5, dtype=np.int16))
    assert buf.mask().all(True) and list(buf.to_vec(CellType.Int16)) == [0, 1, 2, 3, 4]
    buf = MaskedCellBuffer.from_iter(np.arange(5, dtype=np.int16))            # `(0..5i16).collect()`, :458-462
    assert buf.cell_type() == CellType.Int16 and buf.mask().all(True) and list(buf.to_vec(CellType.Int16)) == [0, 1, 2, 3, 4]
    buf = MaskedCellBuffer.from_iter((np.uint8(i), i % 2 == 0) for i in range(5))  # FromIterator<(C, bool)>, :263-278
    assert buf.cell_type() == CellType.UInt8 and buf.counts() == (3, 2) and buf.get_masked(1) is None and buf.get_masked(4) == CellValue.new(np.uint8(4))
    assert MaskedCellBuffer.from_iter([], CellType.Float32).cell_type() == CellType.Float32


def test_masked_convert():  # src/masked/masked_buffer.rs:442-447
    buf = MaskedCellBuffer.fill_with_mask_via(4, filler_masker, u8)
    r = buf.convert(CellType.Float64)
    assert list(r.to_vec(CellType.Float64)) == [0.0, 1.0, 2.0, 3.0]


def test_masked_unary():  # src/masked/masked_buffer.rs:464-479
    mbuf = MaskedCellBuffer.fill_with_mask_via(9, filler_masker, u8)
    v = (-mbuf).to_vec_with_nodata(NoData.default(CellType.Int16))
    m = -32768
    assert list(v) == [0, m, -2, m, -4, m, -6, m, -8] and v.dtype == np.int16


def test_masked_min_max():  # src/masked/masked_buffer.rs:481-485
    mbuf = MaskedCellBuffer.fill_with_mask_via(9, lambda i: (filler(i), i != 0 and i != 8), u8)
    assert mbuf.min_max() == (CellValue(u8, 1), CellValue(u8, 7))


def test_masked_scalar():  # src/masked/masked_buffer.rs:487-509
    mbuf = MaskedCellBuffer.fill_with_mask_via(9, lambda i: (filler(i), True), u8)
    r = mbuf * 2.0
    expected = CellBuffer.fill_via(9, filler, u8) * 2.0
    assert r == expected
    mbuf = MaskedCellBuffer.fill_with_mask_via(9, filler_masker, u8)
    r = mbuf * 2.0
    assert r != expected
    fmin = np.finfo(np.float64).min
    v = r.to_vec_with_nodata(NoData.new(CellType.Float64, fmin))
    assert list(v) == [0.0, fmin, 4.0, fmin, 8.0, fmin, 12.0, fmin, 16.0]


def test_masked_binary():  # src/masked/masked_buffer.rs:511-531
    lhs = MaskedCellBuffer.new(CellBuffer.fill(9, CellValue.new(1.0)), Mask.fill_via(9, masker))
    rhs = MaskedCellBuffer.new(CellBuffer.fill(9, CellValue.new(2.0)), Mask.fill(9, True))
    for op, e in enumerate((3.0, -1.0, 2.0, 0.5)):
        r = lhs._bin(op, rhs)
        assert r.get_masked(0) == CellValue.new(e) and r.get_masked(1) is None
        assert r.get_masked(4) == CellValue.new(e) and r.get_masked(5) is None


def test_masked_example():  # src/masked/masked_buffer.rs:15-38, examples/masked.rs
    buf = MaskedCellBuffer.fill_with_mask_via(4, lambda i: (float(i), i % 2 == 0), CellType.Float64)
    assert buf.mask() == Mask.new([True, False, True, False]) and buf.counts() == (2, 2)
    ones = MaskedCellBuffer.from_vec(np.ones(4))
    r = (buf + ones) * 2.0
    expected = MaskedCellBuffer.new(CellBuffer.from_vec(np.array([2.0, 4.0, 6.0, 8.0])), Mask.new([True, False, True, False]))
    assert r == expected


def test_ndvi_landsat(landsat):  # src/gdal/rasterband.rs:138-163
    red, nir = CellBuffer.from_vec(landsat["red"].ravel()), CellBuffer.from_vec(landsat["nir"].ravel())
    ndvi = (nir - red) / (nir + red)
    mn, mx = ndvi.min_max()
    assert mn.to_f64() - -0.1248899911993 < 1e-8 and mx.to_f64() - 0.66998345719859 < 1e-8
    assert float(mn.value()).hex() == "-0x1.ff8ca5bcc77dcp-4" and float(mx.value()).hex() == "0x1.5708125b0ed28p-1"
    assert nir.normalized_difference(red) == ndvi


def test_ndvi_landsat_masked(landsat):  # src/gdal/rasterband.rs:166-191
    nd = NoData.new(CellType.UInt16, int(landsat["gdal_nodata"][0]))
    red = MaskedCellBuffer.from_vec_with_nodata(landsat["red"].ravel(), nd)
    nir = MaskedCellBuffer.from_vec_with_nodata(landsat["nir_nd"].ravel(), nd)
    nir_data, nir_nodata = nir.counts()
    ndvi = (nir - red) / (nir + red)
    assert ndvi.counts() == (nir_data, nir_nodata) == (31430, 4)
    mn, mx = ndvi.min_max()
    assert float(mn.value()).hex() == "-0x1.ff8ca5bcc77dcp-4" and float(mx.value()).hex() == "0x1.5708125b0ed28p-1"
