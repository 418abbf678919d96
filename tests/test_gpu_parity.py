"""Parity of the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.
Bit-exact for everything: integer, cast, mask and f64 work (NaN results follow the x86 rule the
reference's platform produces). Run on the B200 box: pytest -m gpu."""
import ctypes as C

import numpy as np
import pytest

import erased_cells_b200 as ec
from erased_cells_b200 import CellBuffer, CellType, CellValue, Mask, MaskedCellBuffer, NoData, synth

pytestmark = pytest.mark.gpu
CT = list(CellType)
# ragged sizes around the tile boundaries (tiles are 256*V*4 cells), plus tiny and empty
SIZES = [0, 1, 3, 31, 33, 1000, 4096, 4097, 16384 + 5, 3 * 32768 + 77]


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view({1: "u1", 2: "u2", 4: "u4", 8: "u8"}[a.dtype.itemsize])


def cells(ct, n, seed):
    """full-bit-range cells with specials sprinkled in (NaN payloads, infs, zeros, MIN/MAX)"""
    a = synth.host(ct, n, seed)
    dt = CellType(ct).dtype
    if n >= 16:
        if dt.kind == "f":
            sp = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, -np.nan, 1.0, -1.0], dtype=dt)
        else:
            info = np.iinfo(dt)
            sp = np.array([info.min, info.max, 0, 1, 0, 0, info.max, info.min], dtype=dt)
        a = a.copy()
        a[:: max(n // 8, 1)][:8] = sp[: len(a[:: max(n // 8, 1)][:8])]
    return a


@pytest.mark.parametrize("lct", CT)
def test_binary_all_pairs(orc, lct):
    for rct in CT:
        for n in (0, 5, 4097, 2 * 32768 + 19):
            l, r = cells(lct, n, 0x100 + int(lct)), cells(rct, n + (3 if n else 0), 0x200 + int(rct))  # zip truncates
            dl, dr = CellBuffer.from_vec(l), CellBuffer.from_vec(r)
            for op in range(4):
                got = dl._bin(op, dr)
                want = orc.tight_binary(op, l, r)
                assert got.cell_type() == (CellType.Float64 if n else CellType.UInt8)
                assert got.len() == n
                assert np.array_equal(bits(got.to_vec()), bits(want)), (lct, rct, op, n)


def test_binary_matches_faithful_oracle_small(orc):
    # the per-cell tagged flavour (what the reference executes), small sizes
    for lct in CT:
        for rct in CT:
            l, r = cells(lct, 300, 7 + int(lct)), cells(rct, 300, 70 + int(rct))
            for op in range(4):
                got = CellBuffer.from_vec(l)._bin(op, CellBuffer.from_vec(r)).to_vec()
                assert np.array_equal(bits(got), bits(orc.binary(op, l, r))), (lct, rct, op)


@pytest.mark.parametrize("lct", CT)
def test_scalar_all_pairs(orc, lct):
    for n in (0, 7, 8192 + 3):
        l = cells(lct, n, 0x300 + int(lct))
        dl = CellBuffer.from_vec(l)
        for rct in CT:
            for s in cells(rct, 16, 0x400 + int(rct))[:6]:
                for op in range(4):
                    got = dl._bin(op, CellValue(rct, s))
                    want = orc.tight_scalar(op, l, orc.value(int(rct), s))
                    assert got.cell_type() == (CellType.Float64 if n else CellType.UInt8)
                    assert np.array_equal(bits(got.to_vec()), bits(want)), (lct, rct, op, n, s)
    # python literals: int -> Int32, float -> Float64 (Rust literal defaults)
    assert np.array_equal((CellBuffer.from_vec(np.arange(9, dtype=np.uint8)) * 2.0).to_vec(), np.arange(9) * 2.0)


def test_neg_all_types(orc):
    for ct in CT:
        for n in SIZES:
            a = cells(ct, n, 0x500 + int(ct))
            got = -CellBuffer.from_vec(a)
            want = orc.neg(a)
            assert got.cell_type().dtype == want.dtype and got.len() == n
            assert np.array_equal(bits(got.to_vec()), bits(want)), (ct, n)


def test_convert_all_pairs(orc):
    for s in CT:
        for n in (0, 9, 4097, 32768 + 1):
            a = cells(s, n, 0x600 + int(s))
            da = CellBuffer.from_vec(a)
            for d in CT:
                if not s.can_fit_into(d):
                    with pytest.raises(ec.NarrowingError) as e:
                        da.convert(d)
                    assert (e.value.src, e.value.dst) == (int(s), int(d))
                    assert "Invalid narrowing from cell-type" in str(e.value)
                    continue
                got = da.convert(d)
                want = orc.tight_convert(a, int(d))
                assert got.cell_type().dtype == want.dtype, (s, d, n)  # empty non-identity => UInt8
                assert np.array_equal(bits(got.to_vec()), bits(want)), (s, d, n)
                if n:
                    assert np.array_equal(bits(da.to_vec(d)), bits(want))


def test_min_max_all_types(orc):
    for ct in CT:
        for n in SIZES + [(1 << 20) + 13]:
            a = cells(ct, n, 0x700 + int(ct))
            da = CellBuffer.from_vec(a)
            mn, mx = da.min_max()
            omn, omx = orc.tight_min_max(a)
            assert (int(mn.cell_type()), mn.bits, mx.bits) == (int(ct), omn.bits, omx.bits), (ct, n)
            m = synth.host(CellType.UInt8, n, 0x800 + n) < 100
            mmn, mmx = MaskedCellBuffer(da, Mask.new(m)).min_max()
            omn, omx = orc.tight_min_max(a, m)
            assert (mmn.bits, mmx.bits) == (omn.bits, omx.bits), (ct, n, "masked")
    # total order + seeds
    mn, mx = CellBuffer.from_vec(np.array([np.inf], np.float32)).min_max()
    assert mn.value() == np.finfo(np.float32).max and mx.value() == np.inf
    mn, mx = CellBuffer.from_vec(np.array([0.0, -0.0])).min_max()
    assert (mn.bits, mx.bits) == (0x8000000000000000, 0)
    mn, mx = MaskedCellBuffer(CellBuffer.from_vec(np.arange(5, dtype=np.int16)), Mask.fill(5, False)).min_max()
    assert (mn.value(), mx.value()) == (32767, -32768)


def test_min_max_finite_values_small_faithful(orc):
    for ct in CT:
        a = cells(ct, 999, 0x900 + int(ct))
        mn, mx = CellBuffer.from_vec(a).min_max()
        omn, omx = orc.min_max(a)
        assert (mn.bits, mx.bits) == (omn.bits, omx.bits)


def test_mask_ops(orc):
    for n in SIZES:
        a = synth.host(CellType.UInt8, n, 0xA00 + n) < 128
        b = synth.host(CellType.UInt8, n + 5, 0xB00 + n) < 64
        ma, mb = Mask.new(a), Mask.new(b)
        assert ma.len() == n and np.array_equal(ma.to_vec(), a)
        assert np.array_equal((~ma).to_vec(), orc.mask_not(a))
        assert np.array_equal((ma & mb).to_vec(), orc.mask_and(a, b))
        assert np.array_equal((ma | mb).to_vec(), orc.mask_or(a, b))
        assert np.array_equal((mb & ma).to_vec(), orc.mask_and(b, a))
        assert ma.counts() == orc.mask_counts(a)
        assert (~ma).counts() == orc.mask_counts(orc.mask_not(a))  # tail bits stay clear
        assert (mb | ma).counts() == orc.mask_counts(orc.mask_or(b, a))
        for v in (True, False):
            assert ma.all(v) == orc.mask_all(a, v)
        assert Mask.fill(n, True).counts() == (n, 0) and Mask.fill(n, False).counts() == (0, n)
        assert ma == Mask.new(a) and (n == 0 or ma != ~ma)
    m = Mask.fill(70, True)
    m.put(1, False)
    m[69] = False
    assert not m.get(1) and not m[69] and m[2] and m.counts() == (68, 2)
    with pytest.raises(IndexError):
        m.get(70)
    m.extend([True, False, True])
    assert m.len() == 73 and list(m.to_vec()[-3:]) == [True, False, True]
    # derive(Ord) on Vec<bool>: lexicographic, false < true, then length
    assert Mask.new([False, True]) < Mask.new([True, False]) and Mask.new([True]) < Mask.new([True, False])


def test_mask_counts_follow_mutation(orc):
    """Mask::counts (src/masked/mask.rs:72-80) comes with the mask: the kernel that wrote it counted its set bits. The
    cached count must follow every way a mask can change (put, extend, copy on write of shared words) and never be
    served for different bits; masked scalar / neg results share the operand's mask words instead of copying them."""
    L = ec.lib()
    n = 3 * 32768 + 41
    a = synth.host(CellType.Int16, n, 0xC01, kind=synth.INT_RANGE, lo=-32768, hi=32767, period=20, sentinel=-32768)
    b = synth.host(CellType.Int16, n, 0xC02, kind=synth.INT_RANGE, lo=-32768, hi=32767, period=30, sentinel=-32768)
    nd = NoData.default(CellType.Int16)
    ma, mb = MaskedCellBuffer.from_vec_with_nodata(a, nd), MaskedCellBuffer.from_vec_with_nodata(b, nd)
    wa, wb = orc.mask_from_nodata(a, orc.ND_DEFAULT), orc.mask_from_nodata(b, orc.ND_DEFAULT)
    k0 = L.ec_kernel_launches()
    assert ma.counts() == orc.mask_counts(wa) and mb.counts() == orc.mask_counts(wb)
    r = ma - mb
    assert r.counts() == orc.mask_counts(orc.mask_and(wa, wb))
    assert (ma.mask() | mb.mask()).counts() == orc.mask_counts(orc.mask_or(wa, wb))
    assert (~ma.mask()).counts() == orc.mask_counts(orc.mask_not(wa))
    assert ma.mask().slice(32768, 40000).counts() == orc.mask_counts(wa[32768:72768])
    # none of those counts needed a popcount launch: 1 masked binary + 1 or + 1 not + 1 slice
    assert L.ec_kernel_launches() == k0 + 4, L.ec_kernel_launches() - k0
    # masked * scalar and -masked: the mask words are shared, not copied
    s, g = r * 0.5, -ma
    assert L.ec_mask_device_words(s.mask()._h) == L.ec_mask_device_words(r.mask()._h)
    assert L.ec_mask_device_words(g.mask()._h) == L.ec_mask_device_words(ma.mask()._h)
    assert s.counts() == r.counts() and g.counts() == ma.counts()
    # put on one of the sharers copies first; the cached counts follow the mutation, the other side keeps its bits
    before = r.counts()
    i = int(np.flatnonzero(orc.mask_and(wa, wb))[7])
    s.mask_mut().put(i, False)
    assert s.counts() == (before[0] - 1, before[1] + 1) and r.counts() == before
    assert L.ec_mask_device_words(s.mask()._h) != L.ec_mask_device_words(r.mask()._h)
    assert r.mask().get(i) and not s.mask().get(i)
    s.mask_mut().put(i, False)  # no change: count stays
    assert s.counts() == (before[0] - 1, before[1] + 1)
    s.mask_mut().put(i, True)
    assert s.counts() == before and s.mask() == r.mask()
    m = Mask.new(wa)
    m.extend([True, True, False])
    assert m.counts() == (orc.mask_counts(wa)[0] + 2, orc.mask_counts(wa)[1] + 1)
    assert m.counts() == orc.mask_counts(m.to_vec())


def test_from_nodata_and_fill_nodata(orc):
    for ct in CT:
        for n in (0, 5, 4096 + 33, 65536 + 7):
            a = cells(ct, n, 0xC00 + int(ct))
            if n > 20:  # make sure sentinels occur
                a[3::17] = a[3]
            da = CellBuffer.from_vec(a)
            kinds = [(NoData.default(ct), (orc.ND_DEFAULT, None)), (NoData.none(ct), (orc.ND_NONE, None))]
            if n:
                kinds.append((NoData.new(ct, a[3 % n]), (orc.ND_VALUE, orc.value(int(ct), a[3 % n]))))
            for nd, (ok, ov) in kinds:
                m = MaskedCellBuffer.from_buffer_with_nodata(da, nd)
                want = orc.mask_from_nodata(a, ok, ov)
                assert np.array_equal(m.mask().to_vec(), want), (ct, n, ok)
                assert m.counts() == orc.mask_counts(want)
            # to_vec_with_nodata over every legal target type
            mask = synth.host(CellType.UInt8, n, 0xD00 + n) < 200
            mb = MaskedCellBuffer(da, Mask.new(mask))
            for d in CT:
                fills = [NoData.default(d), NoData.none(d), NoData.new(d, 7)]
                if not ct.can_fit_into(d):
                    with pytest.raises(ec.NarrowingError):
                        mb.to_vec_with_nodata(fills[0])
                    continue
                for nd, (ok, ov) in zip(fills, [(orc.ND_DEFAULT, None), (orc.ND_NONE, None), (orc.ND_VALUE, orc.value(int(d), 7))]):
                    got = mb.to_vec_with_nodata(nd)
                    want = orc.fill_nodata(a, mask, int(d), ok, ov)
                    assert got.dtype == want.dtype and np.array_equal(bits(got), bits(want)), (ct, d, n, ok)
    # bitwise sentinel semantics: only the canonical +qNaN is the float default; Value(0.0) != -0.0
    a = np.array([0x7FF8000000000000, 0xFFF8000000000000, 0x7FF8000000000001], dtype=np.uint64).view(np.float64)
    assert list(MaskedCellBuffer.from_vec_with_nodata(a, NoData.default(CellType.Float64)).mask().to_vec()) == [False, True, True]
    z = np.array([0.0, -0.0])
    assert list(MaskedCellBuffer.from_vec_with_nodata(z, NoData.new(CellType.Float64, 0.0)).mask().to_vec()) == [False, True]


def test_masked_binary_and_chain(orc):
    for (lct, rct) in [(CellType.Int16, CellType.Int16), (CellType.UInt8, CellType.Float32), (CellType.Float64, CellType.UInt64)]:
        for n, extra in ((0, 0), (77, 0), (4096 * 3 + 5, 0), (8192 + 40, 9)):
            l, r = cells(lct, n, 0xE00), cells(rct, n + extra, 0xE01)
            lm = synth.host(CellType.UInt8, n, 0xE02) < 200
            rm = synth.host(CellType.UInt8, n + extra, 0xE03) < 220
            ml = MaskedCellBuffer(CellBuffer.from_vec(l), Mask.new(lm))
            mr = MaskedCellBuffer(CellBuffer.from_vec(r), Mask.new(rm))
            for op in range(4):
                got = ml._bin(op, mr)
                assert np.array_equal(bits(got.to_vec()), bits(orc.tight_binary(op, l, r))), (lct, rct, op, n)
                assert np.array_equal(got.mask().to_vec(), orc.mask_and(lm, rm))
                assert got.counts() == orc.mask_counts(orc.mask_and(lm, rm))
            s = ml * 0.0001
            assert np.array_equal(s.mask().to_vec(), lm)
            assert np.array_equal(bits(s.to_vec()), bits(orc.tight_scalar(orc.MUL, l, orc.value(orc.Float64, 0.0001))))
            ng = -ml
            assert np.array_equal(bits(ng.to_vec()), bits(orc.neg(l))) and np.array_equal(ng.mask().to_vec(), lm)
    with pytest.raises(AssertionError):
        MaskedCellBuffer(CellBuffer.with_defaults(4, CellType.UInt8), Mask.fill(5, True))


def test_fused_chains_equal_unfused(orc):
    for (lct, rct) in [(CellType.UInt16, CellType.UInt16), (CellType.UInt8, CellType.UInt16), (CellType.Float32, CellType.Int64),
                       (CellType.Float64, CellType.Float64), (CellType.Int8, CellType.UInt64)]:
        for n in (0, 100, 32768 + 3):
            l, r = cells(lct, n, 0xF00), cells(rct, n, 0xF01)
            if n > 50 and lct == rct:
                r[5::7] = l[5::7]  # (a-b)/(a+b) with a == b, and 0/0 where both are 0
                l[10::50] = 0
                r[10::50] = 0
            dl, dr = CellBuffer.from_vec(l), CellBuffer.from_vec(r)
            fused = dl.normalized_difference(dr)
            unfused = (dl - dr) / (dl + dr)
            want = orc.tight_binary(orc.DIV, orc.tight_binary(orc.SUB, l, r), orc.tight_binary(orc.ADD, l, r))
            assert fused == unfused and np.array_equal(bits(fused.to_vec()), bits(want)), (lct, rct, n)
            for op1 in range(4):
                for op2 in range(4):
                    f = dl.binary_scalar(op1, dr, op2, 0.5)
                    w = orc.tight_scalar(op2, orc.tight_binary(op1, l, r), orc.value(orc.Float64, 0.5))
                    assert np.array_equal(bits(f.to_vec()), bits(w)), (lct, rct, op1, op2, n)


def test_buffer_cmp(orc):
    for ct in CT:
        a = cells(ct, 5000, 0x111 + int(ct))
        da = CellBuffer.from_vec(a)
        assert da == CellBuffer.from_vec(a) and da.cmp(da.clone()) == 0
        for pos in (0, 1, 2500, 4999):
            b = a.copy()
            bits(b)[pos] ^= 1
            assert da.cmp(CellBuffer.from_vec(b)) == orc.buffer_cmp(a, b) != 0
            assert CellBuffer.from_vec(b).cmp(da) == orc.buffer_cmp(b, a)
        assert da.cmp(CellBuffer.from_vec(a[:-1])) == 1 and CellBuffer.from_vec(a[:-1]).cmp(da) == -1
    # cell type dominates (src/buffer.rs:395-398)
    assert CellBuffer.with_defaults(4, CellType.UInt8) < CellBuffer.with_defaults(4, CellType.Float32)


def test_get_put_fill_extend(orc):
    for ct in CT:
        b = CellBuffer.fill(1000 + int(ct), CellValue(ct, 3))
        assert b.cell_type() == ct and b.len() == 1000 + int(ct) and np.all(b.to_vec() == 3)
        b.put(7, ct.one())
        assert b.get(7) == ct.one() and b.get(8) == CellValue(ct, 3)
        with pytest.raises(IndexError):
            b.get(b.len())
        if ct != CellType.Float64:
            with pytest.raises(ec.NarrowingError):
                b.put(0, CellValue(CellType.Float64, 1.0))
        z = CellBuffer.with_defaults(33, ct)
        assert np.all(z.to_vec() == 0) and z.get(0) == ct.zero()
    b = CellBuffer.fill(3, CellValue(CellType.UInt16, 0))
    b.extend(np.array([1, 2], dtype=np.uint8))
    assert b.cell_type() == CellType.UInt16 and list(b.to_vec()) == [0, 0, 0, 1, 2]


def test_synth_host_equals_device():
    for ct in CT:
        for kind, lo, hi in ((synth.FULL_BITS, 0, 0), (synth.INT_RANGE, 0, 100), (synth.REAL_RANGE, -1e4, 1e4)):
            if kind == synth.INT_RANGE and ct.dtype.kind != "f" and np.iinfo(ct.dtype).min < 0:
                lo = -50
            h = synth.host(ct, 10007, 0xEC40, 123, kind, lo, hi, 50, 1)
            d = synth.device(ct, 10007, 0xEC40, 123, kind, lo, hi, 50, 1).to_vec()
            assert np.array_equal(bits(h), bits(d)), (ct, kind)


def test_async_upload_orders_readers(orc):
    """from_vec(wait=False): the H2D copy runs on the upload stream; every consumer is ordered after it."""
    n = (1 << 24) + 5
    hs = [synth.host(CellType.UInt16, n, 0x5150 + i) for i in range(4)]
    bufs = [CellBuffer.from_vec(h, wait=False) for h in hs]          # four uploads in flight
    outs = [b.convert(CellType.Float64) for b in bufs]                # consumers queued immediately
    mms = [b.min_max() for b in bufs]
    for h, o, mm in zip(hs, outs, mms):
        assert np.array_equal(o.to_vec(), h.astype(np.float64))
        omn, omx = orc.tight_min_max(h)
        assert (mm[0].bits, mm[1].bits) == (omn.bits, omx.bits)
    b = CellBuffer.from_vec(hs[0], wait=False)
    del b                                                             # freeing with the copy in flight is safe
    c = CellBuffer.from_vec(hs[1], wait=False).wait()
    assert c == bufs[1]


def test_extend_value_checked(orc):
    """Extend<C> (src/buffer.rs:205-221): `to_<p>().unwrap()` is VALUE-checked — in-range values of a wider type
    are accepted, anything else is the reference's panic (NarrowingError here), and the buffer stays untouched."""
    for dct in CT:
        for sct in CT:
            base = cells(dct, 40, 0xABC + int(dct))
            # small values fit every target
            more = np.arange(0, 100, 7).astype(sct.dtype)
            b = CellBuffer.from_vec(base)
            b.extend(more)
            want = np.concatenate([base, orc.checked_cast(more, int(dct))])
            assert b.cell_type() == dct and np.array_equal(bits(b.to_vec()), bits(want)), (sct, dct)
            # full-range values: either every cell fits (then bit-exact) or the call is refused
            wild = cells(sct, 64, 0xDEF + int(sct))
            b = CellBuffer.from_vec(base)
            try:
                w = orc.checked_cast(wild, int(dct))
            except orc.NarrowingError:
                with pytest.raises(ec.NarrowingError):
                    b.extend(wild)
                assert b.len() == 40 and np.array_equal(bits(b.to_vec()), bits(base))
            else:
                b.extend(wild)
                assert np.array_equal(bits(b.to_vec()), bits(np.concatenate([base, w]))), (sct, dct)
    # src/buffer.rs:496-505: `buf.extend([1])` — an i32 literal into a u8 buffer
    buf = CellBuffer.fill(3, CellValue(CellType.UInt8, 0))
    buf.extend(np.array([1], dtype=np.int32))
    assert buf.cell_type() == CellType.UInt8 and list(buf.to_vec()) == [0, 0, 0, 1]
    # floats truncate toward zero when in range, NaN never fits an integer
    b = CellBuffer.from_vec(np.zeros(1, np.uint8))
    b.extend(np.array([1.9, 254.99, -0.5]))
    assert list(b.to_vec()) == [0, 1, 254, 0]
    with pytest.raises(ec.NarrowingError):
        b.extend(np.array([np.nan]))


def test_lazy_chains_fuse_and_match_eager(orc):
    """ec.lazy(): operators defer, the first access evaluates — (a-b)/(a+b) and (a op b) op s as ONE kernel each,
    bit-identical to eager evaluation; operands are snapshots (drop / mutate them freely afterwards)."""
    L = ec.lib()
    for (lct, rct) in [(CellType.UInt16, CellType.UInt16), (CellType.UInt8, CellType.UInt16), (CellType.Float32, CellType.Int64), (CellType.Float64, CellType.Float64)]:
        n = 3 * 32768 + 11
        l, r = cells(lct, n, 0x1A2), cells(rct, n, 0x1A3)
        if lct == rct:
            r[5::7] = l[5::7]
        a, b = CellBuffer.from_vec(l), CellBuffer.from_vec(r)
        eager_ndvi = (a - b) / (a + b)
        eager_chain = a / b * 0.5
        with ec.lazy():
            k0 = L.ec_kernel_launches()
            ndvi = (a - b) / (a + b)
            chain = a / b * 0.5
            deep = ((a + b) * 2.0 - a) / b                      # no special shape: evaluated op by op
            assert L.ec_kernel_launches() == k0                 # nothing has run yet
            assert ndvi.cell_type() == CellType.Float64 and ndvi.len() == n
            got = ndvi.to_vec()
            assert L.ec_kernel_launches() == k0 + 1 and L.ec_last_kernel() == b"normalized_difference(lazy)"
            assert np.array_equal(bits(got), bits(eager_ndvi.to_vec()))
            got = chain.to_vec()
            assert L.ec_kernel_launches() == k0 + 2 and L.ec_last_kernel() == b"binary_scalar(lazy)"
            assert np.array_equal(bits(got), bits(eager_chain.to_vec()))
            want = orc.tight_binary(orc.DIV, orc.tight_binary(orc.SUB, orc.tight_scalar(orc.MUL, orc.tight_binary(orc.ADD, l, r), orc.value(orc.Float64, 2.0)), l), r)
            assert np.array_equal(bits(deep.to_vec()), bits(want))
            # scale-and-offset: (x op1 s1) op2 s2 in one pass, for every op pair
            for op1 in range(4):
                for op2 in range(4):
                    k1 = L.ec_kernel_launches()
                    so = a._bin(op1, 0.0001)._bin(op2, 273.15)
                    got = so.to_vec()
                    assert L.ec_kernel_launches() == k1 + 1 and L.ec_last_kernel() == b"scalar_scalar(lazy)"
                    w2 = orc.tight_scalar(op2, orc.tight_scalar(op1, l, orc.value(orc.Float64, 0.0001)), orc.value(orc.Float64, 273.15))
                    assert np.array_equal(bits(got), bits(w2)), (lct, op1, op2)
            # operands are snapshots: dropping or mutating them after the fact does not change a pending result
            x, y = CellBuffer.from_vec(l), CellBuffer.from_vec(r)
            pend = (x - y) / (x + y)
            x.put(0, x.get(1))
            y.extend(r[:3])
            del y
            assert pend == eager_ndvi and x.get(0) == x.get(1)
            # a shared sub-expression used twice is evaluated once and stays correct
            num = a - b
            both = (num / (a + b), num * 3.0)
            assert both[0] == eager_ndvi and np.array_equal(bits(both[1].to_vec()), bits(orc.tight_scalar(orc.MUL, orc.tight_binary(orc.SUB, l, r), orc.value(orc.Float64, 3.0))))
            assert np.array_equal(bits(num.to_vec()), bits(orc.tight_binary(orc.SUB, l, r)))
            # reductions and masked ops see through pending values
            mm = ((a - b) / (a + b)).min_max()
            assert (mm[0].bits, mm[1].bits) == tuple(v.bits for v in eager_ndvi.min_max())
            ma, mb = MaskedCellBuffer.from_vec(l), MaskedCellBuffer.from_vec(r)
            mr = (ma - mb) / (ma + mb)
            assert mr.buffer() == eager_ndvi and mr.counts() == (n, 0)
            assert (CellBuffer.from_vec(l[:0]) + a).cell_type() == CellType.UInt8   # empty stays the reference's UInt8([])
        assert not L.ec_get_lazy()


def test_concurrent_host_threads(orc):
    """The C ABI is callable from any thread (SURVEY.md §8b threading): four host threads issue ops, reductions and
    comparisons at the same time on shared read-only inputs and private outputs; every result is still bit-exact."""
    import threading
    n = (1 << 20) + 77
    a_h, b_h = synth.host(CellType.UInt16, n, 0x7A1), synth.host(CellType.Float32, n, 0x7A2)
    a, b = CellBuffer.from_vec(a_h), CellBuffer.from_vec(b_h)
    want = {op: orc.tight_binary(op, a_h, b_h) for op in range(4)}
    want_mm = orc.tight_min_max(b_h)
    errors = []

    def worker(tid):
        try:
            for it in range(12):
                op = (tid + it) % 4
                r = a._bin(op, b)
                if not np.array_equal(bits(r.to_vec()), bits(want[op])):
                    errors.append((tid, it, "binary"))
                mn, mx = b.min_max()
                if (mn.bits, mx.bits) != (want_mm[0].bits, want_mm[1].bits):
                    errors.append((tid, it, "min_max"))
                if not (r == a._bin(op, b)) or r.convert(CellType.Float64).cmp(r) != 0:
                    errors.append((tid, it, "cmp"))
                m = MaskedCellBuffer.from_buffer_with_nodata(a, NoData.new(CellType.UInt16, a_h[tid]))
                if m.counts() != orc.mask_counts(a_h != a_h[tid]):
                    errors.append((tid, it, "counts"))
        except Exception as e:  # noqa: BLE001
            errors.append((tid, repr(e)))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(4)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert not errors, errors[:5]


def test_lazy_random_trees_match_eager(orc):
    """Pending chains through the operators — the precompiled fused shapes (ec.lazy()) and kernels specialised at run
    time (ec.lazy(jit=True)) — must be bit-identical to eager, op-by-op evaluation: random expression trees over up to
    6 inputs of mixed types, constants, shared sub-expressions, operands of different lengths, and a chain deep enough
    to hit the pending-depth cap (evaluated in pieces instead of recursing)."""
    rng = np.random.default_rng(2024)
    n = 2 * 32768 + 37
    types = [CellType.UInt8, CellType.UInt16, CellType.Int16, CellType.Float32, CellType.Float64, CellType.Int64]
    host = [cells(ct, n + (5 if i == 2 else 0), 0x900 + i) for i, ct in enumerate(types)]
    dev = [CellBuffer.from_vec(h) for h in host]

    def build(depth, leaves):
        """random tree; returns a function evaluating it on a list of buffers"""
        if depth == 0 or rng.random() < 0.25:
            k = int(rng.integers(0, leaves))
            return lambda b: b[k]
        op = int(rng.integers(0, 4))
        if rng.random() < 0.3:
            c = float(rng.choice([0.5, 2.0, -3.0, 1e-4, 0.0, 7.5]))
            sub = build(depth - 1, leaves)
            return lambda b: sub(b)._bin(op, c) if isinstance(sub(b), CellBuffer) else sub(b)
        lt, rt = build(depth - 1, leaves), build(depth - 1, leaves)
        return lambda b: lt(b)._bin(op, rt(b))

    for trial in range(40):
        leaves = int(rng.integers(1, 7))
        f = build(int(rng.integers(2, 6)), leaves)
        e = f(dev)
        for mode in (dict(), dict(jit=True)):
            with ec.lazy(**mode):
                lz = f(dev)
                assert lz == e, (trial, mode)          # device-side bitwise comparison forces the evaluation
    # shared sub-expression: evaluated once, used by two parents and by the user
    with ec.lazy():
        num = dev[1] - dev[2]
        a = (num * 2.0 + dev[0]) / (num - 1.0)
        b = num / 3.0
        ea = ((dev[1] - dev[2]) * 2.0 + dev[0]) / ((dev[1] - dev[2]) - 1.0)
    assert a == ea and b == (dev[1] - dev[2]) / 3.0 and num == dev[1] - dev[2]
    # a chain of 600 pending ops: deeper than the cap on pending depth, so it is evaluated in pieces as it is built
    for mode in (dict(), dict(jit=True)):
        with ec.lazy(**mode):
            x = dev[3]
            for i in range(300):
                x = x * 1.0001 + dev[4]
        y = dev[3]
        for i in range(300):
            y = y * 1.0001 + dev[4]
        assert x == y
    # the interpreted expression VM of ABI 1 is gone
    assert ec.lib().ec_set_lazy(2) == ec._lib.EC_INVALID_ARG


def test_lazy_jit_specialised_kernels_match_eager(orc):
    """ec.lazy(jit=True): a pending chain that no precompiled shape covers becomes ONE kernel specialised at run time
    (NVRTC, ec_jit.cu) — bit-identical to eager evaluation and to the oracle, cached by shape with scalars as
    parameters, ragged sizes, every cell type as an operand, shared sub-expressions, fallback past the size limits."""
    L = ec.lib()
    n = 3 * 32768 + 4099
    types = list(CellType)
    host = [cells(ct, n + (7 if i == 3 else 0), 0xA00 + i) for i, ct in enumerate(types)]
    dev = [CellBuffer.from_vec(h) for h in host]
    u8, u16, i16 = dev[int(CellType.UInt8)], dev[int(CellType.UInt16)], dev[int(CellType.Int16)]
    h8, h16, hi16 = host[int(CellType.UInt8)], host[int(CellType.UInt16)], host[int(CellType.Int16)]
    w = orc.tight_binary
    s = lambda op, a, c: orc.tight_scalar(op, a, orc.value(orc.Float64, c))

    def evi(nir, red, blue, g=2.5, c1=6.0, c2=7.5, l=1.0):
        return ((nir - red) * g) / (((nir + red * c1) - blue * c2) + l)
    eager = evi(u16, i16, u8)
    with ec.lazy(jit=True):
        k0, j0 = L.ec_kernel_launches(), L.ec_jit_cached_kernels()
        got = evi(u16, i16, u8).to_vec()
        if L.ec_last_kernel() != b"expression_jit(lazy)":
            pytest.skip("libnvrtc not loadable on this box: chains fall back to op-by-op evaluation")
        assert L.ec_kernel_launches() == k0 + 1 and L.ec_jit_cached_kernels() == j0 + 1
        # other scalars, same shape: no new kernel
        again = evi(u16, i16, u8, 2.4, 5.5, 7.0, 0.5).to_vec()
        assert L.ec_jit_cached_kernels() == j0 + 1 and L.ec_kernel_launches() == k0 + 2
    assert np.array_equal(bits(got), bits(eager.to_vec()))
    want = w(orc.DIV, s(orc.MUL, w(orc.SUB, h16, hi16), 2.5), s(orc.ADD, w(orc.SUB, w(orc.ADD, h16, s(orc.MUL, hi16, 6.0)), s(orc.MUL, h8, 7.5)), 1.0))
    assert np.array_equal(bits(got), bits(want))
    assert np.array_equal(bits(again), bits(evi(u16, i16, u8, 2.4, 5.5, 7.0, 0.5).to_vec()))
    # every cell type as an operand of a 3-op chain, all four ops, NaN/inf/-0.0/MIN/MAX specials included (cells())
    for ct in types:
        a, b = dev[int(ct)], dev[(int(ct) + 3) % 10]
        ha, hb = host[int(ct)], host[(int(ct) + 3) % 10]
        for op in range(4):
            with ec.lazy(jit=True):
                r = (a._bin(op, b) - a) / b
                assert r.len() == n
                got = r.to_vec()
            want = w(orc.DIV, w(orc.SUB, w(op, ha, hb), ha), hb)
            assert np.array_equal(bits(got), bits(want)), (ct, op)
    # tiny and ragged sizes go through the scalar tail
    for m in (1, 3, 5, 1023, 1025, 4097):
        with ec.lazy(jit=True):
            got = ((u16.view(0, m) + i16.view(0, m)) * 0.5 - u8.view(0, m)).to_vec()
        assert np.array_equal(bits(got), bits(w(orc.SUB, s(orc.MUL, w(orc.ADD, h16[:m], hi16[:m]), 0.5), h8[:m])))
    # a shared sub-expression is evaluated once and enters the kernel as an operand
    with ec.lazy(jit=True):
        num = u16 - i16
        a = (num * 2.0 + u8) / (num - 1.0)
        b = num / 3.0
    assert a == ((u16 - i16) * 2.0 + u8) / ((u16 - i16) - 1.0) and b == (u16 - i16) / 3.0 and num == u16 - i16
    # more than 8 scalars / 48 ops: falls back to op-by-op evaluation, same bits
    with ec.lazy(jit=True):
        x = dev[int(CellType.Float32)]
        for i in range(60):
            x = x * 1.0001 + dev[int(CellType.Float64)]
    y = dev[int(CellType.Float32)]
    for i in range(60):
        y = y * 1.0001 + dev[int(CellType.Float64)]
    assert x == y


def test_lazy_jit_cubins_persist_on_disk(tmp_path):
    """a second process finds the built kernel in $EC_JIT_CACHE and runs it without an NVRTC build"""
    import os
    import subprocess
    import sys
    script = (
        "import sys; sys.path.insert(0, %r)\n"
        "import numpy as np, erased_cells_b200 as ec\n"
        "from erased_cells_b200 import CellBuffer\n"
        "a = CellBuffer.from_vec(np.arange(10000, dtype=np.uint16)); b = CellBuffer.from_vec(np.arange(10000, dtype=np.int32) - 77)\n"
        "want = ((a - b) * 0.25 + a) / (b + 3.0)\n"
        "with ec.lazy(jit=True):\n"
        "    got = ((a - b) * 0.25 + a) / (b + 3.0)\n"
        "    got.device_ptr(); kernel = ec.lib().ec_last_kernel().decode()\n"
        "    same = got == want\n"
        "print('RESULT', same, kernel, ec.lib().ec_jit_builds())\n"
    ) % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, EC_JIT_CACHE=str(tmp_path / "jit"))
    outs = [subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, timeout=300, env=env) for _ in range(2)]
    lines = [[l for l in o.stdout.splitlines() if l.startswith("RESULT")] for o in outs]
    assert all(lines), outs[0].stderr[-2000:] + outs[1].stderr[-2000:]
    first, second = lines[0][0].split(), lines[1][0].split()
    if first[2] != "expression_jit(lazy)":
        pytest.skip("libnvrtc not loadable on this box")
    assert first[1] == "True" and first[3] == "1"
    assert second[1:] == ["True", "expression_jit(lazy)", "0"] and len(list((tmp_path / "jit").glob("*.cubin"))) == 1


def test_views_share_the_allocation(orc):
    """ec_buf_view: a strip of a resident raster as a buffer of its own — no copy, refcounted, copy on write."""
    n = 4 * 32768
    h = synth.host(CellType.Int16, n, 0x51E)
    whole = CellBuffer.from_vec(h)
    strips = [whole.view(g * 32768, 32768) for g in range(4)]
    for g, s in enumerate(strips):
        assert s.len() == 32768 and s.device_ptr() == whole.device_ptr() + g * 32768 * 2
        assert np.array_equal(s.to_vec(), h[g * 32768:(g + 1) * 32768])
        mn, mx = s.min_max()
        omn, omx = orc.tight_min_max(h[g * 32768:(g + 1) * 32768])
        assert (mn.bits, mx.bits) == (omn.bits, omx.bits)
    tail = whole.view(3 * 32768 + 128, 32768 - 128)          # any 32-byte aligned start
    assert np.array_equal(tail.to_vec(), h[3 * 32768 + 128:])
    del whole                                                 # the strips keep the allocation alive
    assert np.array_equal((strips[1] - strips[2]).to_vec(), orc.tight_binary(orc.SUB, h[32768:65536], h[65536:98304]))
    strips[0].put(5, CellValue(CellType.Int16, 1234))         # mutation through a view copies first
    assert strips[0].get(5).value() == 1234 and np.array_equal(strips[1].to_vec(), h[32768:65536])
    with pytest.raises(IndexError):
        strips[0].view(32768, 1)
    with pytest.raises(ec.EcError):
        strips[0].view(3, 10)


def test_masked_row_strips_of_a_resident_raster(orc):
    """MaskedCellBuffer.view = ec_buf_view + ec_mask_slice: strips of a resident masked raster behave like masked
    buffers built from the same cells (mask bits, counts, masked min_max, ops), including a ragged last strip whose
    length is not a multiple of 32 (the slice keeps the bits past its length zero)."""
    n = 3 * 4096 + 1000 + 13
    h = synth.host(CellType.Int16, n, 0x51F, kind=synth.INT_RANGE, lo=-50, hi=50)
    whole = MaskedCellBuffer.from_vec_with_nodata(h, NoData.new(CellType.Int16, 7))
    hm = h != 7
    bounds = [(0, 4096), (4096, 4096), (8192, 4096), (12288, n - 12288), (12288 + 128, 77), (4096, 0)]
    for off, ln in bounds:
        s = whole.view(off, ln)
        assert s.len() == ln and s.mask().len() == ln
        assert np.array_equal(s.mask().to_vec(), hm[off:off + ln])
        assert s.counts() == orc.mask_counts(hm[off:off + ln])
        if ln:
            mn, mx = s.min_max()
            omn, omx = orc.tight_min_max(h[off:off + ln], hm[off:off + ln])
            assert (mn.bits, mx.bits) == (omn.bits, omx.bits)
            assert (~s.mask()).counts() == orc.mask_counts(~hm[off:off + ln])   # tail bits of the slice are zero
    a, b = whole.view(0, 4096), whole.view(4096, 4096)
    d = a - b
    assert np.array_equal(bits(d.buffer().to_vec()), bits(orc.tight_binary(orc.SUB, h[:4096], h[4096:8192])))
    assert np.array_equal(d.mask().to_vec(), hm[:4096] & hm[4096:8192])
    assert whole.mask().slice(0, n) == whole.mask()
    with pytest.raises(IndexError):
        whole.mask().slice(128, n)
    with pytest.raises(ec.EcError):
        whole.mask().slice(32, 64)


def test_allocation_failure_and_trim():
    """An impossible allocation is an error code (MemoryError here), not a crash; the block cache can be handed back."""
    L = ec.lib()
    with pytest.raises(MemoryError):
        CellBuffer.with_defaults(1 << 42, CellType.Float64)      # 32 TiB
    a = CellBuffer.with_defaults(1 << 24, CellType.UInt8)          # the library is still usable afterwards
    assert a.min_max()[1].value() == 0
    del a
    assert L.ec_cached_bytes() >= (1 << 24)
    ec._lib.check(L.ec_trim())
    assert L.ec_cached_bytes() == 0
    assert CellBuffer.fill(100, CellValue(CellType.Int32, 7)).get(99).value() == 7


def test_integer_operand_division_guard_free_sequence(orc):
    """Division of integer cells runs div.rn.f64's Newton sequence without its range guards (div_int_operands,
    ec_common.cuh): every quotient, x/0 = +-inf and 0/0 = the x86 default NaN must still match the reference's
    `(l as f64) / (r as f64)` (src/value.rs:207) bit for bit — plain, fused with a scalar op, and as the quotient
    of the normalized difference (whose operands are sums / differences of cells, up to 2^65 for 64-bit types)."""
    n = (1 << 21) + 333
    rng = np.random.default_rng(0xD1F)
    for lct, rct in ((CellType.UInt8, CellType.UInt16), (CellType.Int16, CellType.Int16), (CellType.Int32, CellType.UInt8),
                     (CellType.UInt32, CellType.Int32), (CellType.Int64, CellType.UInt64), (CellType.UInt64, CellType.Int8),
                     (CellType.Int8, CellType.Int64), (CellType.UInt16, CellType.UInt32)):
        l, r = synth.host(lct, n, 0x501 + int(lct)).copy(), synth.host(rct, n, 0x601 + int(rct)).copy()
        # plenty of zeros, ones, extremes and small magnitudes on both sides
        for a in (l, r):
            info = np.iinfo(a.dtype)
            idx = rng.integers(0, n, n // 4)
            a[idx] = rng.integers(max(info.min, -3), min(info.max, 3) + 1, idx.size).astype(a.dtype)
            a[rng.integers(0, n, 64)] = info.max
            a[rng.integers(0, n, 64)] = info.min
        dl, dr = CellBuffer.from_vec(l), CellBuffer.from_vec(r)
        want = orc.tight_binary(orc.DIV, l, r)
        assert np.array_equal(bits((dl / dr).to_vec()), bits(want)), (lct, rct)
        half = orc.value(orc.Float64, 0.5)
        assert np.array_equal(bits(dl.binary_scalar(ec.DIV, dr, ec.MUL, 0.5).to_vec()), bits(orc.tight_scalar(orc.MUL, want, half))), (lct, rct)
        nd = orc.tight_binary(orc.DIV, orc.tight_binary(orc.SUB, l, r), orc.tight_binary(orc.ADD, l, r))
        assert np.array_equal(bits(dl.normalized_difference(dr).to_vec()), bits(nd)), (lct, rct)
        with ec.lazy():
            lz = (dl - dr) / (dl + dr)
            assert np.array_equal(bits(lz.to_vec()), bits(nd)), (lct, rct)


def test_integer_cells_divided_by_scalars_of_every_magnitude(orc):
    """`buf / scalar` on integer cells with ordinary, extreme, zero and non-finite divisors must give the reference's
    `(cell as f64) / (s as f64)` (src/buffer.rs:346-352, src/value.rs:207) bit for bit, also fused with a second scalar op.
    (A per-thread precomputed reciprocal was tried for this op and measured no faster — 101.9 vs 102.0 us on 8192^2 u16,
    the kernel is HBM-bound already — so the scalar path keeps div.rn.f64.)"""
    n = (1 << 20) + 77
    scalars = [10000.0, 3.0, -7.0, 1e-4, 0.1, -0.3, 2.0 ** 900, 2.0 ** -900, 2.0 ** 901, 2.0 ** -901, 1e300, -1e-300, 5e-324,
               0.0, -0.0, np.inf, -np.inf, np.nan, 1.0, -1.0, 65535.0, 1.7976931348623157e308]
    for ct in (CellType.UInt8, CellType.Int8, CellType.UInt16, CellType.Int16, CellType.UInt32, CellType.Int32, CellType.UInt64, CellType.Int64):
        h = cells(ct, n, 0x700 + int(ct))
        d = CellBuffer.from_vec(h)
        for s in scalars:
            want = orc.tight_scalar(orc.DIV, h, orc.value(orc.Float64, s))
            assert np.array_equal(bits((d / s).to_vec()), bits(want)), (ct, s)
        for s1, op2, s2 in ((10000.0, orc.ADD, 273.15), (-3.0, orc.MUL, 1e-3), (0.0, orc.SUB, 1.0), (2.0 ** 950, orc.DIV, 3.0)):
            want = orc.tight_scalar(op2, orc.tight_scalar(orc.DIV, h, orc.value(orc.Float64, s1)), orc.value(orc.Float64, s2))
            with ec.lazy():
                got = (d / s1)._bin(op2, s2)
                assert np.array_equal(bits(got.to_vec()), bits(want)), (ct, s1, op2, s2)
    # integer scalars too (`s as f64`), and a float buffer keeps the generic path
    h = cells(CellType.Int32, n, 0x7F0)
    d = CellBuffer.from_vec(h)
    for sv in (CellValue(CellType.UInt8, 7), CellValue(CellType.Int64, -(2 ** 62)), CellValue(CellType.UInt64, 2 ** 64 - 1)):
        want = orc.tight_scalar(orc.DIV, h, orc.value(int(sv.cell_type()), sv.value()))
        assert np.array_equal(bits((d / sv).to_vec()), bits(want)), sv
    hf = cells(CellType.Float32, n, 0x7F1)
    assert np.array_equal(bits((CellBuffer.from_vec(hf) / 10000.0).to_vec()), bits(orc.tight_scalar(orc.DIV, hf, orc.value(orc.Float64, 10000.0))))


def _special_cells(ct):
    dt = CellType(ct).dtype
    if dt == np.float32:
        raw = np.array([0x00000000, 0x80000000, 0x7F800000, 0xFF800000, 0x7FC00000, 0xFFC00000, 0x7F800001, 0xFFBFFFFF, 0x3F800000, 0xBF800000,
                        0x00000001, 0x80000001, 0x7F7FFFFF, 0x3FC00000, 0x80800000, 0x4B800000], np.uint32)
        return raw.view(np.float32)
    if dt == np.float64:
        raw = np.array([0, 1 << 63, 0x7FF << 52, 0xFFF << 52, 0x7FF8 << 48, 0xFFF8 << 48, (0x7FF << 52) | 1, (0xFFF << 52) | 0x7FFFFFFFFFFFF, 0x3FF << 52,
                        0xBFF << 52, 1, (1 << 63) | 1, 0x7FEFFFFFFFFFFFFF, 0x3FF8 << 48, 0x0010000000000000, 0x4330000000000000], np.uint64)
        return raw.view(np.float64)
    info = np.iinfo(dt)
    return np.array([0, 1, info.max, info.min, 7, info.max // 3] + ([-1, -7] if info.min < 0 else [2, 255]), dt)


def test_special_value_cross_product_all_ops_and_fused(orc):
    """Every pairing of zeros, infinities, quiet / signalling NaNs of both signs, subnormals and extremes, as a vectorised
    body and as a ragged tail: the four ops, the fused `(l op r) op s` and the normalized difference must match the
    reference's f64 arithmetic on its platform (src/value.rs:207; NaN results by the x86 rule) bit for bit. Covers the
    division variants by operand origin (integer / integer, f32 or integer on both sides, any f64)."""
    pairs = [(CellType.Float32, CellType.Float32), (CellType.Float32, CellType.Int16), (CellType.UInt8, CellType.Float32),
             (CellType.Float32, CellType.UInt64), (CellType.Int64, CellType.Float32), (CellType.Float32, CellType.Float64),
             (CellType.Float64, CellType.Float32), (CellType.Float64, CellType.Float64), (CellType.Int32, CellType.Int8),
             (CellType.Float64, CellType.UInt16)]
    for lct, rct in pairs:
        sl, sr = _special_cells(lct), _special_cells(rct)
        l0, r0 = np.repeat(sl, len(sr)), np.tile(sr, len(sl))
        reps = (3 * 16384 + 77) // len(l0) + 1
        l, r = np.tile(l0, reps)[: 3 * 16384 + 77].copy(), np.tile(r0, reps)[: 3 * 16384 + 77].copy()
        dl, dr = CellBuffer.from_vec(l), CellBuffer.from_vec(r)
        for op in range(4):
            want = orc.tight_binary(op, l, r)
            assert np.array_equal(bits(dl._bin(op, dr).to_vec()), bits(want)), (lct, rct, op)
            for op2, s in ((orc.MUL, 0.5), (orc.ADD, np.inf), (orc.DIV, 0.0)):
                got = dl.binary_scalar(op, dr, op2, s).to_vec()
                assert np.array_equal(bits(got), bits(orc.tight_scalar(op2, want, orc.value(orc.Float64, s)))), (lct, rct, op, op2, s)
        nd = orc.tight_binary(orc.DIV, orc.tight_binary(orc.SUB, l, r), orc.tight_binary(orc.ADD, l, r))
        assert np.array_equal(bits(dl.normalized_difference(dr).to_vec()), bits(nd)), (lct, rct)


def test_handles_cross_streams_and_threads(orc):
    """Handles are Send + Sync for real: a buffer produced on one thread's stream is read, combined and dropped on other
    threads that each enqueue on a stream of their own (ec_set_stream). The allocator orders a foreign stream against the
    block's home stream on first touch and does not recycle the block before every such stream has passed it — so a
    heavy producer followed at once by consumers elsewhere, and blocks freed on the "wrong" thread and reused right away,
    must still give the oracle's bits."""
    import threading

    import torch
    L = ec.lib()
    n = 6 * 1024 * 1024 + 17
    a_h, b_h = cells(CellType.UInt16, n, 0xD01), cells(CellType.Int16, n, 0xD02)
    a, b = CellBuffer.from_vec(a_h), CellBuffer.from_vec(b_h)
    want_sum = orc.tight_binary(orc.ADD, a_h, b_h)
    want = orc.tight_scalar(orc.MUL, want_sum, orc.value(orc.Float64, 0.5))
    errors = []

    def consumer(k, produced, out):
        try:
            s = torch.cuda.Stream()
            ec._lib.check(L.ec_set_stream(C.c_void_p(s.cuda_stream)))
            for rep in range(6):
                r = produced[rep] * 0.5                      # reads a block whose producer ran on the main thread's stream
                got = r.to_vec()
                if not np.array_equal(bits(got), bits(want)):
                    errors.append((k, rep, "wrong bits"))
                produced[rep] = None if k == 0 else produced[rep]   # thread 0 drops the handles: freed on a foreign thread
                del r
            ec._lib.check(L.ec_set_stream(None))
        except Exception as e:  # noqa: BLE001
            errors.append((k, repr(e)))

    for round_ in range(3):
        produced = [a + b for _ in range(6)]                 # queued back to back on the main stream, not waited for
        ts = [threading.Thread(target=consumer, args=(k, list(produced) if k else produced, None)) for k in range(3)]
        [t.start() for t in ts]
        churn = [(a - b) for _ in range(6)]                  # meanwhile the main stream recycles blocks of the same size
        [t.join() for t in ts]
        del churn, produced
        assert (a + b) == CellBuffer.from_vec(want_sum)
    assert not errors, errors[:5]


def test_views_alias_their_parent(orc):
    """ec_buf_view shares the allocation: a put through a view is seen by the parent and the other way round (views are
    aliases, pending lazy operands are snapshots); nothing about the cells is cached while a view lives."""
    h = synth.host(CellType.Float32, 4096, 0xD10, kind=synth.REAL_RANGE, lo=-5.0, hi=5.0)
    p = CellBuffer.from_vec(h)
    mn0, mx0 = p.min_max()
    v = p.view(1024, 512)
    v.put(3, np.float32(99.0))
    assert p.get(1027).value() == np.float32(99.0)
    assert p.min_max()[1].value() == np.float32(99.0)        # not the value remembered before the view existed
    p.put(1030, np.float32(-77.0))
    assert v.get(6).value() == np.float32(-77.0) and v.min_max()[0].value() == np.float32(-77.0)
    with ec.lazy():
        pending = p * 2.0                                    # a snapshot: later puts must not change it
        p.put(0, np.float32(1234.0))
        first = pending.get(0).value()
    assert first == np.float64(h[0]) * 2.0
    del v
    assert p.min_max()[1].value() == np.float32(1234.0)
