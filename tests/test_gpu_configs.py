"""BASELINE.json configs at their full sizes on the B200, checked through size-independent properties
(device-side comparisons, round trips, strip decompositions) plus sampled windows against the CPU
oracle — the oracle cannot chew 10^9 cells in seconds, the windows it can."""
import ctypes as C

import numpy as np
import pytest

import erased_cells_b200 as ec
from erased_cells_b200 import CellBuffer, CellType, CellValue, Mask, MaskedCellBuffer, NoData, sharding, synth

pytestmark = pytest.mark.gpu
T = CellType
WIN = 1 << 20


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view({1: "u1", 2: "u2", 4: "u4", 8: "u8"}[a.dtype.itemsize])


def window(buf: CellBuffer, off: int, n: int) -> np.ndarray:
    """D2H of cells [off, off+n) of a device buffer (off on a 128-cell boundary)."""
    sz = buf.cell_type().size_of()
    return CellBuffer.wrap_device(buf.cell_type(), buf.device_ptr() + off * sz, n).to_vec()


def offsets(n):
    return [0, ((n // 2) // 128) * 128, n - WIN]


def test_config1_readme_4096(orc):
    n = 4096 * 4096
    args_a = dict(kind=synth.INT_RANGE, lo=0, hi=255)
    args_b = dict(kind=synth.INT_RANGE, lo=0, hi=65535)  # includes 0: x/0 = inf, 0/0 = NaN
    a, b = synth.device(T.UInt8, n, 0xEC01, **args_a), synth.device(T.UInt16, n, 0xEC02, **args_b)
    ha, hb = synth.host(T.UInt8, n, 0xEC01, **args_a), synth.host(T.UInt16, n, 0xEC02, **args_b)
    r = a / b * 0.5
    want = orc.tight_scalar(orc.MUL, orc.tight_binary(orc.DIV, ha, hb), orc.value(orc.Float64, 0.5))
    assert r.cell_type() == T.Float64
    got = r.to_vec()
    assert np.array_equal(bits(got), bits(want))  # all 16.8 M cells, bitwise (NaN rule included)
    assert np.isnan(want).any() and np.isinf(want).any()
    assert a.binary_scalar(ec.DIV, b, ec.MUL, 0.5) == r  # fused chain == operator chain


def test_config2_convert_sweep_8192(orc):
    n = 8192 * 8192
    for s in T:
        src = synth.device(s, n, 0xEC10 + int(s))
        as_f64 = src.convert(T.Float64)
        smn, smx = src.min_max()
        for d in T:
            if not s.can_fit_into(d):
                with pytest.raises(ec.NarrowingError) as e:
                    src.convert(d)
                assert (e.value.src, e.value.dst) == (int(s), int(d))
                continue
            out = src.convert(d)
            assert out.cell_type() == d and out.len() == n
            # widening is transitive: S -> D -> f64 == S -> f64 (device-side compare, no D2H)
            assert out.convert(T.Float64) == as_f64, (s, d)
            # widening is monotone under total order, so min_max commutes with convert — except f32 -> f64 on
            # full-bit-range data, where quieting a signalling NaN can lift it above a quiet one
            if not (s == T.Float32 and d == T.Float64):
                dmn, dmx = out.min_max()
                assert (dmn.bits, dmx.bits) == (smn.convert(d).bits, smx.convert(d).bits), (s, d)
            for off in offsets(n)[:2]:
                hw = synth.host(s, WIN, 0xEC10 + int(s), index_offset=off)
                assert np.array_equal(bits(window(out, off, WIN)), bits(orc.tight_convert(hw, int(d)))), (s, d, off)
            del out


def test_config3_masked_i16_16384(orc):
    n = 16384 * 16384
    kw = dict(kind=synth.INT_RANGE, lo=-32768, hi=32767, period=50, sentinel=-32768)
    a, b = synth.device(T.Int16, n, 0xEC31, **kw), synth.device(T.Int16, n, 0xEC32, **kw)
    nd = NoData.default(T.Int16)
    ma, mb = MaskedCellBuffer.from_buffer_with_nodata(a, nd), MaskedCellBuffer.from_buffer_with_nodata(b, nd)
    r = (ma - mb) * 0.0001
    assert r.cell_type() == T.Float64 and r.len() == n
    # mask algebra at full size: |A & B| + |~A | ~B| == n ; counts add up; ~2% + 1/65536 of cells are sentinels
    da, na = ma.counts()
    assert da + na == n and 0.015 * n < na < 0.025 * n
    both = r.counts()
    either_invalid = (~ma.mask() | ~mb.mask()).counts()[0]
    assert both[0] + either_invalid == n and r.mask() == (ma.mask() & mb.mask())
    # fused (masked sub) == unfused (buffer sub, mask and)
    assert (ma - mb).buffer() == a - b
    # windows against the oracle: data, mask, and the masked min_max of the window
    s = orc.value(orc.Float64, 0.0001)
    for off in offsets(n):
        ha, hb = synth.host(T.Int16, WIN, 0xEC31, index_offset=off, **kw), synth.host(T.Int16, WIN, 0xEC32, index_offset=off, **kw)
        want = orc.tight_scalar(orc.MUL, orc.tight_binary(orc.SUB, ha, hb), s)
        wm = orc.mask_and(orc.mask_from_nodata(ha, orc.ND_DEFAULT), orc.mask_from_nodata(hb, orc.ND_DEFAULT))
        assert np.array_equal(bits(window(r.buffer(), off, WIN)), bits(want))
        words = CellBuffer.wrap_device(T.UInt32, ec.lib().ec_mask_device_words(r.mask()._h) + off // 8, WIN // 32).to_vec()
        assert np.array_equal(np.unpackbits(words.view(np.uint8), bitorder="little").astype(bool), wm)
    # masked min_max: equals the strip-wise combination (4 strips) and lies inside the value range
    mn, mx = r.min_max()
    keys = []
    for g in range(4):
        off, ln = sharding.row_strip(16384, 16384, 4, g)
        sb = CellBuffer.wrap_device(T.Float64, r.buffer().device_ptr() + off * 8, ln)
        sm = Mask.new(np.zeros(0, bool))  # placeholder, replaced below
        words_ptr = ec.lib().ec_mask_device_words(r.mask()._h) + off // 8
        # a strip of the mask is a word range of the same allocation: wrap it through a device->device copy
        mw = CellBuffer.wrap_device(T.UInt32, words_ptr, ln // 32).to_vec()
        sm = Mask.new(np.unpackbits(mw.view(np.uint8), bitorder="little").astype(bool))
        smn, smx = MaskedCellBuffer(sb, sm).min_max()
        keys.append(sharding.keys_of(smn, smx))
    cmn, cmx = sharding.values_of(T.Float64, np.min(np.stack(keys), axis=0))
    assert (mn.bits, mx.bits) == (cmn.bits, cmx.bits)
    assert -6.5536 <= float(mn.value()) < float(mx.value()) <= 6.5536


def test_config4_f32_32768_min_max_strips(orc):
    n = 32768 * 32768
    buf = synth.device(T.Float32, n, 0xEC40, kind=synth.REAL_RANGE, lo=-1e4, hi=1e4)
    mn, mx = buf.min_max()
    assert -1e4 <= float(mn.value()) < -9999.9 and 9999.9 < float(mx.value()) <= 1e4
    # identical for every strip count (1/2/4/8): combine per-strip keys with MIN, as the all-reduce does
    for g_count in (2, 4, 8):
        keys = []
        for g in range(g_count):
            off, ln = sharding.row_strip(32768, 32768, g_count, g)
            strip = CellBuffer.wrap_device(T.Float32, buf.device_ptr() + off * 4, ln)
            keys.append(sharding.keys_of(*strip.min_max()))
        cmn, cmx = sharding.values_of(T.Float32, np.min(np.stack(keys), axis=0))
        assert (cmn.bits, cmx.bits) == (mn.bits, mx.bits), g_count
    # the extremes are where the oracle says they are, for the windows holding them
    for off in offsets(n):
        hw = synth.host(T.Float32, WIN, 0xEC40, index_offset=off, kind=synth.REAL_RANGE, lo=-1e4, hi=1e4)
        w = CellBuffer.wrap_device(T.Float32, buf.device_ptr() + off * 4, WIN)
        omn, omx = orc.tight_min_max(hw)
        assert tuple(v.bits for v in w.min_max()) == (omn.bits, omx.bits)
    # adversarial variant: total order and the (MAX, MIN) seeds
    specials = {5: 0x7FC00001, 123456789: 0xFFC00000, 77: 0x7F800000, 999999999: 0xFF800000, 31: 0x80000000, 64: 0}
    for i, b in specials.items():
        buf.put(i, CellValue(T.Float32, np.array([b], np.uint32).view(np.float32)[0]))
    mn, mx = buf.min_max()
    assert (mn.bits, mx.bits) == (0xFFC00000, 0x7FC00001)  # -NaN is the minimum, +NaN(payload 1) the maximum
    f64mn, f64mx = buf.convert(T.Float64).min_max()
    assert (f64mn.bits, f64mx.bits) == (0xFFF8000000000000, 0x7FF8000020000000)


def test_config5_ndvi_u16_32768_tile(orc):
    n = 32768 * 32768
    kw = dict(kind=synth.INT_RANGE, lo=5000, hi=40000, period=1000, sentinel=0)
    nir, red = synth.device(T.UInt16, n, 0xEC50, **kw), synth.device(T.UInt16, n, 0xEC58, **kw)
    fused = nir.normalized_difference(red)
    unfused = (nir - red) / (nir + red)
    assert fused.cell_type() == T.Float64 and fused == unfused  # 8.6 GB compared on the device
    fmn, fmx = fused.min_max()
    assert (fmn.bits, fmx.bits) == tuple(v.bits for v in unfused.min_max())
    del unfused
    # ~0.1 % zeros per band: (0-r)/(0+r) = -1, (n-0)/(n+0) = 1, and where both are 0 the 0/0 is the x86 default
    # NaN, which is NEGATIVE and therefore the minimum under total order (SURVEY.md §7 risk 2)
    assert fmn.bits == 0xFFF8000000000000 and float(fmx.value()) == 1.0
    for off in offsets(n):
        hn, hr = synth.host(T.UInt16, WIN, 0xEC50, index_offset=off, **kw), synth.host(T.UInt16, WIN, 0xEC58, index_offset=off, **kw)
        want = orc.tight_binary(orc.DIV, orc.tight_binary(orc.SUB, hn, hr), orc.tight_binary(orc.ADD, hn, hr))
        assert np.array_equal(bits(window(fused, off, WIN)), bits(want))


def test_more_than_2_pow_32_cells(orc):
    """Maximum sizes: 2^32 + 4099 u8 cells (4.3 GB) — 64-bit cell indices everywhere. convert / binary are checked on
    windows that straddle the 2^32 boundary and end at the last (ragged) cell; min_max, counts and statistics over the
    whole buffer must equal what three unequal views (one across the boundary) combine to."""
    n = (1 << 32) + 4099
    a = synth.device(T.UInt8, n, 0xEC99, kind=synth.INT_RANGE, lo=1, hi=250, period=1 << 20, sentinel=0)
    assert a.len() == n
    wide = a.convert(T.UInt16)                       # 8.6 GB written
    assert wide.len() == n and wide.cell_type() == T.UInt16
    for off in (0, (1 << 32) - WIN // 2 - ((1 << 32) - WIN // 2) % 128, ((n - WIN) // 128) * 128):
        m = min(WIN, n - off)
        want = synth.host(T.UInt8, m, 0xEC99, index_offset=off, kind=synth.INT_RANGE, lo=1, hi=250, period=1 << 20, sentinel=0)
        assert np.array_equal(window(a, off, m), want), off
        assert np.array_equal(window(wide, off, m), want.astype(np.uint16)), off
    tail = n - ((n - WIN) // 128) * 128
    assert np.array_equal(wide.view(n - tail, tail).to_vec()[-5:], synth.host(T.UInt8, 5, 0xEC99, index_offset=n - 5, kind=synth.INT_RANGE, lo=1, hi=250,
                                                                           period=1 << 20, sentinel=0).astype(np.uint16))
    del wide
    mn, mx = a.min_max()
    assert (mn.value(), mx.value()) == (0, 250)
    masked = MaskedCellBuffer.from_buffer_with_nodata(a, NoData.new(T.UInt8, 0))
    data, nodata = masked.counts()
    assert data + nodata == n and 0 < nodata < n // (1 << 19)
    st, mst = a.statistics(), masked.statistics()
    assert st.count == n and mst.count == data and mst.min.value() == 1
    cuts = [0, (1 << 31) + 128 * 7, (1 << 32) + 128 * 3, n]  # the middle view starts below and ends above 2^32
    kind, p, e = sharding.statistics_plan(st.min, st.max)
    raws = [sharding.moments(a.view(i, j - i), None, p, e) for i, j in zip(cuts[:-1], cuts[1:])]
    again = sharding.finish_statistics(raws, st.min, st.max)
    assert (again.count, again.mean, again.stddev) == (st.count, st.mean, st.stddev)
    assert abs(st.mean - 125.5) < 0.01 and abs(mst.mean - 125.5) < 0.01  # uniform 1..250 (+ a few zeros)
    part_mm = [a.view(i, j - i).min_max() for i, j in zip(cuts[:-1], cuts[1:])]
    assert min(v[0].value() for v in part_mm) == 0 and max(v[1].value() for v in part_mm) == 250
