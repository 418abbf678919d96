"""Cross-checks inside the oracle: the faithful (per-cell tagged) flavour, the tight typed loops and
an independent numpy statement must agree bit for bit; plus the vectors no reference test pins
(u64/i64 -> f64 rounding, NaN payload rule, total order, signed MIN negation). CPU only."""
import numpy as np
import pytest

CT = range(10)


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view({1: "u1", 2: "u2", 4: "u4", 8: "u8"}[a.dtype.itemsize])


def rand_cells(rng, ct, n, orc, full_bits=True):
    dt = orc.DTYPES[ct]
    raw = rng.integers(0, 256, size=n * dt.itemsize, dtype=np.uint8)
    a = raw.view(dt).copy()
    if dt.kind == "f" and n >= 8:  # sprinkle specials
        sp = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, -np.nan, 1.0, -1.0], dtype=dt)
        a[rng.integers(0, n, size=8)] = sp
    return a


@pytest.mark.parametrize("lct", CT)
def test_faithful_equals_tight_binary(orc, lct):
    rng = np.random.default_rng(100 + lct)
    for rct in CT:
        l, r = rand_cells(rng, lct, 257, orc), rand_cells(rng, rct, 257, orc)
        for op in orc.OPS:
            f, t = orc.binary(op, l, r), orc.tight_binary(op, l, r)
            assert f.dtype == np.float64
            assert np.array_equal(bits(f), bits(t)), (lct, rct, op)
            # independent numpy statement, compared where the result is not NaN
            with np.errstate(all="ignore"):
                a, b = l.astype(np.float64), r.astype(np.float64)
                e = [a + b, a - b, a * b, a / b][op]
            ok = ~np.isnan(e)
            assert np.array_equal(np.isnan(f), np.isnan(e))
            assert np.array_equal(bits(f)[ok], bits(e)[ok]), (lct, rct, op)


def test_scalar_equals_binary_with_broadcast(orc):
    rng = np.random.default_rng(7)
    for lct in CT:
        l = rand_cells(rng, lct, 100, orc)
        for rct in CT:
            s = rand_cells(rng, rct, 1, orc)
            sv = orc.value(rct, s[0])
            for op in orc.OPS:
                assert np.array_equal(bits(orc.scalar(op, l, sv)), bits(orc.binary(op, l, np.repeat(s, 100))))
                assert np.array_equal(bits(orc.scalar(op, l, sv)), bits(orc.tight_scalar(op, l, sv)))


def test_convert_faithful_equals_tight_and_numpy(orc):
    rng = np.random.default_rng(11)
    for s in CT:
        a = rand_cells(rng, s, 300, orc)
        for d in CT:
            if not orc.can_fit_into(s, d):
                with pytest.raises(orc.NarrowingError):
                    orc.tight_convert(a, d)
                continue
            f, t = orc.convert(a, d), orc.tight_convert(a, d)
            assert f.dtype == orc.DTYPES[d]
            assert np.array_equal(bits(f), bits(t)), (s, d)
            with np.errstate(all="ignore"):
                e = a.astype(orc.DTYPES[d])
            ok = ~np.isnan(e) if e.dtype.kind == "f" else np.ones(len(e), bool)
            assert np.array_equal(bits(f)[ok], bits(e)[ok]), (s, d)


def test_int64_to_f64_round_to_nearest_even(orc):
    # Rust `as f64` == RNE; the reference's tests never exercise these (parity unpinned), vectors
    # from SURVEY.md §8c checked against g++ and numpy.
    u = np.array([2**64 - 1, 2**53 + 1, 2**53 + 2, 2**53 + 3, 2**63, 2**63 + 1024, 2**63 + 1025, 0], dtype=np.uint64)
    r = orc.convert(u, orc.Float64)
    assert r[0].hex() == "0x1.0000000000000p+64"
    assert r[1] == 2.0**53 and r[2] == 2.0**53 + 2 and r[3] == 2.0**53 + 4
    assert r[5] == 2.0**63 and r[6] == 2.0**63 + 2048
    i = np.array([0x7FFFFFFFFFFFFBFF, 0x7FFFFFFFFFFFFE00, -(2**63), -(2**53) - 1, 2**53 + 1], dtype=np.int64)
    r = orc.convert(i, orc.Float64)
    assert bits(r)[0] == 0x43DFFFFFFFFFFFFF
    assert r[1] == 2.0**63 and r[2] == -(2.0**63) and r[3] == -(2.0**53) and r[4] == 2.0**53
    # and through arithmetic: u64 + u8 goes through Float64
    assert orc.binary(orc.ADD, u[:1], np.zeros(1, np.uint8))[0] == 2.0**64


def test_nan_rule_x86(orc):
    """invalid ops produce the negative default NaN; a NaN lhs wins, else a NaN rhs (quieted)."""
    z = np.zeros(1)
    inf = np.array([np.inf])
    dflt = 0xFFF8000000000000
    assert bits(orc.binary(orc.DIV, z, z))[0] == dflt
    assert bits(orc.binary(orc.SUB, inf, inf))[0] == dflt
    assert bits(orc.binary(orc.MUL, inf, z))[0] == dflt
    assert bits(orc.binary(orc.DIV, np.zeros(1, np.uint8), np.zeros(1, np.uint16)))[0] == dflt
    a = np.array([0x7FF0000000000123], dtype=np.uint64).view(np.float64)  # signalling, payload 0x123
    b = np.array([0xFFF8000000000456], dtype=np.uint64).view(np.float64)
    one = np.ones(1)
    for op in orc.OPS:
        assert bits(orc.binary(op, a, b))[0] == 0x7FF8000000000123
        assert bits(orc.binary(op, b, a))[0] == 0xFFF8000000000456
        assert bits(orc.binary(op, one, a))[0] == 0x7FF8000000000123
        assert bits(orc.binary(op, a, one))[0] == 0x7FF8000000000123
    # f32 NaN operands are widened first (payload kept, quieted)
    f = np.array([0xFF800001], dtype=np.uint32).view(np.float32)
    assert bits(orc.binary(orc.ADD, f, one))[0] == 0xFFF8000020000000
    assert bits(orc.convert(f, orc.Float64))[0] == 0xFFF8000020000000


def test_total_order_min_max(orc):
    # seeds participate: min_max([+inf]) == (f32::MAX, +inf) — SURVEY.md fact 3
    mn, mx = orc.min_max(np.array([np.inf], dtype=np.float32))
    assert mn.numpy() == np.finfo(np.float32).max and mx.numpy() == np.inf
    mn, mx = orc.min_max(np.zeros(0, dtype=np.int16))
    assert (mn.numpy(), mx.numpy()) == (32767, -32768)
    a = np.array([1.0, -0.0, 0.0, np.nan, -np.nan, -np.inf, np.inf])
    mn, mx = orc.min_max(a)
    assert mn.bits == bits(np.array([-np.nan]))[0] and mx.bits == bits(np.array([np.nan]))[0]
    mn, mx = orc.min_max(np.array([0.0, -0.0]))
    assert mn.bits == 0x8000000000000000 and mx.bits == 0
    rng = np.random.default_rng(5)
    for ct in CT:
        a = rand_cells(rng, ct, 1000, orc)
        m = rng.random(1000) < 0.5
        for mask in (None, m):
            f, t = orc.min_max(a, mask), orc.tight_min_max(a, mask)
            assert (f[0].key(), f[1].key()) == (t[0].key(), t[1].key()), ct


def test_neg_all_types(orc):
    rng = np.random.default_rng(9)
    out_ct = [orc.Int16, orc.Int32, orc.Float64, orc.Float64, orc.Int8, orc.Int16, orc.Int32, orc.Int64, orc.Float32, orc.Float64]
    for ct in CT:
        a = rand_cells(rng, ct, 200, orc)
        r = orc.neg(a)
        assert r.dtype == orc.DTYPES[out_ct[ct]]
        if ct in (orc.Float32, orc.Float64):
            assert np.array_equal(bits(r), bits(a) ^ (1 << (a.dtype.itemsize * 8 - 1)))
        elif ct in (orc.UInt32, orc.UInt64):
            assert np.array_equal(bits(r), bits(-(a.astype(np.float64))))
        else:
            with np.errstate(over="ignore"):
                assert np.array_equal(r, (-(a.astype(r.dtype))).astype(r.dtype))
    # signed MIN wraps (release-mode behaviour of the reference)
    assert orc.neg(np.array([-128], np.int8))[0] == -128
    assert orc.neg(np.array([-(2**63)], np.int64))[0] == -(2**63)
    # -0u32 is -0.0
    assert bits(orc.neg(np.zeros(1, np.uint32)))[0] == 0x8000000000000000


def test_nodata_bitwise_semantics(orc):
    # Default matches only the canonical positive quiet NaN; Value(0.0) does not match -0.0
    a = np.array([0x7FF8000000000000, 0xFFF8000000000000, 0x7FF8000000000001], dtype=np.uint64).view(np.float64)
    assert list(orc.mask_from_nodata(a, orc.ND_DEFAULT)) == [False, True, True]
    z = np.array([0.0, -0.0])
    assert list(orc.mask_from_nodata(z, orc.ND_VALUE, orc.value(orc.Float64, 0.0))) == [False, True]
    rng = np.random.default_rng(3)
    for ct in CT:
        a = rand_cells(rng, ct, 500, orc)
        nd = orc.value(ct, a[17])
        m = orc.mask_from_nodata(a, orc.ND_VALUE, nd)
        assert np.array_equal(m, bits(a) != bits(a[17:18])[0])


def test_fill_nodata_and_truncation(orc):
    a = np.arange(10, dtype=np.uint8)
    m = np.arange(10) % 3 != 0
    out = orc.fill_nodata(a, m, orc.Float32, orc.ND_VALUE, orc.value(orc.Float32, -9999.0))
    assert out.dtype == np.float32 and list(out[:4]) == [-9999.0, 1.0, 2.0, -9999.0]
    assert np.array_equal(orc.fill_nodata(a, m, orc.UInt16, orc.ND_NONE), a.astype(np.uint16))
    with pytest.raises(orc.NarrowingError):
        orc.fill_nodata(a.astype(np.float64), m, orc.Int32, orc.ND_DEFAULT)
    # length mismatch truncates to the shorter operand (zip) — src/buffer.rs:327
    assert len(orc.binary(orc.ADD, np.ones(5, np.uint8), np.ones(3, np.float32))) == 3


def test_special_value_cross_product_against_an_independent_statement(orc):
    """Every pairing of zeros, infinities, NaNs of both signs and kinds, subnormals and extremes through the four ops:
    the oracle's non-NaN results equal numpy's IEEE arithmetic bit for bit, and its NaN results follow the x86 rule
    stated here independently (first NaN operand, quieted; else the negative default NaN). This is the table the GPU
    test of the same name checks the kernels against."""
    raw64 = np.array([0, 1 << 63, 0x7FF << 52, 0xFFF << 52, 0x7FF8 << 48, 0xFFF8 << 48, (0x7FF << 52) | 1, (0xFFF << 52) | 0x7FFFFFFFFFFFF,
                      0x3FF << 52, 0xBFF << 52, 1, (1 << 63) | 1, 0x7FEFFFFFFFFFFFFF, 0x3FF8 << 48, 0x0010000000000000, 0x4330000000000000], np.uint64)
    raw32 = np.array([0x00000000, 0x80000000, 0x7F800000, 0xFF800000, 0x7FC00000, 0xFFC00000, 0x7F800001, 0xFFBFFFFF, 0x3F800000, 0xBF800000,
                      0x00000001, 0x80000001, 0x7F7FFFFF, 0x3FC00000, 0x80800000, 0x4B800000], np.uint32)

    def widen32(u):  # cvtss2sd on NaNs: sign and payload kept, quiet bit set
        f = u.view(np.float32)
        with np.errstate(all="ignore"):
            d = f.astype(np.float64).view(np.uint64).copy()
        nan = np.isnan(f)
        d[nan] = ((u[nan].astype(np.uint64) & 0x80000000) << 32) | 0x7FF8000000000000 | ((u[nan].astype(np.uint64) & 0x007FFFFF) << 29)
        return d

    for lraw, rraw in ((raw64, raw64), (raw32, raw32), (raw32, raw64), (raw64, raw32)):
        l0, r0 = np.repeat(lraw, len(rraw)), np.tile(rraw, len(lraw))
        lcells = l0.view(np.float32 if l0.dtype == np.uint32 else np.float64)
        rcells = r0.view(np.float32 if r0.dtype == np.uint32 else np.float64)
        lb = widen32(l0) if l0.dtype == np.uint32 else l0.copy()
        rb = widen32(r0) if r0.dtype == np.uint32 else r0.copy()
        ld, rd = lb.view(np.float64), rb.view(np.float64)
        for op, fn in ((orc.ADD, np.add), (orc.SUB, np.subtract), (orc.MUL, np.multiply), (orc.DIV, np.divide)):
            with np.errstate(all="ignore"):
                plain = fn(ld, rd)
            want = plain.view(np.uint64).copy()
            isn = np.isnan(plain)
            want[isn] = 0xFFF8000000000000
            rn, ln = np.isnan(rd), np.isnan(ld)
            want[rn] = rb[rn] | 0x0008000000000000
            want[ln] = lb[ln] | 0x0008000000000000
            got = bits(orc.tight_binary(op, lcells, rcells))
            assert np.array_equal(got, want), (lraw.dtype, rraw.dtype, op, np.flatnonzero(got != want)[:5])
            assert np.array_equal(bits(orc.binary(op, lcells, rcells)), want)
