"""Statistics extension on the device (ec_buf_statistics / ec_buf_moments through the C ABI) against the numpy
restatement of the same definition in oracle/oracle.py. The reference has no statistics: parity is
UNPINNED, what is pinned is that the device result equals the order-independent definition bit for bit — count, min,
max, mean, stddev and the raw 128-bit window sums — for every cell type, ragged sizes, masks and non-finite cells."""
import math
import struct

import numpy as np
import pytest

from erased_cells_b200 import CellBuffer, CellType, Mask, MaskedCellBuffer, sharding, synth

pytestmark = pytest.mark.gpu
CT = list(CellType)
SIZES = [1, 3, 31, 33, 1000, 4097, 16384 + 5, 5 * 32768 + 77]


def fbits(x: float) -> int:
    return struct.unpack("<Q", struct.pack("<d", x))[0]


def finite_cells(ct, n, seed, kind):
    dt = ct.dtype
    if kind == "range":
        return synth.host(ct, n, seed, kind=synth.REAL_RANGE, lo=-100.0 if ct.is_signed() else 5.0, hi=125.0)
    a = synth.host(ct, n, seed).copy()  # full bit range: every exponent for floats, MIN..MAX for integers
    if dt.kind == "f":
        a[~np.isfinite(a)] = dt.type(1.5)
    return a


def same(got, want):
    assert got.count == want["count"]
    assert (got.min.bits, got.max.bits) == (want["min"].bits, want["max"].bits)
    for g, w in ((got.mean, want["mean"]), (got.stddev, want["stddev"])):
        assert (math.isnan(g) and math.isnan(w)) or fbits(g) == fbits(w), (g, w)


def raw_words(raw):
    w = [raw[0]]
    for k in range(4):
        v = raw[1 + k] & ((1 << 128) - 1)
        w += [v & (2 ** 64 - 1), v >> 64]
    return w


@pytest.mark.parametrize("ct", CT)
@pytest.mark.parametrize("kind", ["range", "full"])
def test_statistics_parity(orc, ct, kind):
    for n in SIZES:
        a = finite_cells(ct, n, 0x5A00 + int(ct) + n, kind)
        valid = synth.host(CellType.UInt8, n, 0x5B00 + n) < 180
        buf = CellBuffer.from_vec(a)
        same(buf.statistics(), orc.statistics(a))
        same(MaskedCellBuffer(buf, Mask.new(valid)).statistics(), orc.statistics(a, valid))


@pytest.mark.parametrize("ct", CT)
def test_raw_window_sums(orc, ct):
    n = 3 * 32768 + 1234
    a = finite_cells(ct, n, 0x5C00 + int(ct), "full")
    valid = synth.host(CellType.UInt8, n, 0x5C7F) < 128
    buf, mask = CellBuffer.from_vec(a), Mask.new(valid)
    for m, hm in ((None, None), (mask, valid)):
        mn, mx = orc.min_max(a, hm)
        kind, p, e = orc.statistics_plan(mn, mx)
        assert kind == orc.ST_REGULAR
        assert [int(x) for x in sharding.moments(buf, m, p, e)] == raw_words(orc.moments_raw(a, hm, p, e))


def test_non_finite_and_empty(orc):
    cases = [
        (np.array([], dtype=np.float32), None),
        (np.array([], dtype=np.uint16), None),
        (np.array([1.0, np.inf, 3.0], dtype=np.float64), None),
        (np.array([1.0, -np.inf, 3.0], dtype=np.float32), None),
        (np.array([np.inf, -np.inf], dtype=np.float32), None),
        (np.array([1.0, np.nan], dtype=np.float32), None),
        (np.array([1.0, -np.nan], dtype=np.float64), None),
        (np.array([np.nan, 5.0, np.inf], dtype=np.float32), np.array([0, 1, 0], dtype=bool)),  # specials masked out
        (np.array([3, 4, 5], dtype=np.int16), np.array([0, 0, 0], dtype=bool)),
        (np.full(70000, 7, dtype=np.uint8), None),
        (np.array([np.finfo(np.float64).max, -np.finfo(np.float64).max, 0.0]), None),
        (np.array([5e-324, 0.0, 1e-323]), None),
        (np.array([2 ** 64 - 1, 0, 2 ** 63], dtype=np.uint64), None),
        (np.array([-2 ** 63, 2 ** 63 - 1], dtype=np.int64), None),
    ]
    for a, valid in cases:
        buf = CellBuffer.from_vec(a) if len(a) else CellBuffer.with_defaults(0, CellType.of(a))
        got = buf.statistics() if valid is None else MaskedCellBuffer(buf, Mask.new(valid)).statistics()
        same(got, orc.statistics(a, valid))


def test_strip_and_grid_independence_at_scale(orc):
    """2^26 + ragged cells made on the device; the whole-buffer result must equal the finish over three unequal strips'
    raw sums (different grids, different tails) and the oracle on the downloaded cells."""
    n = (1 << 26) + 12345
    for ct, lo, hi in ((CellType.UInt16, 5000, 40000), (CellType.Int32, -2.0e9, 2.0e9), (CellType.Float32, -1.0e4, 1.0e4)):
        buf = synth.device(ct, n, 0xEC77, kind=synth.REAL_RANGE, lo=lo, hi=hi)
        whole = buf.statistics()
        kind, p, e = sharding.statistics_plan(whole.min, whole.max)
        cuts = [0, 1 << 20, (1 << 25) + 4096, n]
        raws = [sharding.moments(buf.view(i, j - i), None, p, e) for i, j in zip(cuts[:-1], cuts[1:])]
        again = sharding.finish_statistics(raws, whole.min, whole.max)
        assert (again.count, fbits(again.mean), fbits(again.stddev)) == (whole.count, fbits(whole.mean), fbits(whole.stddev))
        host = buf.to_vec()
        same(whole, orc.statistics(host))
        # Float32 cells are summed on the grid 2^(E - 26): the mean is within half a grid step, in practice far closer
        slack = 2.0 ** (14 - 26) * 1e-3 if ct == CellType.Float32 else 0.0
        assert abs(whole.mean - host.astype(np.float64).mean()) <= 1e-9 * abs(whole.mean) + 1e-9 + slack
        assert abs(whole.stddev - host.astype(np.float64).std()) <= 1e-9 * whole.stddev + slack
