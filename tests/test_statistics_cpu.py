"""Statistics extension (count / min / max / mean / population stddev), CPU side.

The reference has no statistics beyond min_max and counts, so parity is UNPINNED: what is checked
here is (1) that the repo's own definition (oracle/oracle.py) is accurate — against exact rational
arithmetic, tolerance 2 ULP (of the larger of the mean and the largest cell magnitude), (2) that it is independent of cell order and of how the cells are split into strips, and
(3) that the library's host-side plan/finish (no GPU involved) agree bit for bit with the oracle's."""
import ctypes as C
import math
import struct
from fractions import Fraction

import numpy as np
import pytest

from erased_cells_b200 import _lib

ALL = ["u1", "u2", "u4", "u8", "i1", "i2", "i4", "i8", "f4", "f8"]


def bits(x: float) -> int:
    return struct.unpack("<Q", struct.pack("<d", x))[0]


def ulps(a: float, b: float) -> int:
    def k(x):
        u = bits(x)
        return (1 << 63) - (u & ~(1 << 63)) if u >> 63 else u + (1 << 63)
    return abs(k(a) - k(b))


def sample(rng, dt, n, kind="full"):
    dt = np.dtype(dt)
    if dt.kind == "f":
        if kind == "full":
            return (rng.standard_normal(n) * 10.0 ** rng.integers(-20, 20)).astype(dt)
        return (rng.standard_normal(n) + 1.0e6).astype(dt)  # large mean, small spread: cancellation-prone
    ii = np.iinfo(dt)
    if kind == "full":
        return rng.integers(ii.min, ii.max, n, dtype=dt, endpoint=True)
    lo = ii.max - 100
    return rng.integers(lo, ii.max, n, dtype=dt, endpoint=True)


def exact(a):
    xs = [Fraction(float(v)) for v in a.astype(np.float64)]
    mean = sum(xs) / len(xs)
    var = sum((v - mean) ** 2 for v in xs) / len(xs)
    # sqrt of an exact rational, correctly rounded enough for a 2-ULP check: scale to an integer square root
    sd = math.sqrt(var) if var < 2 ** 900 else math.sqrt(var / 4 ** 400) * 2.0 ** 400
    return float(mean), sd


@pytest.mark.parametrize("dt", ALL)
@pytest.mark.parametrize("kind", ["full", "narrow"])
def test_definition_is_accurate(orc, dt, kind):
    rng = np.random.default_rng(hash((dt, kind)) & 0xFFFF)
    a = sample(rng, dt, 3000, kind)
    s = orc.statistics(a)
    mean, sd = exact(a)
    assert s["count"] == len(a)
    # 2 ULP of the mean, or — when the mean is small against the data (cancellation in pivot + offset) — 2 ULP of
    # the largest cell magnitude
    top = float(np.abs(a.astype(np.float64)).max())
    if dt == "f4":
        # Float32 cells are summed on the grid 2^(E - 26), 2^E > the largest magnitude: every cell is off by at most half
        # a grid step (an eighth of the ulp of the largest cells), so mean and stddev are within one grid step
        grid = 2.0 ** (math.frexp(top)[1] - 26)
        assert abs(s["mean"] - mean) <= 0.5 * grid + 2 * 2.0 ** -52 * top, (s["mean"], mean)
        assert abs(s["stddev"] - sd) <= grid + 1e-13 * abs(sd), (s["stddev"], sd)
        # and exact whenever every cell sits on the grid (all magnitudes within 2^3 of the largest): integers here
        ints = rng.integers(-4000, 4000, 3000).astype(np.float32)
        si, (mi, sdi) = orc.statistics(ints), exact(ints)
        assert ulps(si["mean"], mi) <= 2 or abs(si["mean"] - mi) <= 2 * 2.0 ** -52 * 4000, (si["mean"], mi)
        assert abs(si["stddev"] - sdi) <= 4e-16 * sdi + 1e-13 * abs(si["stddev"])
        return
    assert ulps(s["mean"], mean) <= 2 or abs(s["mean"] - mean) <= 2 * 2.0 ** -52 * top, (s["mean"], mean)
    if sd == 0:
        assert s["stddev"] == 0
    else:
        # m2 - m1^2 loses what the pivot cannot remove: allow 2 ULP of stddev plus the cancellation of the definition
        assert abs(s["stddev"] - sd) <= 4e-16 * sd + 1e-13 * abs(s["stddev"]), (s["stddev"], sd)


@pytest.mark.parametrize("dt", ALL)
def test_order_and_strip_independence(orc, dt):
    rng = np.random.default_rng(7)
    a = sample(rng, dt, 5000)
    m = rng.random(5000) < 0.7
    whole = orc.statistics(a, m)
    perm = rng.permutation(5000)
    again = orc.statistics(a[perm], m[perm])
    assert (bits(whole["mean"]), bits(whole["stddev"])) == (bits(again["mean"]), bits(again["stddev"]))
    mn, mx = orc.min_max(a, m)
    kind, p, e = orc.statistics_plan(mn, mx)
    assert kind == orc.ST_REGULAR
    for cuts in ([0, 5000], [0, 1, 5000], [0, 1234, 1234, 4096, 5000]):
        raws = [orc.moments_raw(a[i:j], m[i:j], p, e) for i, j in zip(cuts[:-1], cuts[1:])]
        s = orc.statistics_finish(raws, mn, mx)
        assert s["count"] == int(m.sum())
        assert (bits(s["mean"]), bits(s["stddev"])) == (bits(whole["mean"]), bits(whole["stddev"]))


def test_degenerate(orc):
    s = orc.statistics(np.array([], dtype=np.int16))
    assert s["count"] == 0 and math.isnan(s["mean"]) and math.isnan(s["stddev"])
    s = orc.statistics(np.array([1, 2, 3], dtype=np.float32), np.array([0, 0, 0], dtype=bool))
    assert s["count"] == 0 and math.isnan(s["mean"])
    s = orc.statistics(np.array([1, np.inf], dtype=np.float64))
    assert s["mean"] == math.inf and math.isnan(s["stddev"]) and s["count"] == 2
    s = orc.statistics(np.array([-np.inf, np.inf], dtype=np.float32))
    assert math.isnan(s["mean"])
    s = orc.statistics(np.array([1, np.nan], dtype=np.float32))
    assert math.isnan(s["mean"]) and math.isnan(s["stddev"])
    s = orc.statistics(np.array([np.nan, 5.0], dtype=np.float32), np.array([0, 1], dtype=bool))  # NaN masked out
    assert s["mean"] == 5.0 and s["stddev"] == 0.0 and s["count"] == 1
    s = orc.statistics(np.full(100, 7, dtype=np.uint8))
    assert s["mean"] == 7.0 and s["stddev"] == 0.0
    big = np.array([np.finfo(np.float64).max, -np.finfo(np.float64).max, 0.0])
    s = orc.statistics(big)
    assert s["mean"] == 0.0 and math.isfinite(s["stddev"]) and s["stddev"] > 1e307
    tiny = np.array([5e-324, 0.0, 1e-323])
    s = orc.statistics(tiny)
    assert ulps(s["mean"], 5e-324) <= 1


def lib_value(orc, v):
    out = _lib.Value()
    out.ct, out.bits = v.ct, v.bits
    return out


@pytest.mark.parametrize("dt", ALL)
def test_library_host_plan_and_finish_match_oracle(orc, dt):
    """ec_statistics_plan / ec_statistics_finish are host code: checked here without a GPU, on oracle-made raw sums."""
    lib = _lib.lib()
    rng = np.random.default_rng(11)
    for kind_ in ("full", "narrow"):
        a = sample(rng, dt, 4000, kind_)
        m = rng.random(4000) < 0.9
        mn, mx = orc.min_max(a, m)
        okind, p, e = orc.statistics_plan(mn, mx)
        lmn, lmx = lib_value(orc, mn), lib_value(orc, mx)
        kind, piv, ex = C.c_int(), C.c_double(), C.c_int()
        _lib.check(lib.ec_statistics_plan(C.byref(lmn), C.byref(lmx), C.byref(kind), C.byref(piv), C.byref(ex)))
        assert (kind.value, bits(piv.value), ex.value) == (okind, bits(p), e)
        parts = [orc.moments_raw(a[i:j], m[i:j], p, e) for i, j in ((0, 1500), (1500, 4000))]
        words = (C.c_uint64 * (9 * len(parts)))()
        for i, r in enumerate(parts):
            words[9 * i] = r[0]
            for k in range(4):
                v = r[1 + k] & ((1 << 128) - 1)
                words[9 * i + 1 + 2 * k], words[9 * i + 2 + 2 * k] = v & (2 ** 64 - 1), v >> 64
        out = _lib.Statistics()
        _lib.check(lib.ec_statistics_finish(words, len(parts), C.byref(lmn), C.byref(lmx), C.byref(out)))
        want = orc.statistics(a, m)
        assert out.count == want["count"]
        assert (bits(out.mean), bits(out.stddev)) == (bits(want["mean"]), bits(want["stddev"]))


def test_library_host_finish_degenerate(orc):
    lib = _lib.lib()
    for arr, mask in ((np.array([], dtype=np.float32), None), (np.array([1, np.inf], dtype=np.float64), None),
                      (np.array([-np.inf, np.inf, np.nan], dtype=np.float32), None),
                      (np.array([3, 4], dtype=np.uint16), np.array([0, 0], dtype=bool))):
        want = orc.statistics(arr, mask)
        mn, mx = lib_value(orc, want["min"]), lib_value(orc, want["max"])
        words = (C.c_uint64 * 9)(want["count"], *([0] * 8))
        out = _lib.Statistics()
        _lib.check(lib.ec_statistics_finish(words, 1, C.byref(mn), C.byref(mx), C.byref(out)))
        assert out.count == want["count"]
        for got, w in ((out.mean, want["mean"]), (out.stddev, want["stddev"])):
            assert (math.isnan(got) and math.isnan(w)) or got == w
