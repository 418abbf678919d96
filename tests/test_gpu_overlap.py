"""Launch overlap (programmatic dependent launch, ec_set_launch_overlap): every op of the reference's API is its own
launch, and a dependent chain `a / b * 0.5` (README.md:28, src/buffer.rs:321-352) hands each result to the next kernel.
With overlap on, a kernel's CTAs may become resident while the previous grid drains and must not touch memory before
it has completed. These tests run dependent chains — results feeding the next op, freed blocks re-used by the caching
allocator as the next output — with overlap on and off and against the oracle: all three must agree bit for bit."""
import numpy as np
import pytest

import erased_cells_b200 as ec
from erased_cells_b200 import CellBuffer, CellType, MaskedCellBuffer, NoData, synth

pytestmark = pytest.mark.gpu


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view({1: "u1", 2: "u2", 4: "u4", 8: "u8"}[a.dtype.itemsize])


@pytest.fixture
def overlap():
    L = ec.lib()
    prev = L.ec_set_launch_overlap(1)
    assert prev in (0, 1)
    yield L
    L.ec_set_launch_overlap(prev)


def chain(a: CellBuffer, b: CellBuffer, rounds: int):
    """dependent ops; every temporary is dropped at once, so its block is the next op's output"""
    x = a - b
    for i in range(rounds):
        y = x * 1.0001          # reads x
        x = y + a               # reads y; the old x goes back to the allocator and is re-used
        y = x / b               # x/0 and 0/0 cells: the NaN rule path
        x = y - x
    m = -x
    return x, m


# a few CTAs (both grids resident at once), several waves, and a ragged multi-wave size
@pytest.mark.parametrize("n", [1000, 16384 * 3 + 5, (1 << 22) + 777, (1 << 25) + 4099])
def test_dependent_chain_on_off_identical(overlap, orc, n):
    L = overlap
    ha = synth.host(CellType.UInt16, n, 0x0A01, kind=synth.INT_RANGE, lo=0, hi=9)
    hb = synth.host(CellType.UInt8, n, 0x0A02, kind=synth.INT_RANGE, lo=0, hi=3)
    a, b = CellBuffer.from_vec(ha), CellBuffer.from_vec(hb)
    res = {}
    for mode in (0, 1, 1, 0, 1):
        L.ec_set_launch_overlap(mode)
        x, m = chain(a, b, 6)
        mn, mx = x.min_max()
        got = (bits(x.to_vec()).copy(), bits(m.to_vec()).copy(), mn.bits, mx.bits)
        if mode in res:
            ref = res[mode]
        else:
            res[mode] = ref = got
        assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1]) and got[2:] == ref[2:]
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1]) and res[0][2:] == res[1][2:]
    # and the oracle, one round (the CPU path is slow): same bits
    if n <= (1 << 22) + 777:
        L.ec_set_launch_overlap(1)
        x, _ = chain(a, b, 1)
        w = orc.tight_binary(orc.SUB, ha, hb)
        y = orc.tight_scalar(orc.MUL, w, orc.value(orc.Float64, 1.0001))
        w = orc.tight_binary(orc.ADD, y, ha)
        y = orc.tight_binary(orc.DIV, w, hb)
        w = orc.tight_binary(orc.SUB, y, w)
        assert np.array_equal(bits(x.to_vec()), bits(w))


def test_convert_ladder_and_masked_chain(overlap, orc):
    L = overlap
    n = (1 << 23) + 33
    h = synth.host(CellType.UInt8, n, 0x0B01)
    src = CellBuffer.from_vec(h)
    out = {}
    for mode in (0, 1):
        L.ec_set_launch_overlap(mode)
        # widening ladder: each cast reads the previous cast's output
        c = src.convert(CellType.UInt16).convert(CellType.UInt32).convert(CellType.UInt64).convert(CellType.Float64)
        # masked: mask build -> masked binary (mask AND fused) -> scalar -> min_max / counts -> convert + NoData fill
        ma = MaskedCellBuffer.from_buffer_with_nodata(src, NoData.default(CellType.UInt8))
        mb = MaskedCellBuffer.from_buffer_with_nodata(src.convert(CellType.Int16), NoData.new(CellType.Int16, 7))
        r = (ma - mb) * 0.5 + ma
        mn, mx = r.min_max()
        filled = r.to_vec_with_nodata(NoData.new(CellType.Float64, -1.0))
        out[mode] = (bits(c.to_vec()).copy(), bits(filled).copy(), r.mask().to_vec().copy(), r.counts(), mn.bits, mx.bits)
    for g, w in zip(out[1], out[0]):
        assert np.array_equal(g, w) if isinstance(g, np.ndarray) else g == w
    assert np.array_equal(out[1][0], bits(h.astype(np.float64)))
    wm = orc.mask_and(orc.mask_from_nodata(h, orc.ND_DEFAULT), h != 7)
    assert np.array_equal(out[1][2], wm)
    assert out[1][3] == orc.mask_counts(wm)


def test_overlap_switch_reports_previous(overlap):
    L = overlap
    assert L.ec_set_launch_overlap(0) == 1
    assert L.ec_set_launch_overlap(0) == 0
    assert L.ec_set_launch_overlap(1) == 0
