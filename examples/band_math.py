"""Band math on device-resident rasters: NDVI and EVI through the ordinary operators.

Eagerly every operator is one pass over HBM; inside `ec.lazy()` a chain is evaluated on first use — `(a-b)/(a+b)` and
`(a op b) op s` by precompiled fused kernels, anything longer (with jit=True) by one kernel specialised at run time.
Results are bit-identical in every mode. Statistics of the valid cells (an extension: the reference crate stops at
min_max) come from one more pass.

  python examples/band_math.py [side]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # run from a checkout
import time

import erased_cells_b200 as ec
from erased_cells_b200 import CellType, MaskedCellBuffer, NoData, synth

side = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
n = side * side
nodata = NoData.new(CellType.UInt16, 0)
nir, red, blue = [MaskedCellBuffer.from_buffer_with_nodata(
    synth.device(CellType.UInt16, n, 0xBA0 + i, kind=synth.INT_RANGE, lo=200, hi=40000, period=500, sentinel=0), nodata) for i in range(3)]


def ndvi():
    return (nir - red) / (nir + red)


def evi():
    return ((nir - red) * 2.5) / (((nir + red * 6.0) - blue * 7.5) + 1.0)


for name, fn in (("NDVI", ndvi), ("EVI", evi)):
    eager = fn()
    for label, mode in (("eager", None), ("lazy", dict()), ("lazy + jit", dict(jit=True))):
        for attempt in range(2):  # the second round is the steady state (kernels built, allocator warm)
            t0 = time.perf_counter()
            if mode is None:
                r = fn()
            else:
                with ec.lazy(**mode):
                    r = fn()
                    r.buffer().device_ptr()
            ec.lib().ec_synchronize()
            dt = time.perf_counter() - t0
        assert r == eager
        print(f"{name} {side}x{side} {label:11s} {dt * 1e3:8.3f} ms")
    st = eager.statistics()
    print(f"{name}: valid {st.count} of {n}, min {st.min.value():.6f} max {st.max.value():.6f} mean {st.mean:.6f} stddev {st.stddev:.6f}")
