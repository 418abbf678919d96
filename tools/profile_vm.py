"""One launch of the expression VM on EVI (development tool, for ncu)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import erased_cells_b200 as ec
from erased_cells_b200 import CellType as T, synth

N = 8192 * 8192
nir, red, blue = [synth.device(T.UInt16, N, 0xEC60 + i, kind=synth.INT_RANGE, lo=100, hi=40000) for i in range(3)]
for _ in range(2):
    with ec.lazy():
        r = ((nir - red) * 2.5) / (((nir + red * 6.0) - blue * 7.5) + 1.0)
        r.device_ptr()
    ec.lib().ec_synchronize()
print("ok", ec.lib().ec_last_kernel().decode())
