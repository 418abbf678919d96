"""One launch of two run-time specialised kernels (EVI, SAVI on 16384^2 u16 bands), for ncu:
  ncu --set full --clock-control none -k regex:ecj_kernel --csv --page raw --log-file gpurun_out/jit_full_raw.csv python tools/profile_jit.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import erased_cells_b200 as ec
from erased_cells_b200 import CellType as T, synth

N = 16384 * 16384
nir, red, blue = [synth.device(T.UInt16, N, 0xEC60 + i, kind=synth.INT_RANGE, lo=100, hi=40000) for i in range(3)]
with ec.lazy(jit=True):
    evi = ((nir - red) * 2.5) / (((nir + red * 6.0) - blue * 7.5) + 1.0)
    evi.device_ptr()
    print("EVI", ec.lib().ec_last_kernel().decode())
    del evi
    savi = ((nir - red) * 1.5) / ((nir + red) + 0.5)
    savi.device_ptr()
    print("SAVI", ec.lib().ec_last_kernel().decode())
ec.lib().ec_synchronize()
print("ok")
