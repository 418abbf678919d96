"""Where does the e2e convert sweep spend its time? (development tool)"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import erased_cells_b200 as ec
from erased_cells_b200 import CellBuffer, CellType

L = ec.lib()
cells = 8192 * 8192
SZ = [1, 2, 4, 8, 1, 2, 4, 8, 4, 8]


def pinned(nbytes):
    p = C.c_void_p()
    ec._lib.check(L.ec_host_alloc(nbytes, C.byref(p)))
    return np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(p.value))


hsrc = [pinned(cells * SZ[ct]).view(CellType(ct).dtype) for ct in range(10)]
for a in hsrc:
    a[:] = 1
out = pinned(cells * 8)
for step in range(3):
    t = dict(h2d=0.0, conv=0.0, d2h=0.0, free=0.0, err=0.0)
    T0 = time.perf_counter()
    for s in range(10):
        t0 = time.perf_counter(); buf = CellBuffer.from_vec(hsrc[s]); t["h2d"] += time.perf_counter() - t0
        for d in range(10):
            if CellType(s).can_fit_into(CellType(d)):
                t0 = time.perf_counter(); c = buf.convert(CellType(d)); L.ec_synchronize(); t1 = time.perf_counter()
                c.to_vec(out=out[: cells * SZ[d]].view(CellType(d).dtype)); t2 = time.perf_counter()
                del c; t3 = time.perf_counter()
                t["conv"] += t1 - t0; t["d2h"] += t2 - t1; t["free"] += t3 - t2
            else:
                t0 = time.perf_counter()
                try:
                    buf.convert(CellType(d))
                except ec.NarrowingError:
                    pass
                t["err"] += time.perf_counter() - t0
    print(f"step {step}: total {1e3 * (time.perf_counter() - T0):.1f} ms", {k: round(v * 1e3, 1) for k, v in t.items()})
