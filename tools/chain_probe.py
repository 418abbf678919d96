"""Longer op chains through the three evaluation modes of lazy operators: op by op (mode 1 when no precompiled shape
matches), the expression VM (mode 2) and run-time specialised kernels (mode 3, NVRTC). Wall clock around synchronised
batches; bytes = operands read once + f64 result written once."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import erased_cells_b200 as ec
from erased_cells_b200 import CellType as T, synth

side = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
N = side * side
L = ec.lib()
nir, red, blue = [synth.device(T.UInt16, N, 0xEC60 + i, kind=synth.INT_RANGE, lo=100, hi=40000) for i in range(3)]
f32 = synth.device(T.Float32, N, 0xEC70, kind=synth.REAL_RANGE, lo=-1e4, hi=1e4)

chains = {
    "EVI 2.5*(nir-red)/(nir+6*red-7.5*blue+1), 3 x u16 -> f64 (14 B/cell)": (lambda: ((nir - red) * 2.5) / (((nir + red * 6.0) - blue * 7.5) + 1.0), 14),
    "SAVI 1.5*(nir-red)/(nir+red+0.5), 2 x u16 -> f64 (12 B/cell)": (lambda: ((nir - red) * 1.5) / ((nir + red) + 0.5), 12),
    "(f32*0.0001+273.15)*(f32*0.0001+273.15), 1 x f32 -> f64 (12 B/cell)": (lambda: (f32 * 0.0001 + 273.15) * (f32 * 0.0001 + 273.15), 12),
}
for name, (fn, bpc) in chains.items():
    print(name)
    ref = fn()
    for label, mode in (("eager, op by op", None), ("lazy, precompiled shapes only", dict()), 
                        ("lazy + run-time specialised kernel", dict(jit=True))):
        def run():
            if mode is None:
                return fn()
            with ec.lazy(**mode):
                r = fn()
                r.device_ptr()
            return r
        t0 = time.perf_counter()
        out = run()   # first call: includes the NVRTC build in mode 3
        L.ec_synchronize()
        first = time.perf_counter() - t0
        assert out == ref
        k0 = L.ec_kernel_launches()
        reps = 5
        t0 = time.perf_counter()
        for _ in range(reps):
            out = run()
        L.ec_synchronize()
        dt = (time.perf_counter() - t0) / reps
        print(f"  {label:38s} {dt * 1e3:8.3f} ms  {N * bpc / dt / 1e9:7.0f} GB/s algorithmic  {(L.ec_kernel_launches() - k0) // reps} launches  "
              f"(first call {first * 1e3:.0f} ms)  last={L.ec_last_kernel().decode()}")
        del out
