"""Times the statistics extension on the device: moments pass alone and the whole ec_buf_statistics
(min_max + moments), per cell type, masked and unmasked. Wall clock around synchronised batches of launches."""
import sys
import time

sys.path.insert(0, ".")
import numpy as np  # noqa: E402

from erased_cells_b200 import CellType, MaskedCellBuffer, NoData, sharding, synth  # noqa: E402
from erased_cells_b200._lib import check, lib  # noqa: E402

n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 28)
reps = 10
print(f"cells = {n}")
for ct in CellType:
    buf = synth.device(ct, n, 0xEC90 + int(ct), kind=synth.REAL_RANGE, lo=5.0, hi=120.0)
    valid = synth.device(CellType.UInt8, n, 0xEC9F, kind=synth.INT_RANGE, lo=0, hi=9)
    mb = MaskedCellBuffer.from_buffer_with_nodata(valid, NoData.new(CellType.UInt8, 0))
    mask = mb.mask()
    st = buf.statistics()
    kind, p, e = sharding.statistics_plan(st.min, st.max)
    sz = ct.dtype.itemsize
    for label, m, bpc in (("unmasked", None, sz), ("masked", mask, sz + 0.125)):
        for name, fn in (("moments", lambda: sharding.moments(buf, m, p, e)),
                         ("min_max", lambda: (MaskedCellBuffer(buf, m).min_max() if m is not None else buf.min_max())),
                         ("statistics", lambda: (MaskedCellBuffer(buf, m).statistics() if m is not None else buf.statistics()))):
            fn()
            check(lib().ec_synchronize())
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            check(lib().ec_synchronize())
            dt = (time.perf_counter() - t0) / reps
            print(f"{ct.name:8s} {label:9s} {name:10s} {dt * 1e3:8.3f} ms  {n * bpc / dt / 1e9:8.1f} GB/s (one read of the cells)")
    del buf, valid, mb, mask
