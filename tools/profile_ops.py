"""One launch of each representative kernel of the path, for ncu (tools/ is development tooling).

  python tools/profile_ops.py                # plain run (must exit 0 before profiling)
  ncu --set full --clock-control none --import-source on -o gpurun_out/prof python tools/profile_ops.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import erased_cells_b200 as ec
from erased_cells_b200 import CellType as T, MaskedCellBuffer, NoData, synth

N = int(os.environ.get("EC_PROFILE_CELLS", 8192 * 8192))
u8 = synth.device(T.UInt8, N, 1, kind=synth.INT_RANGE, lo=0, hi=255)
u16 = synth.device(T.UInt16, N, 2, kind=synth.INT_RANGE, lo=0, hi=65535)
u16b = synth.device(T.UInt16, N, 3, kind=synth.INT_RANGE, lo=0, hi=65535)
i16a = synth.device(T.Int16, N, 4, kind=synth.INT_RANGE, lo=-32768, hi=32767, period=50, sentinel=-32768)
i16b = synth.device(T.Int16, N, 5, kind=synth.INT_RANGE, lo=-32768, hi=32767, period=50, sentinel=-32768)
u64 = synth.device(T.UInt64, N, 6)
f32 = synth.device(T.Float32, N, 7, kind=synth.REAL_RANGE, lo=-1e4, hi=1e4)
ec.lib().ec_synchronize()

ops = []
ops.append(("convert u64->f64 (16 B/cell)", lambda: u64.convert(T.Float64)))
ops.append(("convert u8->u16 (3 B/cell)", lambda: u8.convert(T.UInt16)))
ops.append(("convert f32->f64 (12 B/cell)", lambda: f32.convert(T.Float64)))
ops.append(("clone u64 (16 B/cell)", lambda: u64.clone()))
ops.append(("u8 / u16 (11 B/cell)", lambda: u8 / u16))
ops.append(("(u8 / u16) * 0.5 fused (11 B/cell)", lambda: u8.binary_scalar(ec.DIV, u16, ec.MUL, 0.5)))
ops.append(("normalized_difference u16 (12 B/cell)", lambda: u16.normalized_difference(u16b)))
ma = MaskedCellBuffer.from_buffer_with_nodata(i16a, NoData.default(T.Int16))   # mask_build
mb = MaskedCellBuffer.from_buffer_with_nodata(i16b, NoData.default(T.Int16))
ops.append(("masked i16 - i16 (12.375 B/cell)", lambda: ma - mb))
r = ma - mb
ops.append(("f64 * scalar (16 B/cell)", lambda: r.buffer() * 0.0001))
ops.append(("masked min_max f64 (8.125 B/cell)", lambda: r.min_max()))
ops.append(("min_max f32 (4 B/cell)", lambda: f32.min_max()))
ops.append(("min_max u8 (1 B/cell)", lambda: u8.min_max()))
ops.append(("to_vec_with_nodata i16->f32 fill (6.125 B/cell)", lambda: ec.CellBuffer._take(_fill())))
ops.append(("counts (1/8 B/cell)", lambda: r.counts()))


def _fill():
    import ctypes as C
    h = C.c_void_p()
    nd = NoData.new(T.Float32, -9999.0)
    ec._lib.check(ec.lib().ec_buf_fill_nodata(ma.buffer()._h, ma.mask()._h, int(T.Float32), nd.kind, nd._ptr(), C.byref(h)))
    return h


for name, fn in ops:
    before = ec.lib().ec_kernel_launches()
    out = fn()
    ec.lib().ec_synchronize()
    print(f"{name}: {ec.lib().ec_kernel_launches() - before} launch(es), family={ec.lib().ec_last_kernel().decode()}")
    del out
print("ok")
