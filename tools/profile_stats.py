"""One launch of each statistics kernel (extension) on 2^28-cell buffers, for ncu.

  python tools/profile_stats.py        # plain run first (must exit 0)
  ncu --set full --clock-control none -k regex:"int_stats_kernel|moments_kernel" --csv --page raw \
      --log-file gpurun_out/stats_full_raw.csv python tools/profile_stats.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import erased_cells_b200 as ec
from erased_cells_b200 import CellType as T, MaskedCellBuffer, NoData, synth

N = int(os.environ.get("EC_PROFILE_CELLS", 1 << 28))
u8 = synth.device(T.UInt8, N, 1, kind=synth.INT_RANGE, lo=0, hi=255)
u16 = synth.device(T.UInt16, N, 2, kind=synth.INT_RANGE, lo=5000, hi=40000)
i16 = synth.device(T.Int16, N, 4, kind=synth.INT_RANGE, lo=-32768, hi=32767, period=50, sentinel=-32768)
u32 = synth.device(T.UInt32, N, 5)
i32 = synth.device(T.Int32, N, 6)
f32 = synth.device(T.Float32, N, 7, kind=synth.REAL_RANGE, lo=-1e4, hi=1e4)
f64 = synth.device(T.Float64, N, 8, kind=synth.REAL_RANGE, lo=-1e4, hi=1e4)
mi16 = MaskedCellBuffer.from_buffer_with_nodata(i16, NoData.default(T.Int16))
ec.lib().ec_synchronize()
for name, b in (("u8", u8), ("u16", u16), ("masked i16", mi16), ("u32", u32), ("i32", i32), ("f32", f32), ("f64", f64)):
    s = b.statistics()
    print(name, s)
print("ok")
