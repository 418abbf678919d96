"""Throughput of the kernels the headline configs do not touch (development tool): GB/s algorithmic, CUDA events."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import erased_cells_b200 as ec
from erased_cells_b200 import CellBuffer, CellType as T, CellValue, Mask, MaskedCellBuffer, NoData, synth

L = ec.lib()
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
ec._lib.check(L.ec_set_stream(C.c_void_p(st.cuda_stream)))
N = int(os.environ.get("EC_CELLS", 1 << 28))


def timed(fn, iters=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def row(name, ms, bpc):
    print(f"{name:46s} {ms:8.3f} ms  {bpc * N / ms / 1e6:8.0f} GB/s")


def _fill(m, ct):
    h = C.c_void_p()
    nd = NoData.new(ct, 1)
    ec._lib.check(L.ec_buf_fill_nodata(m.buffer()._h, m.mask()._h, int(ct), nd.kind, nd._ptr(), C.byref(h)))
    return h



for ct in (T.UInt8, T.Int8, T.UInt16, T.Int16, T.UInt32, T.Float32, T.Int64, T.Float64):
    a = synth.device(ct, N, 1 + int(ct))
    sz = ct.size_of()
    m = MaskedCellBuffer.from_buffer_with_nodata(a, NoData.new(ct, a.get(5).value()))
    row(f"min_max {ct}", timed(lambda: a.min_max()), sz)
    row(f"masked min_max {ct}", timed(lambda: m.min_max()), sz + 0.125)
    row(f"from_nodata {ct}", timed(lambda: MaskedCellBuffer.from_buffer_with_nodata(a, NoData.default(ct))), sz + 0.125)
    row(f"neg {ct}", timed(lambda: -a), sz + (-a).cell_type().size_of())
    row(f"fill_nodata {ct}->{ct}", timed(lambda: CellBuffer._take(_fill(m, ct))), 2 * sz + 0.125)
    row(f"cmp (equal buffers) {ct}", timed(lambda: a.cmp(a)), 2 * sz)
    row(f"scalar mul {ct}", timed(lambda: a * 0.5), sz + 8)
    row(f"fill {ct}", timed(lambda: CellBuffer.fill(N, CellValue(ct, 3))), sz)
    del a, m


ma, mb = Mask.new(synth.host(T.UInt8, 1 << 20, 1) < 128), None
big = MaskedCellBuffer.from_buffer_with_nodata(synth.device(T.UInt8, N, 5), NoData.new(T.UInt8, 7)).mask()
row("mask and", timed(lambda: big & big), 3 / 8)
row("mask not", timed(lambda: ~big), 2 / 8)
row("mask counts", timed(lambda: big.counts()), 1 / 8)
