"""Config 1 (README example scaled: 4096^2 u8 / u16 * 0.5 -> f64) timed three ways (development tool): through the Python
mirror, through bare C-ABI calls in a tight loop, and the same loop on tiny buffers (= host cost per op), so a
host-bound number is not mistaken for kernel time. CUDA events on the launching stream."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import erased_cells_b200 as ec
from erased_cells_b200 import CellType as T, synth

L = ec.lib()
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
ec._lib.check(L.ec_set_stream(C.c_void_p(st.cuda_stream)))
N = int(os.environ.get("EC_CELLS", 4096 * 4096))
ITERS = 64


def sets(n):
    return [(synth.device(T.UInt8, n, 0xEC01 + 16 * i, kind=synth.INT_RANGE, lo=0, hi=255),
             synth.device(T.UInt16, n, 0xEC02 + 16 * i, kind=synth.INT_RANGE, lo=0, hi=65535)) for i in range(8)]


def timed(fn, iters=ITERS):
    for i in range(8):
        fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3  # us


def raw_div(S):
    h = C.c_void_p()
    binary, free, ref = L.ec_buf_binary, L.ec_buf_free, C.byref(h)
    hs = [(a._h, b._h) for a, b in S]

    def f(i):
        a, b = hs[i & 7]
        binary(3, a, b, ref)
        free(h)
    return f


def raw_fused(S):
    h = C.c_void_p()
    half = ec.CellValue(T.Float64, 0.5)._v
    fn, free, ref, hp = L.ec_buf_binary_scalar, L.ec_buf_free, C.byref(h), C.byref(half)
    hs = [(a._h, b._h) for a, b in S]

    def f(i):
        a, b = hs[i & 7]
        fn(3, a, b, 2, hp, ref)
        free(h)
    return f


for overlap in (1, 0):
    L.ec_set_launch_overlap(overlap)
    big, tiny = sets(N), sets(1024)
    rows = [("python mirror  a / b", timed(lambda i: big[i & 7][0] / big[i & 7][1]), 11),
            ("C ABI loop     a / b", timed(raw_div(big)), 11),
            ("C ABI loop     (a / b) * 0.5 fused", timed(raw_fused(big)), 11),
            ("python mirror  a / b * 0.5 unfused", timed(lambda i: big[i & 7][0] / big[i & 7][1] * 0.5), 27),
            ("C ABI loop     a / b, 1024 cells (host cost per op)", timed(raw_div(tiny)), 0),
            ("python mirror  a / b, 1024 cells (host cost per op)", timed(lambda i: tiny[i & 7][0] / tiny[i & 7][1]), 0)]
    rows += [("python mirror  u16 / 10000.0", timed(lambda i: big[i & 7][1] / 10000.0), 10),
             ("python mirror  u16 / 1e300", timed(lambda i: big[i & 7][1] / 1e300), 10),
             ("python mirror  u16 * 0.0001", timed(lambda i: big[i & 7][1] * 0.0001), 10)]
    print(f"launch overlap {overlap}, {N} cells, {ITERS} back-to-back ops")
    for name, us, bpc in rows:
        print(f"  {name:56s} {us:8.2f} us" + (f"  {bpc * N / us / 1e3:8.0f} GB/s" if bpc else ""))
    del big, tiny
