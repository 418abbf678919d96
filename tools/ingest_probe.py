"""Ingest + NoData mask of a 32768^2 u16 band (2.1 GB) against the host -> device link:
   link          bare pinned cudaMemcpyAsync H2D of the same bytes
   one shot      from_vec(pinned band) then from_nodata: the copy, then the mask kernel
   chunked       ec_ingest_*: 32 MiB chunks through pinned staging, mask kernel of chunk k - 1 under the copy of chunk k;
                 the "reader" that fills the staging buffers is (a) nothing (buffers pre-filled: the pipeline's own ceiling),
                 (b) one thread copying from a pageable array, (c) 8 threads copying slices of each chunk
Prints GB/s and the fraction of the link; the result of (b) is compared with the one-shot upload bit for bit."""
import ctypes as C
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import erased_cells_b200 as ec
from erased_cells_b200 import CellBuffer, CellType, MaskedCellBuffer, NoData, raster_io, synth

side = int(os.environ.get("EC_SIDE", 32768))
n = side * side
L = ec.lib()
ec._lib.check(L.ec_init(0))
band = synth.host(CellType.UInt16, n, 0xEC50, kind=synth.INT_RANGE, lo=0, hi=40000) if n <= (1 << 26) else None
if band is None:  # 2^30 cells: generate in slices
    band = np.empty(n, dtype=np.uint16)
    for o in range(0, n, 1 << 26):
        band[o:o + (1 << 26)] = synth.host(CellType.UInt16, 1 << 26, 0xEC50, index_offset=o, kind=synth.INT_RANGE, lo=0, hi=40000)
nbytes = band.nbytes
pinned = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
pinned.numpy().view(np.uint16)[:] = band
dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
nd = NoData.new(CellType.UInt16, 0)


def best(fn, reps=3):
    t = []
    for _ in range(reps):
        torch.cuda.synchronize(); ec._lib.check(L.ec_synchronize())
        t0 = time.perf_counter()
        out = fn()
        ec._lib.check(L.ec_synchronize()); torch.cuda.synchronize()
        t.append(time.perf_counter() - t0)
        del out
    return min(t)


link = best(lambda: dev.copy_(pinned, non_blocking=True))
print(f"link: bare pinned H2D of {nbytes / 1e9:.2f} GB: {nbytes / link / 1e9:.1f} GB/s")
src = pinned.numpy().view(np.uint16)


def one_shot():
    b = CellBuffer.from_vec(src, wait=False)
    return MaskedCellBuffer.from_buffer_with_nodata(b, nd)


t = best(one_shot)
print(f"one shot (upload, then mask): {nbytes / t / 1e9:.1f} GB/s = {link / t:.3f} of the link")


def chunked(reader):
    g = raster_io.Ingest(CellType.UInt16, n, nd, masked=True)
    pos = 0
    while True:
        buf = g.next_buffer()
        if buf is None:
            break
        k = min(buf.size, n - pos)
        reader(buf, pos, k)
        g.submit(k)
        pos += k
    return g.finish()


def one_shot_pageable():
    b = CellBuffer.from_vec(band, wait=True)
    return MaskedCellBuffer.from_buffer_with_nodata(b, nd)


t = best(one_shot_pageable)
print(f"one shot from PAGEABLE memory (what from_vec(Vec<T>) does): {nbytes / t / 1e9:.1f} GB/s = {link / t:.3f} of the link")
pool = ThreadPoolExecutor(8)


def copy1(buf, pos, k):
    buf[:k] = band[pos:pos + k]


def copy8(buf, pos, k):
    step = (k + 7) // 8
    list(pool.map(lambda i: buf.__setitem__(slice(i * step, min(k, (i + 1) * step)), band[pos + i * step: pos + min(k, (i + 1) * step)]), range(8)))


for label, reader in (("reader: none (staging pre-filled)", lambda buf, pos, k: None), ("reader: 1 thread copying from pageable memory", copy1),
                      ("reader: 8 threads copying from pageable memory", copy8)):
    t = best(lambda: chunked(reader))
    print(f"chunked ingest + mask, {label}: {nbytes / t / 1e9:.1f} GB/s = {link / t:.3f} of the link")
m = chunked(copy1)
ref = one_shot()
assert m.buffer() == ref.buffer() and m.mask() == ref.mask() and m.counts() == ref.counts(), "chunked ingest differs from the one-shot upload"
print("chunked == one shot: buffer, mask and counts identical;", m.counts())
