"""from_vec / to_vec on PAGEABLE host memory (a numpy array here, a Vec<T> in the crate) against the link:
   link           bare pinned cudaMemcpyAsync of the same bytes, each way
   driver         ec_set_host_copy_threads(0): cudaMemcpy on pageable memory (the driver's own single-thread staging)
   staged, T      8 MiB chunks through pinned staging, T host threads moving them while the DMA engine copies
to_vec is timed into a fresh numpy allocation (np.empty: pages fault in on first touch; numpy itself asks for huge pages),
into an array that has been written before, and into a fresh anonymous mmap (what Vec::with_capacity gets from malloc:
no huge-page advice unless the library gives it, $EC_HOST_COPY_THP=0 turns that off). Every transfer is compared with the source bit for bit.
Usage: python tools/host_copy_probe.py [threads ...]; $EC_SIDE (default 32768), $EC_DEVICES for several GPUs."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import erased_cells_b200 as ec
from erased_cells_b200 import CellBuffer, CellType, synth

side = int(os.environ.get("EC_SIDE", 32768))
n = side * side
L = ec.lib()
if os.environ.get("EC_DEVICES"):
    ec.set_shard_min_cells(1 << 20)
else:
    ec._lib.check(L.ec_init(0))
band = np.empty(n, dtype=np.uint16)
for o in range(0, n, 1 << 26):
    k = min(1 << 26, n - o)
    band[o:o + k] = synth.host(CellType.UInt16, k, 0xEC51, index_offset=o, kind=synth.INT_RANGE, lo=0, hi=40000)
nbytes = band.nbytes
pinned = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
pinned.numpy().view(np.uint16)[:] = band
dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")


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


h2d = best(lambda: dev.copy_(pinned, non_blocking=True))
d2h = best(lambda: pinned.copy_(dev, non_blocking=True))
print(f"link, pinned, {nbytes / 1e9:.2f} GB: H2D {nbytes / h2d / 1e9:.1f} GB/s, D2H {nbytes / d2h / 1e9:.1f} GB/s; "
      f"logical devices: {ec.device_count()}, host cores: {os.cpu_count()}")
try:
    print("transparent huge pages:", open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip(),
          "| EC_HOST_COPY_THP =", os.environ.get("EC_HOST_COPY_THP", "1"), "| EC_HOST_COPY_STREAM =", os.environ.get("EC_HOST_COPY_STREAM", "1"))
except OSError:
    pass
resident = CellBuffer.from_vec(pinned.numpy().view(np.uint16))
print("shards:", resident.shard_count())
touched = np.zeros(n, dtype=np.uint16)


def to_fresh_mmap():  # anonymous pages straight from the kernel, no huge-page advice by the allocator: Vec::with_capacity
    import mmap
    m = mmap.mmap(-1, nbytes, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
    out = np.frombuffer(m, dtype=np.uint16)
    resident.to_vec(out=out)
    return out


for threads in [int(a) for a in sys.argv[1:]] or [0, -1, 1, 4, 8, 12, 16]:
    ec.set_host_copy_threads(threads)
    up = best(lambda: CellBuffer.from_vec(band))
    down_fresh = best(lambda: resident.to_vec())
    down = best(lambda: resident.to_vec(out=touched))
    down_mmap = best(to_fresh_mmap)
    b = CellBuffer.from_vec(band)
    assert b == resident and np.array_equal(resident.to_vec(), band) and np.array_equal(touched, band), "staged copy differs"
    label = "driver (pageable cudaMemcpy)" if threads == 0 else "staged, default threads" if threads < 0 else f"staged, {threads} thread{'s' if threads > 1 else ''}"
    print(f"{label}: from_vec {nbytes / up / 1e9:.1f} GB/s ({h2d / up:.2f} of the link) | to_vec into fresh memory "
          f"{nbytes / down_fresh / 1e9:.1f} GB/s ({d2h / down_fresh:.2f}) | to_vec into touched memory {nbytes / down / 1e9:.1f} GB/s ({d2h / down:.2f}) "
          f"| to_vec into a fresh plain mmap {nbytes / down_mmap / 1e9:.1f} GB/s ({d2h / down_mmap:.2f})")
