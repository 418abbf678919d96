"""Host<->device copy bandwidth on this box (development tool): pinned vs pageable, each way and both at once."""
import time

import torch

dev = torch.device("cuda", 0)
n = 1 << 30  # 1 GiB
d0 = torch.empty(n, dtype=torch.uint8, device=dev)
d1 = torch.empty(n, dtype=torch.uint8, device=dev)
hp0 = torch.empty(n, dtype=torch.uint8).pin_memory()
hp1 = torch.empty(n, dtype=torch.uint8).pin_memory()
hpage = torch.empty(n, dtype=torch.uint8)
hpage.fill_(1)


def t(fn, iters=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / iters


s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
print("H2D pinned   GB/s", n / t(lambda: d0.copy_(hp0, non_blocking=True)) / 1e9)
print("D2H pinned   GB/s", n / t(lambda: hp0.copy_(d0, non_blocking=True)) / 1e9)
print("H2D pageable GB/s", n / t(lambda: d0.copy_(hpage)) / 1e9)
print("D2H pageable GB/s", n / t(lambda: hpage.copy_(d0)) / 1e9)


def both():
    with torch.cuda.stream(s1):
        d0.copy_(hp0, non_blocking=True)
    with torch.cuda.stream(s2):
        hp1.copy_(d1, non_blocking=True)


print("H2D+D2H concurrent, GB/s each way", n / t(both) / 1e9)
for sz in (1 << 20, 1 << 24, 1 << 26, 1 << 28):
    print(f"D2H pinned {sz >> 20} MiB chunks GB/s", sz / t(lambda: hp0[:sz].copy_(d0[:sz], non_blocking=True), 20) / 1e9)
import subprocess
print(subprocess.run("nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.gen.max,pcie.link.width.current --format=csv; nvidia-smi topo -m | head -5; numactl -H 2>/dev/null | head -5; lscpu | grep -i -E 'numa|socket|model name'", shell=True, capture_output=True, text=True).stdout)

# ---- the same copies through the C ABI (ec_buf_from_host / ec_buf_to_host) ------------------------------
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import erased_cells_b200 as ec
from erased_cells_b200 import CellBuffer, CellType

L = ec.lib()
cells = 8192 * 8192


def pinned(nbytes):
    p = C.c_void_p()
    ec._lib.check(L.ec_host_alloc(nbytes, C.byref(p)))
    return np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(p.value))


for label, mk in (("ec_host_alloc", pinned), ("torch pin_memory", lambda nb: torch.empty(nb, dtype=torch.uint8).pin_memory().numpy())):
    src = mk(cells * 8).view(np.float64)
    dst = mk(cells * 8).view(np.float64)
    src[:] = 1.0
    for rep in range(2):
        t0 = time.perf_counter(); b = CellBuffer.from_vec(src); t1 = time.perf_counter()
        o = b.to_vec(out=dst); t2 = time.perf_counter()
        c = b.convert(CellType.Float64); L.ec_synchronize(); t3 = time.perf_counter()
        o = c.to_vec(out=dst); t4 = time.perf_counter()
        del c; t5 = time.perf_counter()
        print(f"{label} rep{rep}: H2D {cells * 8 / (t1 - t0) / 1e9:.1f} GB/s, D2H {cells * 8 / (t2 - t1) / 1e9:.1f} GB/s, clone {(t3 - t2) * 1e3:.2f} ms, "
              f"D2H of fresh buffer {cells * 8 / (t4 - t3) / 1e9:.1f} GB/s, free {(t5 - t4) * 1e3:.3f} ms")
    del b
# alternating sizes, as the sweep does
srcs = {ct: CellBuffer.with_defaults(cells, ct) for ct in (CellType.UInt8, CellType.UInt32, CellType.Float64)}
big = pinned(cells * 8)
for rep in range(2):
    for ct, b in srcs.items():
        for d in (CellType.UInt8, CellType.UInt32, CellType.Float64):
            if not ct.can_fit_into(d):
                continue
            t0 = time.perf_counter(); c = b.convert(d); L.ec_synchronize(); t1 = time.perf_counter()
            c.to_vec(out=big[: cells * d.size_of()].view(d.dtype)); t2 = time.perf_counter()
            del c; t3 = time.perf_counter()
            print(f"rep{rep} {ct}->{d}: convert {1e3 * (t1 - t0):.2f} ms, D2H {cells * d.size_of() / (t2 - t1) / 1e9:.1f} GB/s, free {1e3 * (t3 - t2):.3f} ms")
