"""Synthetic rasters: counter-based splitmix64(seed ^ index), identical on host (numpy, here) and on
the device (ec_buf_synth -> synth_kernel in csrc/ec_tu_reduce.cu). Used by tests and bench only."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import Value, check, lib
from .api import CellBuffer, CellType, CellValue

FULL_BITS, INT_RANGE, REAL_RANGE = 0, 1, 2
_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M
        x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M
        x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M
        return x ^ (x >> np.uint64(31))


def host(ct: CellType, n: int, seed: int, index_offset: int = 0, kind: int = FULL_BITS, lo: float = 0.0,
         hi: float = 0.0, period: int = 0, sentinel=None) -> np.ndarray:
    dt = CellType(ct).dtype
    idx = np.arange(index_offset, index_offset + n, dtype=np.uint64)
    h = splitmix64(np.uint64(seed) ^ idx)
    if kind == FULL_BITS:
        v = h.astype({1: "u1", 2: "u2", 4: "u4", 8: "u8"}[dt.itemsize]).view(dt)  # low bytes of h
    elif kind == INT_RANGE:
        a, b = int(lo), int(hi)
        span = np.uint64(b - a + 1)
        x = np.int64(a) + ((h >> np.uint64(11)) % span).astype(np.int64)
        v = x.astype(dt)
    else:
        u = (h >> np.uint64(11)).astype(np.float64) * 2.0 ** -53
        x = np.float64(lo) + u * (np.float64(hi) - np.float64(lo))
        v = x.astype(dt) if dt.kind == "f" else x.astype(np.int64).astype(dt)
    if sentinel is not None and period:
        v = v.copy()
        v[splitmix64(h) % np.uint64(period) == 0] = sentinel
    return v


def device(ct: CellType, n: int, seed: int, index_offset: int = 0, kind: int = FULL_BITS, lo: float = 0.0,
           hi: float = 0.0, period: int = 0, sentinel=None) -> CellBuffer:
    h = C.c_void_p()
    sv = CellValue(ct, sentinel)._v if sentinel is not None else None
    check(lib().ec_buf_synth(int(ct), n, seed, index_offset, kind, float(lo), float(hi), period,
                             C.byref(sv) if sv is not None else None, C.byref(h)))
    return CellBuffer._take(h)
