"""erased_cells_b200 — B200-native per-cell compute path of s22s/erased-cells behind the crate's API.

The product is ``lib/liberased_cells_b200.so`` (hand-written sm_100a kernels behind the C ABI of
``include/erased_cells_b200.h``); this package is the Python mirror of the reference's public types
used by the parity tests and the bench. Nothing here computes cells on the CPU.
"""
from ._lib import EcError, NarrowingError, NoDeviceError, ParseError, build, lib  # noqa: F401
from .api import CellBuffer, CellType, CellValue, Mask, MaskedCellBuffer, NoData, Statistics, from_serde, to_serde  # noqa: F401

import contextlib as _contextlib

ADD, SUB, MUL, DIV = 0, 1, 2, 3
FINISH_HOST, FINISH_PEER, FINISH_NCCL = 0, 1, 2


def init_devices(devices) -> int:
    """One process, several GPUs: bind the library to these CUDA devices (before its first use). From then on a
    CellBuffer / Mask of at least shard_min_cells() cells is kept as row strips, one per device, behind the same
    objects and operators; reductions finish across the GPUs. The same device may be listed twice (tests)."""
    import ctypes as C

    from ._lib import check, lib
    arr = (C.c_int * len(devices))(*[int(d) for d in devices])
    check(lib().ec_init_devices(arr, len(devices)))
    return lib().ec_device_count()


def device_count() -> int:
    from ._lib import lib
    return lib().ec_device_count()


def set_shard_min_cells(cells: int) -> int:
    """threshold from which new buffers / masks are sharded; returns the previous one"""
    from ._lib import lib
    return lib().ec_set_shard_min_cells(int(cells))


def set_host_copy_threads(threads: int) -> int:
    """Host threads that move a pageable array (numpy, a `Vec<T>`) to / from pinned staging while the DMA engine copies
    the previous chunks (`ec_set_host_copy_threads`); 0 leaves pageable copies to the CUDA driver. Returns the previous
    setting."""
    return lib().ec_set_host_copy_threads(int(threads))


def set_shard_finish(mode: int) -> int:
    """FINISH_HOST (host folds the strips' partials), FINISH_PEER (GPU-to-GPU exchange inside the reduction kernels) or
    FINISH_NCCL (ncclAllReduce); returns the previous mode"""
    from ._lib import lib
    prev = lib().ec_set_shard_finish(int(mode))
    if prev < 0:
        raise ValueError("shard finish mode")
    return prev


@_contextlib.contextmanager
def lazy(on: bool = True, jit: bool = False):
    """Defer buffer arithmetic inside the block so that op chains fuse into single passes over HBM
    (`(a - b) / (a + b)`, `(a op b) op scalar`); results are bit-identical to eager evaluation.
    jit=True compiles longer chains (e.g. EVI) into one kernel specialised at run time with NVRTC (cached by shape)."""
    from ._lib import check, lib
    prev = lib().ec_get_lazy()
    check(lib().ec_set_lazy((3 if jit else 1) if on else 0))
    try:
        yield
    finally:
        check(lib().ec_set_lazy(prev))

__all__ = ["CellBuffer", "CellType", "CellValue", "Mask", "MaskedCellBuffer", "NoData", "Statistics", "NarrowingError",
           "NoDeviceError", "EcError", "ParseError", "build", "lib", "lazy", "ADD", "SUB", "MUL", "DIV", "init_devices", "device_count", "set_shard_min_cells",
           "set_host_copy_threads", "set_shard_finish", "FINISH_HOST", "FINISH_PEER", "FINISH_NCCL"]
