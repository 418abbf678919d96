"""erased_cells_b200 — B200-native per-cell compute path of s22s/erased-cells behind the crate's API.

The product is ``lib/liberased_cells_b200.so`` (hand-written sm_100a kernels behind the C ABI of
``include/erased_cells_b200.h``); this package is the Python mirror of the reference's public types
used by the parity tests and the bench. Nothing here computes cells on the CPU.
"""
from ._lib import EcError, NarrowingError, NoDeviceError, ParseError, build, lib  # noqa: F401
from .api import CellBuffer, CellType, CellValue, Mask, MaskedCellBuffer, NoData, Statistics, from_serde, to_serde  # noqa: F401

import contextlib as _contextlib

ADD, SUB, MUL, DIV = 0, 1, 2, 3


@_contextlib.contextmanager
def lazy(on: bool = True, vm: bool = False, jit: bool = False):
    """Defer buffer arithmetic inside the block so that op chains fuse into single passes over HBM
    (`(a - b) / (a + b)`, `(a op b) op scalar`); results are bit-identical to eager evaluation.
    jit=True compiles longer chains (e.g. EVI) into one kernel specialised at run time with NVRTC (cached by shape);
    vm=True routes them through the experimental expression VM instead (one interpreted pass, slower)."""
    from ._lib import check, lib
    prev = lib().ec_get_lazy()
    check(lib().ec_set_lazy((3 if jit else 2 if vm else 1) if on else 0))
    try:
        yield
    finally:
        check(lib().ec_set_lazy(prev))

__all__ = ["CellBuffer", "CellType", "CellValue", "Mask", "MaskedCellBuffer", "NoData", "Statistics", "NarrowingError",
           "NoDeviceError", "EcError", "ParseError", "build", "lib", "lazy", "ADD", "SUB", "MUL", "DIV"]
