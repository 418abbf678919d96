"""erased_cells_b200 — B200-native per-cell compute path of s22s/erased-cells behind the crate's API.

The product is ``lib/liberased_cells_b200.so`` (hand-written sm_100a kernels behind the C ABI of
``include/erased_cells_b200.h``); this package is the Python mirror of the reference's public types
used by the parity tests and the bench. Nothing here computes cells on the CPU.
"""
from ._lib import EcError, NarrowingError, NoDeviceError, ParseError, build, lib  # noqa: F401
from .api import CellBuffer, CellType, CellValue, Mask, MaskedCellBuffer, NoData  # noqa: F401

ADD, SUB, MUL, DIV = 0, 1, 2, 3

__all__ = ["CellBuffer", "CellType", "CellValue", "Mask", "MaskedCellBuffer", "NoData", "NarrowingError",
           "NoDeviceError", "EcError", "ParseError", "build", "lib", "ADD", "SUB", "MUL", "DIV"]
