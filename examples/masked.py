"""The reference's examples/masked.rs on the device: masks are packed bits in HBM and propagate through the operators."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # run from a checkout
import numpy as np

from erased_cells_b200 import CellType, Mask, MaskedCellBuffer

# the numbers 0..=3 with mask [true, false, true, false]
buf = MaskedCellBuffer.fill_with_mask_via(4, lambda i: (float(i), i % 2 == 0), CellType.Float64)
assert buf.mask() == Mask.new([True, False, True, False])
assert buf.counts() == (2, 2)

ones = MaskedCellBuffer.from_vec(np.ones(4))
r = (buf + ones) * 2.0
expected = MaskedCellBuffer(MaskedCellBuffer.from_vec(np.array([2.0, 4.0, 6.0, 8.0])).buffer(), Mask.new([True, False, True, False]))
assert r == expected
print(r.to_vec(), r.mask().to_vec(), r.statistics())
