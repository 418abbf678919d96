"""The reference's examples/quick.rs on the device: same values, same result, the buffers live in HBM."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # run from a checkout
import numpy as np

from erased_cells_b200 import CellBuffer

buf1 = CellBuffer.from_vec(np.array([1, 2, 3], dtype=np.uint8))    # CellBuffer::from(vec![1u8, 2, 3])
buf2 = CellBuffer.from_vec(np.array([2, 4, 6], dtype=np.uint16))   # CellBuffer::from(vec![2u16, 4, 6])
result = buf1 / buf2 * 0.5                                         # division coerces to f64
assert result == CellBuffer.from_vec(np.array([0.25, 0.25, 0.25]))
print(result.cell_type().name, result.to_vec())
