"""Property-based check of the product's host-side scalar logic (ec_ctype_* / ec_value_*) against the oracle:
arbitrary bit patterns of every cell type through binary ops, neg, cmp, convert and the value-checked to_<p>()."""
import numpy as np
from hypothesis import given, settings, strategies as st

from erased_cells_b200 import CellType, CellValue

cts = st.sampled_from(list(CellType))
bits64 = st.integers(min_value=0, max_value=2**64 - 1)


def mk(ct, b):
    a = np.array([b], dtype="<u8").view(ct.dtype)
    return a[0]


@settings(max_examples=400, deadline=None)
@given(cts, bits64, cts, bits64, st.integers(0, 3))
def test_value_binary_cmp(orc, lct, lb, rct, rb, op):
    x, y = mk(lct, lb), mk(rct, rb)
    a, b = CellValue(lct, x), CellValue(rct, y)
    oa, ob = orc.value(int(lct), x), orc.value(int(rct), y)
    got, want = a._bin(op, b), orc.value_binary(op, oa, ob)
    assert (int(got.cell_type()), got.bits) == want.key()
    assert a.cmp(b) == orc.value_cmp(oa, ob)


@settings(max_examples=400, deadline=None)
@given(cts, bits64, cts)
def test_value_unary_and_casts(orc, ct, b, dst):
    x = mk(ct, b)
    a, oa = CellValue(ct, x), orc.value(int(ct), x)
    n, on = -a, orc.value_neg(oa)
    assert (int(n.cell_type()), n.bits) == on.key()
    got, want = a.to_prim(dst), orc.value_to_prim(oa, int(dst))
    assert (got is None) == (want is None)
    if got is not None:
        assert (int(got.cell_type()), got.bits) == want.key()
    if ct.can_fit_into(dst):
        c = a.convert(dst)
        assert (int(c.cell_type()), c.bits) == orc.value_convert(oa, int(dst)).key()
