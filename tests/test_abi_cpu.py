"""CPU-side checks of the product: the C-ABI library loads, exports every symbol include/*.h
declares, its host-side CellType/CellValue logic agrees with the oracle, and device entry points
fail loudly (no CPU fallback) when no GPU is present. No device compute here."""
import ctypes as C
import os
import re
import sys

import numpy as np
import pytest

import erased_cells_b200 as ec
from erased_cells_b200 import CellType, CellValue, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "erased_cells_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ec_[a-z0-9_]+)\s*\(", src)))


def test_library_loads_and_exports_every_declared_symbol():
    L = ec.lib()
    assert L.ec_abi_version() == 2
    names = declared_symbols()
    assert len(names) > 80
    for n in names:
        assert hasattr(L, n), f"{n} declared in the header but not exported"
    # and the ctypes table binds exactly the declared set
    assert sorted(_lib._signatures().keys()) == names


def test_value_struct_layout():
    assert C.sizeof(_lib.Value) == 16 and _lib.Value.bits.offset == 8


def test_union_lattice_matches_oracle(orc):
    for a in CellType:
        for b in CellType:
            assert int(a.union(b)) == orc.union(int(a), int(b))
            assert a.can_fit_into(b) == orc.can_fit_into(int(a), int(b))
        assert a.size_of() == orc.size_of(int(a))
        assert a.is_integral() == orc.is_integral(int(a)) and a.is_signed() == orc.is_signed(int(a))
        assert str(a) == orc.name(int(a)) and CellType.from_str(str(a)) == a
        for f, g in ((a.min_value, orc.min_value), (a.max_value, orc.max_value), (a.zero, orc.zero), (a.one, orc.one)):
            v, w = f(), g(int(a))
            assert (int(v.cell_type()), v.bits) == w.key()
    with pytest.raises(ec.ParseError):
        CellType.from_str("UInt57")  # src/ctype.rs:263


def _specials(ct, rng):
    dt = CellType(ct).dtype
    raw = rng.integers(0, 256, size=64 * dt.itemsize, dtype=np.uint8).view(dt).copy()
    if dt.kind == "f":
        raw[:8] = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, -np.nan, 1.5, -2.5], dtype=dt)
    else:
        info = np.iinfo(dt)
        raw[:4] = np.array([info.min, info.max, 0, 1], dtype=dt)
    return raw


def test_scalar_ops_match_oracle(orc):
    rng = np.random.default_rng(42)
    for lct in CellType:
        ls = _specials(lct, rng)
        for rct in CellType:
            rs = _specials(rct, rng)
            for x, y in zip(ls[:24], rs[:24]):
                a, b = CellValue(lct, x), CellValue(rct, y)
                oa, ob = orc.value(int(lct), x), orc.value(int(rct), y)
                for op, got in enumerate((a + b, a - b, a * b, a / b)):
                    want = orc.value_binary(op, oa, ob)
                    assert (int(got.cell_type()), got.bits) == want.key(), (lct, rct, op, x, y)
                assert a.cmp(b) == orc.value_cmp(oa, ob)
        for x in ls:
            a, oa = CellValue(lct, x), orc.value(int(lct), x)
            n, on = -a, orc.value_neg(oa)
            assert (int(n.cell_type()), n.bits) == on.key()
            assert a.to_f64() == orc.value_to_f64(oa) or (np.isnan(a.to_f64()) and np.isnan(orc.value_to_f64(oa)))
            assert a.to_i64() == orc.value_to_i64(oa) and a.to_u64() == orc.value_to_u64(oa)
            for d in CellType:  # value-checked to_<p>() (Extend / GDAL nodata conversion)
                got, want = a.to_prim(d), orc.value_to_prim(oa, int(d))
                assert (got is None) == (want is None), (lct, d, x)
                if got is not None:
                    assert (int(got.cell_type()), got.bits) == want.key(), (lct, d, x)
            for d in CellType:
                if lct.can_fit_into(d):
                    c = a.convert(d)
                    assert (int(c.cell_type()), c.bits) == orc.value_convert(oa, int(d)).key()
                else:
                    with pytest.raises(ec.NarrowingError) as e:
                        a.convert(d)
                    assert (e.value.src, e.value.dst) == (int(lct), int(d))


def test_reference_scalar_kats():
    # src/value.rs:313-329, :338-346, :349-391 through the product's host scalar path
    assert CellValue(CellType.UInt8, 43).convert(CellType.Int16) == CellValue(CellType.Int16, 43)
    with pytest.raises(ec.NarrowingError):
        CellValue(CellType.Float32, 3.11111).convert(CellType.Int32)
    assert CellValue(CellType.UInt16, 33).convert(CellType.Float32).cell_type() == CellType.Float32
    n = -CellValue(CellType.UInt8, 1)
    assert n.cell_type() == CellType.Int16 and n.value() == -1
    l, r = CellValue(CellType.UInt8, 1), CellValue(CellType.UInt8, 2)
    assert l + r == CellValue(CellType.Float64, 3.0) and l + 2 == CellValue(CellType.Float64, 3.0)
    assert (l / r).cell_type() == CellType.Float64 and (l / r).value() == 0.5
    # doc-test src/buffer.rs:36: ((max - min + 1) / 2) == 4.5
    assert ((CellValue(CellType.UInt8, 8) - CellValue(CellType.UInt8, 0) + 1) / 2) == CellValue.new(4.5)


def test_nodata_value_defaults():
    from erased_cells_b200 import NoData
    assert NoData.none(CellType.Int16).value() is None
    assert NoData.default(CellType.UInt8).value() == 0
    assert np.isnan(NoData.default(CellType.Float32).value())
    assert NoData.new(CellType.UInt16, 6).value() == 6
    assert NoData.default(CellType.Int16).value() == -32768
    assert CellValue.new(np.float64("nan")).is_nodata(NoData.default(CellType.Float64))  # src/masked/nodata.rs:92-94


def test_row_strips():
    L = ec.lib()
    off, ln = C.c_size_t(), C.c_size_t()
    cover = 0
    for g in range(8):
        assert L.ec_row_strip(32768, 32768, 8, g, C.byref(off), C.byref(ln)) == 0
        assert off.value == cover and off.value % 128 == 0 and ln.value == 32768 * 4096
        cover += ln.value
    assert cover == 32768 * 32768
    cover = 0
    for g in range(3):  # ragged: remainder rows go to the last strip; unaligned width falls back to cell ranges
        assert L.ec_row_strip(186, 169, 3, g, C.byref(off), C.byref(ln)) == 0
        assert off.value == cover and off.value % 128 == 0
        cover += ln.value
    assert cover == 186 * 169


def test_device_calls_fail_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(ec.NoDeviceError):
        ec.CellBuffer.from_vec(np.arange(4, dtype=np.uint8))
    with pytest.raises(ec.NoDeviceError):
        ec.Mask.fill(4, True)


def test_jit_kernels_build_for_sm100a_without_a_gpu():
    """ec_jit_dry_build: the run-time specialised kernel of a pending chain compiles with NVRTC for sm_100a (no load, no
    launch). Skipped where libnvrtc cannot be loaded (the product then evaluates chains op by op)."""
    import ctypes as C

    import pytest

    from erased_cells_b200 import _lib
    L = _lib.lib()
    log = C.create_string_buffer(4096)
    evi = b"ecj_div(ecj_mul(ecj_sub(v0, v1), c0), ecj_add(ecj_sub(ecj_add(v0, ecj_mul(v1, c1)), ecj_mul(v2, c2)), c3))"
    cts = (C.c_uint8 * 3)(1, 5, 0)
    rc = L.ec_jit_dry_build(cts, 3, 4, evi, log, len(log))
    if rc == _lib.EC_NO_DEVICE:
        pytest.skip("libnvrtc not loadable here")
    assert rc == _lib.EC_OK, log.value.decode()
    for ct in range(10):  # every cell type as operand 0, eight operands of the widest type
        cts = (C.c_uint8 * 2)(ct, 9 - ct)
        assert L.ec_jit_dry_build(cts, 2, 1, b"ecj_mul(ecj_add(v0, v1), c0)", log, len(log)) == _lib.EC_OK, log.value.decode()
    cts = (C.c_uint8 * 8)(*([9] * 8))
    # integer-typed sub-expressions take the guard-free quotient (ecj_divi): (v0 - v1) / (v0 + v1) * c0 over u16 cells
    cts2 = (C.c_uint8 * 2)(1, 1)
    assert L.ec_jit_dry_build(cts2, 2, 1, b"ecj_mul(ecj_divi(ecj_sub(v0, v1), ecj_add(v0, v1)), c0)", log, len(log)) == _lib.EC_OK, log.value.decode()
    expr = b"ecj_add(ecj_add(ecj_add(v0, v1), ecj_add(v2, v3)), ecj_add(ecj_add(v4, v5), ecj_add(v6, v7)))"
    assert L.ec_jit_dry_build(cts, 8, 0, expr, log, len(log)) == _lib.EC_OK, log.value.decode()
    assert L.ec_jit_dry_build(cts, 2, 0, b"not_a_function(v0, v1)", log, len(log)) == _lib.EC_INVALID_ARG and b"not_a_function" in log.value


def test_rust_ffi_and_ctypes_bindings_match_the_header():
    """rust/ffi.rs is generated from include/erased_cells_b200.h (tools/gen_ffi_rs.py): it must be up to date, declare
    every function of the header with the header's arity and types, and the ctypes binding the tests drive must agree with
    the same prototypes — so neither host binding can drift from the C ABI unnoticed."""
    import ctypes as C
    import re
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_ffi_rs as gen
    from erased_cells_b200 import _lib

    header = open(gen.HEADER).read()
    assert open(gen.OUT).read() == gen.generate(), "rust/ffi.rs is stale: python tools/gen_ffi_rs.py"
    protos = gen.prototypes(header)
    assert len(protos) >= 120 and len({n for n, _, _ in protos}) == len(protos)
    rust = open(gen.OUT).read()
    decl = dict(re.findall(r"pub fn (ec_\w+)\((.*?)\)(?: -> [^;]+)?;", rust))
    L = ec.lib()
    sigs = _lib._signatures()
    assert set(sigs) == {n for n, _, _ in protos}, set(sigs) ^ {n for n, _, _ in protos}
    scalar = {"int": C.c_int, "uint8_t": C.c_uint8, "uint64_t": C.c_uint64, "int64_t": C.c_int64, "size_t": C.c_size_t,
              "double": C.c_double, "float": C.c_float, "ec_status": C.c_int}

    def size_class(ctype_c):
        """what the calling convention sees: pointer, or a scalar of a given ctypes type"""
        c = ctype_c.replace(" *", "*").strip()
        if c.endswith("*"):
            return "ptr"
        return scalar[c.replace("const ", "").strip()]

    def py_class(t):
        if t is None:
            return None
        if t in (C.c_void_p, C.c_char_p) or hasattr(t, "contents") or issubclass(t, C._Pointer):
            return "ptr"
        return t

    for name, ret, args in protos:
        assert hasattr(L, name), f"{name} is declared in the header but not exported"
        n_rust = 0 if not decl[name].strip() else decl[name].count(":")
        assert n_rust == len(args), (name, decl[name], args)
        res, argtypes = sigs[name]
        assert len(argtypes) == len(args), (name, argtypes, args)
        for (ctype_c, pname), at in zip(args, argtypes):
            assert size_class(ctype_c) == py_class(at), (name, pname, ctype_c, at)
        assert (None if ret == "void" else size_class(ret)) == py_class(res), (name, ret, res)


def test_every_environment_variable_is_documented():
    """Each EC_* variable the library reads (env_int / getenv in csrc/) appears in INTEGRATION.md's table."""
    import glob
    import re

    used = set()
    for path in glob.glob(os.path.join(ROOT, "erased_cells_b200", "csrc", "*")):
        if path.endswith((".cu", ".cuh", ".hpp", ".inc")):
            used |= set(re.findall(r'(?:env_int|getenv)\("(EC_[A-Z_]+)"', open(path).read()))
    assert len(used) >= 15, used
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    missing = sorted(v for v in used if f"`{v}`" not in doc)
    assert not missing, f"undocumented environment variables: {missing}"
