"""ORACLE — TEST INFRASTRUCTURE ONLY. ctypes/numpy front-end of the CPU restatement in
``erased_cells_oracle.hpp`` (which cites the reference file:line it follows).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this
module; the product package ``erased_cells_b200`` never does.

Buffers are numpy arrays whose dtype carries the cell type; scalars are ``(cell_type, python value)``
pairs or ``Value`` structs (16 bytes: tag + 8 payload bytes, the layout of the product's ec_value).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "liberased_cells_oracle.so")

# with_ct! order — /root/reference/src/lib.rs:85-101
UInt8, UInt16, UInt32, UInt64, Int8, Int16, Int32, Int64, Float32, Float64 = range(10)
CELL_TYPES = list(range(10))
NAMES = ["UInt8", "UInt16", "UInt32", "UInt64", "Int8", "Int16", "Int32", "Int64", "Float32", "Float64"]
DTYPES = [np.dtype(d) for d in ("u1", "u2", "u4", "u8", "i1", "i2", "i4", "i8", "f4", "f8")]
ADD, SUB, MUL, DIV = range(4)
OPS = [ADD, SUB, MUL, DIV]
ND_NONE, ND_DEFAULT, ND_VALUE = range(3)
OK, NARROWING, OOB, LEN_MISMATCH = range(4)


def ct_of(arr_or_dtype) -> int:
    dt = np.dtype(arr_or_dtype.dtype if hasattr(arr_or_dtype, "dtype") else arr_or_dtype)
    return DTYPES.index(dt)


class Value(C.Structure):
    _fields_ = [("ct", C.c_uint8), ("pad", C.c_uint8 * 7), ("bits", C.c_uint64)]

    def numpy(self):
        """The payload as a numpy scalar of the tagged type."""
        return np.array([self.bits], dtype="<u8").view(DTYPES[self.ct])[0]

    def key(self):
        return (int(self.ct), int(self.bits))

    def __repr__(self):
        return f"{NAMES[self.ct]}({self.numpy()!r})"


def value(ct: int, x) -> Value:
    """Build a tagged scalar; ``x`` may be a python number or a numpy scalar (cast to ``ct``)."""
    a = np.zeros(1, dtype="<u8")
    a.view(DTYPES[ct])[0] = x
    v = Value()
    v.ct = ct
    v.bits = int(a[0])
    return v


def build(force: bool = False) -> str:
    """Compile the restatement (g++, seconds). Building the checker is not using it."""
    src = [os.path.join(HERE, f) for f in ("oracle_capi.cpp", "erased_cells_oracle.hpp", "Makefile")]
    if force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in src):
        subprocess.run(["make", "-C", HERE, "-s"], check=True)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        VP, SZ, I = C.c_void_p, C.c_size_t, C.c_int
        PV = C.POINTER(Value)
        sigs = {
            "eco_abi_version": (I, []),
            "eco_last_op_seconds": (C.c_double, []),
            "eco_ctype_union": (I, [I, I]),
            "eco_ctype_can_fit_into": (I, [I, I]),
            "eco_ctype_size_of": (I, [I]),
            "eco_ctype_is_integral": (I, [I]),
            "eco_ctype_is_signed": (I, [I]),
            "eco_ctype_name": (C.c_char_p, [I]),
            "eco_ctype_from_str": (I, [C.c_char_p]),
            "eco_ctype_min_value": (None, [I, PV]),
            "eco_ctype_max_value": (None, [I, PV]),
            "eco_ctype_zero": (None, [I, PV]),
            "eco_ctype_one": (None, [I, PV]),
            "eco_value_convert": (I, [PV, I, PV]),
            "eco_value_binary": (None, [I, PV, PV, PV]),
            "eco_value_neg": (None, [PV, PV]),
            "eco_value_to_prim": (I, [PV, I, PV]),
            "eco_value_cmp": (I, [PV, PV]),
            "eco_value_to_f64": (I, [PV, C.POINTER(C.c_double)]),
            "eco_value_to_i64": (I, [PV, C.POINTER(C.c_int64)]),
            "eco_value_to_u64": (I, [PV, C.POINTER(C.c_uint64)]),
            "eco_buf_binary": (None, [I, I, VP, SZ, I, VP, SZ, C.POINTER(I), C.POINTER(SZ), VP]),
            "eco_buf_scalar": (None, [I, I, VP, SZ, PV, C.POINTER(I), C.POINTER(SZ), VP]),
            "eco_buf_neg": (None, [I, VP, SZ, C.POINTER(I), C.POINTER(SZ), VP]),
            "eco_buf_convert": (I, [I, VP, SZ, I, C.POINTER(I), C.POINTER(SZ), VP]),
            "eco_checked_cast": (I, [I, VP, SZ, I, VP]),
            "eco_buf_min_max": (None, [I, VP, SZ, VP, PV, PV]),
            "eco_buf_cmp": (I, [I, VP, SZ, I, VP, SZ]),
            "eco_buf_fill": (None, [SZ, PV, VP]),
            "eco_buf_put": (I, [I, VP, SZ, SZ, PV]),
            "eco_mask_and": (SZ, [VP, SZ, VP, SZ, VP]),
            "eco_mask_or": (SZ, [VP, SZ, VP, SZ, VP]),
            "eco_mask_not": (None, [VP, SZ, VP]),
            "eco_mask_counts": (None, [VP, SZ, C.POINTER(SZ), C.POINTER(SZ)]),
            "eco_mask_all": (I, [VP, SZ, I]),
            "eco_nodata_value": (I, [I, I, PV, PV]),
            "eco_nodata_is": (I, [I, I, PV, PV]),
            "eco_mask_from_nodata": (None, [I, VP, SZ, I, PV, VP]),
            "eco_fill_nodata": (I, [I, VP, SZ, VP, SZ, I, I, PV, C.POINTER(SZ), VP]),
            "eco_tight_binary": (None, [I, I, VP, I, VP, SZ, VP]),
            "eco_tight_convert": (I, [I, VP, SZ, I, VP]),
            "eco_tight_min_max": (None, [I, VP, SZ, VP, PV, PV]),
        }
        for name, (res, args) in sigs.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _c(a) -> np.ndarray:
    a = np.ascontiguousarray(a)
    assert a.ndim == 1
    return a


class NarrowingError(Exception):
    """src/error.rs:14-15"""

    def __init__(self, src, dst):
        super().__init__(f"Invalid narrowing from cell-type {NAMES[src]} to {NAMES[dst]}")
        self.src, self.dst = src, dst


# ---------------------------------------------------------------------------------------------
# CellType
# ---------------------------------------------------------------------------------------------
def union(a, b): return lib().eco_ctype_union(a, b)
def can_fit_into(a, b): return bool(lib().eco_ctype_can_fit_into(a, b))
def size_of(a): return lib().eco_ctype_size_of(a)
def is_integral(a): return bool(lib().eco_ctype_is_integral(a))
def is_signed(a): return bool(lib().eco_ctype_is_signed(a))
def name(a): return lib().eco_ctype_name(a).decode()
def from_str(s): return lib().eco_ctype_from_str(s.encode())


def _v(fn, *a):
    o = Value()
    fn(*a, C.byref(o))
    return o


def min_value(ct): return _v(lib().eco_ctype_min_value, ct)
def max_value(ct): return _v(lib().eco_ctype_max_value, ct)
def zero(ct): return _v(lib().eco_ctype_zero, ct)
def one(ct): return _v(lib().eco_ctype_one, ct)


# ---------------------------------------------------------------------------------------------
# CellValue
# ---------------------------------------------------------------------------------------------
def value_convert(v: Value, ct: int) -> Value:
    o = Value()
    if lib().eco_value_convert(C.byref(v), ct, C.byref(o)) != OK:
        raise NarrowingError(v.ct, ct)
    return o


def value_binary(op, l: Value, r: Value) -> Value:
    o = Value()
    lib().eco_value_binary(op, C.byref(l), C.byref(r), C.byref(o))
    return o


def value_neg(v: Value) -> Value:
    return _v(lib().eco_value_neg, C.byref(v))


def value_to_prim(v: Value, ct: int):
    o = Value()
    return o if lib().eco_value_to_prim(C.byref(v), ct, C.byref(o)) else None


def value_cmp(l: Value, r: Value) -> int:
    return lib().eco_value_cmp(C.byref(l), C.byref(r))


def value_to_f64(v: Value):
    o = C.c_double()
    return None if lib().eco_value_to_f64(C.byref(v), C.byref(o)) else o.value


def value_to_i64(v: Value):
    o = C.c_int64()
    return None if lib().eco_value_to_i64(C.byref(v), C.byref(o)) else o.value


def value_to_u64(v: Value):
    o = C.c_uint64()
    return None if lib().eco_value_to_u64(C.byref(v), C.byref(o)) else o.value


# ---------------------------------------------------------------------------------------------
# CellBuffer ops (faithful: per-cell tagged dispatch, as the reference executes them)
# ---------------------------------------------------------------------------------------------
def _result(ct: C.c_int, n: C.c_size_t, raw: np.ndarray) -> np.ndarray:
    dt = DTYPES[ct.value]
    return raw[: n.value * dt.itemsize].view(dt).copy()


def binary(op, l, r) -> np.ndarray:
    l, r = _c(l), _c(r)
    n = min(len(l), len(r))
    raw = np.empty(max(n, 1) * 8, dtype=np.uint8)
    ct, ln = C.c_int(), C.c_size_t()
    lib().eco_buf_binary(op, ct_of(l), _p(l), len(l), ct_of(r), _p(r), len(r), C.byref(ct), C.byref(ln), _p(raw))
    return _result(ct, ln, raw)


def scalar(op, l, r: Value) -> np.ndarray:
    l = _c(l)
    raw = np.empty(max(len(l), 1) * 8, dtype=np.uint8)
    ct, ln = C.c_int(), C.c_size_t()
    lib().eco_buf_scalar(op, ct_of(l), _p(l), len(l), C.byref(r), C.byref(ct), C.byref(ln), _p(raw))
    return _result(ct, ln, raw)


def neg(a) -> np.ndarray:
    a = _c(a)
    raw = np.empty(max(len(a), 1) * 8, dtype=np.uint8)
    ct, ln = C.c_int(), C.c_size_t()
    lib().eco_buf_neg(ct_of(a), _p(a), len(a), C.byref(ct), C.byref(ln), _p(raw))
    return _result(ct, ln, raw)


def convert(a, dst: int) -> np.ndarray:
    a = _c(a)
    raw = np.empty(max(len(a), 1) * 8, dtype=np.uint8)
    ct, ln = C.c_int(), C.c_size_t()
    if lib().eco_buf_convert(ct_of(a), _p(a), len(a), dst, C.byref(ct), C.byref(ln), _p(raw)) != OK:
        raise NarrowingError(ct_of(a), dst)
    return _result(ct, ln, raw)


def checked_cast(a, dst: int) -> np.ndarray:
    """Extend<C> semantics (src/buffer.rs:205-221): value-checked `to_<p>()`; a failing cell is the reference's panic."""
    a = _c(a)
    o = np.empty(len(a), dtype=DTYPES[dst])
    if lib().eco_checked_cast(ct_of(a), _p(a), len(a), dst, _p(o)) != OK:
        raise NarrowingError(ct_of(a), dst)
    return o


def min_max(a, mask=None):
    a = _c(a)
    mn, mx = Value(), Value()
    m = None
    if mask is not None:
        m = np.ascontiguousarray(mask, dtype=np.uint8)
        assert len(m) == len(a)
    lib().eco_buf_min_max(ct_of(a), _p(a), len(a), _p(m) if m is not None else None, C.byref(mn), C.byref(mx))
    return mn, mx


def buffer_cmp(l, r) -> int:
    l, r = _c(l), _c(r)
    return lib().eco_buf_cmp(ct_of(l), _p(l), len(l), ct_of(r), _p(r), len(r))


def fill(n: int, v: Value) -> np.ndarray:
    o = np.empty(n, dtype=DTYPES[v.ct])
    lib().eco_buf_fill(n, C.byref(v), _p(o))
    return o


def put(a: np.ndarray, idx: int, v: Value) -> int:
    """In place; returns the status (OK / NARROWING / OOB)."""
    return lib().eco_buf_put(ct_of(a), _p(a), len(a), idx, C.byref(v))


def last_op_seconds() -> float:
    return lib().eco_last_op_seconds()


# ---------------------------------------------------------------------------------------------
# Mask / NoData (masks are bool arrays, one byte per cell, as in the reference)
# ---------------------------------------------------------------------------------------------
def _m(m) -> np.ndarray:
    return np.ascontiguousarray(m, dtype=np.uint8)


def mask_and(l, r):
    l, r = _m(l), _m(r)
    o = np.empty(min(len(l), len(r)), dtype=np.uint8)
    lib().eco_mask_and(_p(l), len(l), _p(r), len(r), _p(o))
    return o.astype(bool)


def mask_or(l, r):
    l, r = _m(l), _m(r)
    o = np.empty(min(len(l), len(r)), dtype=np.uint8)
    lib().eco_mask_or(_p(l), len(l), _p(r), len(r), _p(o))
    return o.astype(bool)


def mask_not(m):
    m = _m(m)
    o = np.empty(len(m), dtype=np.uint8)
    lib().eco_mask_not(_p(m), len(m), _p(o))
    return o.astype(bool)


def mask_counts(m):
    m = _m(m)
    d, nd = C.c_size_t(), C.c_size_t()
    lib().eco_mask_counts(_p(m), len(m), C.byref(d), C.byref(nd))
    return d.value, nd.value


def mask_all(m, v: bool) -> bool:
    m = _m(m)
    return bool(lib().eco_mask_all(_p(m), len(m), int(v)))


def nodata_value(kind: int, ct: int, v: Value | None = None):
    o = Value()
    has = lib().eco_nodata_value(kind, ct, C.byref(v) if v is not None else None, C.byref(o))
    return o if has else None


def nodata_is(kind: int, ct: int, nd: Value | None, v: Value) -> bool:
    return bool(lib().eco_nodata_is(kind, ct, C.byref(nd) if nd is not None else None, C.byref(v)))


def mask_from_nodata(a, kind: int, nd: Value | None = None):
    a = _c(a)
    o = np.empty(len(a), dtype=np.uint8)
    lib().eco_mask_from_nodata(ct_of(a), _p(a), len(a), kind, C.byref(nd) if nd is not None else None, _p(o))
    return o.astype(bool)


def fill_nodata(a, mask, dst: int, kind: int, nd: Value | None = None) -> np.ndarray:
    a, m = _c(a), _m(mask)
    raw = np.empty(max(len(a), 1) * 8, dtype=np.uint8)
    ln = C.c_size_t()
    st = lib().eco_fill_nodata(ct_of(a), _p(a), len(a), _p(m), len(m), dst, kind,
                               C.byref(nd) if nd is not None else None, C.byref(ln), _p(raw))
    if st != OK:
        raise NarrowingError(ct_of(a), dst)
    return raw[: ln.value * DTYPES[dst].itemsize].view(DTYPES[dst]).copy()


# ---------------------------------------------------------------------------------------------
# tight flavour (typed loops; same arithmetic). For big sweeps.
# ---------------------------------------------------------------------------------------------
def tight_binary(op, l, r) -> np.ndarray:
    l, r = _c(l), _c(r)
    n = min(len(l), len(r))
    if n == 0:
        return np.empty(0, dtype=np.uint8)
    o = np.empty(n, dtype=np.float64)
    lib().eco_tight_binary(op, ct_of(l), _p(l), ct_of(r), _p(r), n, _p(o))
    return o


def tight_scalar(op, l, r: Value) -> np.ndarray:
    l = _c(l)
    if len(l) == 0:
        return np.empty(0, dtype=np.uint8)
    rr = np.full(len(l), r.numpy(), dtype=DTYPES[r.ct])
    return tight_binary(op, l, rr)


def tight_convert(a, dst: int) -> np.ndarray:
    a = _c(a)
    if ct_of(a) != dst and len(a) == 0:
        if not can_fit_into(ct_of(a), dst):
            raise NarrowingError(ct_of(a), dst)
        return np.empty(0, dtype=np.uint8)
    o = np.empty(len(a), dtype=DTYPES[dst])
    if lib().eco_tight_convert(ct_of(a), _p(a), len(a), dst, _p(o)) != OK:
        raise NarrowingError(ct_of(a), dst)
    return o


def tight_min_max(a, mask=None):
    a = _c(a)
    mn, mx = Value(), Value()
    m = _m(mask) if mask is not None else None
    lib().eco_tight_min_max(ct_of(a), _p(a), len(a), _p(m) if m is not None else None, C.byref(mn), C.byref(mx))
    return mn, mx
