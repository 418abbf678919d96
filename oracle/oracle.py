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


# ---------------------------------------------------------------------------------------------
# EXTENSION — statistics (count / min / max / mean / population stddev).
# The reference has NO statistics beyond min_max and Mask::counts (STATISTICS_MEAN /
# STATISTICS_STDDEV occur only in a comment quoting gdal_calc.py output, src/gdal/rasterband.rs:152-156), so
# there is nothing to pin against: PARITY UNPINNED. What is restated here is this repo's own DEFINITION
# (DESIGN.md §4.6), chosen so that the result is a function of the multiset of valid cells only — independent
# of summation order, grid size and the number of GPUs:
#   1. (mn, mx) = min_max of the valid cells; x = to_f64(cell)
#   2. pivot p = fl(xmin/2 + xmax/2); E = binary exponent with max|x - p| < 2^E (clamped to [-1000, 1024])
#   3a. integer cells of at most 32 bits: y = (x - p) * 2^-E is exact; S1 = RN(sum y), S2 = RN(sum y^2), the sums
#       taken exactly (they follow from count, A = sum x, B = sum x^2 and s = 2p = min + max)
#   3b. 64-bit integers and floats: y = fl(fl(x - p) * 2^-E), z = fl(y * y); each is split into two 48-bit
#       fixed-point windows (units 2^-47 and 2^-95, round-to-nearest-even at the second), the four integer
#       streams are summed EXACTLY; S1 = fl(fl(X1) * 2^-47 + fl(X2) * 2^-95), S2 likewise from the z windows
#   4. m1 = S1/n, m2 = S2/n, var = max(m2 - m1*m1, 0); mean = p + m1*2^E; stddev = sqrt(var)*2^E
# numpy float64 arithmetic below is IEEE binary64 with one rounding per operation, like the device code
# (__dadd_rn/__dmul_rn) and the host finish.
_C1 = 48.0                      # 1.5 * 2^5: ulp = 2^-47
_C2 = 48.0 * 2.0 ** -48         # ulp = 2^-95
_K1 = int(np.float64(_C1).view(np.int64))
_K2 = int(np.float64(_C2).view(np.int64))
_QNAN = float(np.uint64(0x7FF8000000000000).view(np.float64))
ST_REGULAR, ST_EMPTY, ST_NONFINITE = range(3)


def _value_f64(v: Value) -> float:
    return float(v.numpy().astype(np.float64)) if v.ct != Float64 else float(v.numpy())


def statistics_plan(mn: Value, mx: Value):
    """(kind, pivot, exp2) from the min/max of the valid cells."""
    if value_cmp(mn, mx) > 0:  # the (T::MAX, T::MIN) seeds survived: no valid cell
        return ST_EMPTY, 0.0, 0
    lo, hi = np.float64(_value_f64(mn)), np.float64(_value_f64(mx))
    if not (np.isfinite(lo) and np.isfinite(hi)):
        return ST_NONFINITE, 0.0, 0
    if mn.ct == Float32:  # quantised route: exp2 = E with every |cell| < 2^E, no pivot on the way in
        top = max(abs(float(lo)), abs(float(hi)))
        return ST_REGULAR, 0.0, (0 if top == 0 else int(np.frexp(top)[1]))
    p = np.float64(lo * np.float64(0.5)) + np.float64(hi * np.float64(0.5))
    d = max(np.float64(hi - p), np.float64(p - lo))
    e = 0 if d == 0 else int(np.frexp(d)[1])
    return ST_REGULAR, float(p), max(-1000, min(1024, e))


def _windows(v: np.ndarray):
    t1 = v + np.float64(_C1)
    x1 = t1.view(np.int64) - np.int64(_K1)
    r = v - (t1 - np.float64(_C1))
    t2 = r + np.float64(_C2)
    x2 = t2.view(np.int64) - np.int64(_K2)
    return x1, x2


def _isum(x: np.ndarray) -> int:
    """exact integer sum of int64 terms below 2^48 in magnitude"""
    return sum(int(x[i:i + (1 << 14)].sum(dtype=np.int64)) for i in range(0, len(x), 1 << 14))


def _isum_sq(q: np.ndarray) -> int:
    """exact sum of squares of int64 terms below 2^27 in magnitude (squares below 2^54: 512 of them fit int64)"""
    sq = q * q
    return sum(int(sq[i:i + 512].sum(dtype=np.int64)) for i in range(0, len(sq), 512))


def integer_route(ct: int) -> bool:
    return is_integral(ct) and size_of(ct) <= 4


def moments_raw(a, mask, pivot: float, exp2: int):
    """[count, sum X1, sum X2, sum Z1, sum Z2] — or, on the integer route, [count, sum x, sum x^2, 0, 0] — as python ints."""
    a = _c(a)
    if mask is not None:
        a = a[np.ascontiguousarray(mask, dtype=bool)]
    if integer_route(ct_of(a)):
        xs = [int(v) for v in a]
        return [len(xs), sum(xs), sum(v * v for v in xs), 0, 0]
    if ct_of(a) == Float32:
        # Float32 cells are read as the int32 raster q = rint(x * 2^(26 - E)) (half to even; the scaling is exact in f64)
        # and summed exactly like integer cells
        with np.errstate(all="ignore"):
            q = np.rint(np.ldexp(a.astype(np.float64), 26 - exp2)).astype(np.int64)
        return [len(q), _isum(q), sum(int(v) * int(v) for v in q) if len(q) < (1 << 16) else _isum_sq(q), 0, 0]
    with np.errstate(all="ignore"):
        y = (a.astype(np.float64) - np.float64(pivot)) * np.float64(np.ldexp(1.0, -exp2))
        x1, x2 = _windows(y)
        z1, z2 = _windows(y * y)
    return [len(a), _isum(x1), _isum(x2), _isum(z1), _isum(z2)]


def statistics_finish(raws, mn: Value, mx: Value):
    """combine per-shard raw accumulators -> dict(count, min, max, mean, stddev)"""
    import math
    kind, p, e = statistics_plan(mn, mx)
    tot = [sum(r[i] for r in raws) for i in range(5)]
    out = dict(count=tot[0], min=mn, max=mx, mean=_QNAN, stddev=_QNAN)
    if kind == ST_EMPTY or tot[0] == 0:
        return out
    if kind == ST_NONFINITE:
        lo, hi = _value_f64(mn), _value_f64(mx)
        if not (math.isnan(lo) or math.isnan(hi)) and not (lo == -math.inf and hi == math.inf):
            out["mean"] = lo if lo == -math.inf else hi
        return out
    if mn.ct == Float32:
        # finish the int32 raster of quantised cells (min / max quantise monotonically), scale back by 2^-(26 - E): exact
        k = 26 - e
        qmn = value(Int32, int(np.rint(math.ldexp(_value_f64(mn), k))))
        qmx = value(Int32, int(np.rint(math.ldexp(_value_f64(mx), k))))
        q = statistics_finish(raws, qmn, qmx)
        out["mean"], out["stddev"] = math.ldexp(q["mean"], -k), math.ldexp(q["stddev"], -k)
        return out
    n = float(tot[0])
    if integer_route(mn.ct):
        cnt, A, B, s2p = tot[0], tot[1], tot[2], int(p * 2.0)  # y = (2x - s) / 2 exactly
        s1 = math.ldexp(float(2 * A - cnt * s2p), -1 - e)   # python int -> float is round-to-nearest-even
        s2 = math.ldexp(float(4 * B - 4 * s2p * A + cnt * s2p * s2p), -2 - 2 * e)
    else:
        s1 = math.ldexp(float(tot[1]), -47) + math.ldexp(float(tot[2]), -95)
        s2 = math.ldexp(float(tot[3]), -47) + math.ldexp(float(tot[4]), -95)
    m1, m2 = s1 / n, s2 / n
    var = m2 - m1 * m1
    if var < 0:
        var = 0.0
    out["mean"] = p + math.ldexp(m1, e)
    out["stddev"] = math.ldexp(math.sqrt(var), e)
    return out


def statistics(a, mask=None):
    mn, mx = min_max(a, mask)
    kind, p, e = statistics_plan(mn, mx)
    if kind == ST_REGULAR:
        raw = moments_raw(a, mask, p, e)
    else:
        raw = [int(len(a) if mask is None else np.count_nonzero(mask)), 0, 0, 0, 0]
    return statistics_finish([raw], mn, mx)
