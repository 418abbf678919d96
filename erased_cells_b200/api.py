"""Python mirror of the erased-cells public API over the B200 C ABI.

Same names, argument meaning and error behaviour as the reference crate (s22s/erased-cells v0.1.1):
``CellType`` (src/ctype.rs), ``CellValue`` (src/value.rs), ``CellBuffer`` + ``BufferOps``
(src/buffer.rs, src/lib.rs:104-163), ``Mask`` (src/masked/mask.rs), ``NoData``
(src/masked/nodata.rs) and ``MaskedCellBuffer`` (src/masked/masked_buffer.rs). Every buffer and mask
lives in GPU memory behind an ``ec_buf`` / ``ec_mask`` handle; all per-cell work runs in the CUDA
library. Rust idioms map as: ``Result::Err(NarrowingError)`` -> ``NarrowingError`` exception, index
panics -> ``IndexError``, ``assert_eq!`` panics -> ``AssertionError``, ``Drop`` -> ``__del__``.
"""
from __future__ import annotations

import ctypes as C
import enum

import numpy as np

from . import _lib
from ._lib import NarrowingError, Value, check, lib

_DTYPES = [np.dtype(d) for d in ("u1", "u2", "u4", "u8", "i1", "i2", "i4", "i8", "f4", "f8")]


# ---------------------------------------------------------------------------------------------
# CellType — src/ctype.rs
# ---------------------------------------------------------------------------------------------
class CellType(enum.IntEnum):
    UInt8 = 0
    UInt16 = 1
    UInt32 = 2
    UInt64 = 3
    Int8 = 4
    Int16 = 5
    Int32 = 6
    Int64 = 7
    Float32 = 8
    Float64 = 9

    @staticmethod
    def iter():
        return iter(CellType)

    def is_integral(self) -> bool:
        return bool(lib().ec_ctype_is_integral(int(self)))

    def is_signed(self) -> bool:
        return bool(lib().ec_ctype_is_signed(int(self)))

    def size_of(self) -> int:
        return lib().ec_ctype_size_of(int(self))

    def union(self, other: "CellType") -> "CellType":
        return CellType(lib().ec_ctype_union(int(self), int(other)))

    def can_fit_into(self, other: "CellType") -> bool:
        return bool(lib().ec_ctype_can_fit_into(int(self), int(other)))

    def _v(self, fn) -> "CellValue":
        v = Value()
        check(fn(int(self), C.byref(v)))
        return CellValue._wrap(v)

    def zero(self): return self._v(lib().ec_ctype_zero)
    def one(self): return self._v(lib().ec_ctype_one)
    def min_value(self): return self._v(lib().ec_ctype_min_value)
    def max_value(self): return self._v(lib().ec_ctype_max_value)

    def __str__(self):
        return lib().ec_ctype_name(int(self)).decode()

    @staticmethod
    def from_str(s: str) -> "CellType":
        out = C.c_uint8()
        check(lib().ec_ctype_from_name(s.encode(), C.byref(out)))
        return CellType(out.value)

    @property
    def dtype(self) -> np.dtype:
        return _DTYPES[int(self)]

    @staticmethod
    def of(x) -> "CellType":
        """CellEncoding::cell_type (src/encoding.rs:9-40) for numpy dtypes / arrays / scalars."""
        dt = np.dtype(x.dtype if hasattr(x, "dtype") else x)
        try:
            return CellType(_DTYPES.index(dt))
        except ValueError:
            raise TypeError(f"{dt} has no CellType (not CellEncoding)") from None


# ---------------------------------------------------------------------------------------------
# CellValue — src/value.rs
# ---------------------------------------------------------------------------------------------
def _into_value(x) -> "CellValue":
    """`impl<T: CellEncoding> From<T> for CellValue` with Rust literal defaults: int -> i32, float -> f64."""
    if isinstance(x, CellValue):
        return x
    if isinstance(x, (np.generic,)):
        return CellValue(CellType.of(x), x)
    if isinstance(x, bool):
        raise TypeError("bool is not CellEncoding")
    if isinstance(x, int):
        return CellValue(CellType.Int32, x)
    if isinstance(x, float):
        return CellValue(CellType.Float64, x)
    raise TypeError(f"cannot convert {type(x).__name__} into CellValue")


class CellValue:
    __slots__ = ("_v",)

    def __init__(self, cell_type: CellType, value):
        a = np.zeros(1, dtype="<u8")
        a.view(_DTYPES[int(cell_type)])[0] = value
        self._v = Value()
        self._v.ct = int(cell_type)
        self._v.bits = int(a[0])

    @staticmethod
    def new(value) -> "CellValue":
        return _into_value(value)

    @staticmethod
    def _wrap(v: Value) -> "CellValue":
        o = CellValue.__new__(CellValue)
        o._v = Value()
        o._v.ct, o._v.bits = v.ct, v.bits
        return o

    def cell_type(self) -> CellType:
        return CellType(self._v.ct)

    @property
    def bits(self) -> int:
        return int(self._v.bits)

    def value(self):
        """Payload as a numpy scalar of the tagged type."""
        return np.array([self._v.bits], dtype="<u8").view(_DTYPES[self._v.ct])[0]

    def convert(self, cell_type: CellType) -> "CellValue":
        out = Value()
        check(lib().ec_value_convert(C.byref(self._v), int(cell_type), C.byref(out)))
        return CellValue._wrap(out)

    def get(self, cell_type: CellType):
        """`get::<T>()`: the payload as T (NarrowingError if T is narrower)."""
        return self.convert(cell_type).value()

    def unify(self, other: "CellValue"):
        dest = self.cell_type().union(other.cell_type())
        return self.convert(dest), other.convert(dest)

    def to_f64(self):
        o, some = C.c_double(), C.c_int()
        check(lib().ec_value_to_f64(C.byref(self._v), C.byref(o), C.byref(some)))
        return o.value if some.value else None

    def to_i64(self):
        o, some = C.c_int64(), C.c_int()
        check(lib().ec_value_to_i64(C.byref(self._v), C.byref(o), C.byref(some)))
        return o.value if some.value else None

    def to_u64(self):
        o, some = C.c_uint64(), C.c_int()
        check(lib().ec_value_to_u64(C.byref(self._v), C.byref(o), C.byref(some)))
        return o.value if some.value else None

    def to_prim(self, cell_type: CellType):
        """`self.to_<p>()`: value-checked ToPrimitive chain; None where the reference yields None."""
        out, some = Value(), C.c_int()
        check(lib().ec_value_to_prim(C.byref(self._v), int(cell_type), C.byref(out), C.byref(some)))
        return CellValue._wrap(out) if some.value else None

    def _bin(self, op, rhs):
        r = _into_value(rhs)
        out = Value()
        check(lib().ec_value_binary(op, C.byref(self._v), C.byref(r._v), C.byref(out)))
        return CellValue._wrap(out)

    def __add__(self, r): return self._bin(0, r)
    def __sub__(self, r): return self._bin(1, r)
    def __mul__(self, r): return self._bin(2, r)
    def __truediv__(self, r): return self._bin(3, r)

    def __neg__(self):
        out = Value()
        check(lib().ec_value_neg(C.byref(self._v), C.byref(out)))
        return CellValue._wrap(out)

    def cmp(self, other) -> int:
        o = C.c_int()
        check(lib().ec_value_cmp(C.byref(self._v), C.byref(_into_value(other)._v), C.byref(o)))
        return o.value

    def __eq__(self, other):
        try:
            return self.cmp(other) == 0
        except TypeError:
            return NotImplemented

    def __lt__(self, o): return self.cmp(o) < 0
    def __le__(self, o): return self.cmp(o) <= 0
    def __gt__(self, o): return self.cmp(o) > 0
    def __ge__(self, o): return self.cmp(o) >= 0
    __hash__ = None

    def is_nodata(self, no_data: "NoData") -> bool:
        """IsNodata::is (src/masked/nodata.rs:59-62)."""
        return no_data.is_(self)

    def __repr__(self):
        return f"{self.cell_type()}({self.value()!r})"


# ---------------------------------------------------------------------------------------------
# CellBuffer — src/buffer.rs (BufferOps: src/lib.rs:104-163)
# ---------------------------------------------------------------------------------------------
def _host(a, ct=None) -> np.ndarray:
    a = np.ascontiguousarray(a if ct is None else np.asarray(a, dtype=_DTYPES[int(ct)]))
    if a.ndim != 1:
        a = a.reshape(-1)
    return a


class Statistics:
    """count / min / max / mean / population stddev of the valid cells (extension; ec_statistics)."""
    __slots__ = ("count", "min", "max", "mean", "stddev")

    def __init__(self, raw: "_lib.Statistics"):
        self.count = int(raw.count)
        self.min, self.max = CellValue._wrap(raw.min), CellValue._wrap(raw.max)
        self.mean, self.stddev = float(raw.mean), float(raw.stddev)

    def __repr__(self):
        return f"Statistics(count={self.count}, min={self.min}, max={self.max}, mean={self.mean}, stddev={self.stddev})"


def _statistics(buf_h, mask_h) -> Statistics:
    raw = _lib.Statistics()
    check(lib().ec_buf_statistics(buf_h, mask_h, C.byref(raw)))
    return Statistics(raw)


class CellBuffer:
    __slots__ = ("_h", "_keep")

    def __init__(self, handle):
        self._h = handle
        self._keep = None  # host array of an upload still in flight (from_vec(..., wait=False))

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib._lib is not None:
            _lib._lib.ec_buf_free(h)

    @staticmethod
    def _take(h: C.c_void_p) -> "CellBuffer":
        return CellBuffer(h.value)

    # -- constructors ----------------------------------------------------------------------
    @staticmethod
    def new(data) -> "CellBuffer":
        return CellBuffer.from_vec(data)

    @staticmethod
    def from_vec(data, wait: bool = True) -> "CellBuffer":
        """from_vec / From<Vec<T>> (src/buffer.rs:64-66, :252-263): one H2D copy.

        wait=False returns at once: the copy runs on the upload stream (overlapping D2H traffic) and the
        buffer keeps `data` alive until it lands — the analogue of from_vec taking ownership of the Vec."""
        a = _host(data)
        ct = CellType.of(a)
        h = C.c_void_p()
        if wait:
            check(lib().ec_buf_from_host(int(ct), a.ctypes.data_as(C.c_void_p), a.size, C.byref(h)))
            check(lib().ec_synchronize())  # `a` may be a temporary
            return CellBuffer._take(h)
        check(lib().ec_buf_from_host_async(int(ct), a.ctypes.data_as(C.c_void_p), a.size, C.byref(h)))
        b = CellBuffer._take(h)
        b._keep = a
        return b

    def wait(self) -> "CellBuffer":
        check(lib().ec_buf_wait(self._h))
        self._keep = None
        return self

    @staticmethod
    def with_defaults(len_: int, ct: CellType) -> "CellBuffer":
        h = C.c_void_p()
        check(lib().ec_buf_with_defaults(len_, int(ct), C.byref(h)))
        return CellBuffer._take(h)

    @staticmethod
    def fill(len_: int, value) -> "CellBuffer":
        h = C.c_void_p()
        check(lib().ec_buf_fill(len_, C.byref(_into_value(value)._v), C.byref(h)))
        return CellBuffer._take(h)

    @staticmethod
    def fill_via(len_: int, f, ct: CellType) -> "CellBuffer":
        """fill_via::<T, F>: the closure runs on the host (it is host code in the reference too), one H2D."""
        return CellBuffer.from_vec(np.array([f(i) for i in range(len_)], dtype=_DTYPES[int(ct)]))

    @staticmethod
    def from_iter(values) -> "CellBuffer":
        """FromIterator<CellValue> (src/buffer.rs:229-250): type of the first element; empty => UInt8."""
        values = list(values)
        if not values:
            return CellBuffer.with_defaults(0, CellType.UInt8)
        if isinstance(values[0], CellValue):
            ct = values[0].cell_type()
            return CellBuffer.from_vec(np.array([v.get(ct) for v in values], dtype=_DTYPES[int(ct)]))
        return CellBuffer.from_vec(np.asarray(values))

    @staticmethod
    def wrap_device(ct: CellType, device_ptr: int, len_: int) -> "CellBuffer":
        """A row strip of a raster that is already in HBM (e.g. a torch tensor's storage); not owned."""
        h = C.c_void_p()
        check(lib().ec_buf_wrap_device(int(ct), C.c_void_p(device_ptr), len_, C.byref(h)))
        return CellBuffer._take(h)

    # -- BufferOps ---------------------------------------------------------------------------
    def len(self) -> int:
        return lib().ec_buf_len(self._h)

    __len__ = len

    def is_empty(self) -> bool:
        return self.len() == 0

    def cell_type(self) -> CellType:
        return CellType(lib().ec_buf_ctype(self._h))

    def device_ptr(self) -> int:
        return lib().ec_buf_device_ptr(self._h) or 0

    def shard_count(self) -> int:
        """0 for a buffer on one GPU; otherwise the number of row strips it is kept as (erased_cells_b200.init_devices)."""
        return lib().ec_buf_shard_count(self._h)

    def shards(self) -> list:
        """[(logical_device, cuda_device, offset, len, device_ptr)] of the strips (one entry for a plain buffer)"""
        out = []
        for g in range(max(1, self.shard_count())):
            info = _lib.ShardInfo()
            check(lib().ec_buf_shard(self._h, g, C.byref(info), None))
            out.append((info.logical_device, info.cuda_device, info.offset, info.len, info.device_ptr or 0))
        return out

    def get(self, index: int) -> CellValue:
        v = Value()
        check(lib().ec_buf_get(self._h, index, C.byref(v)))
        return CellValue._wrap(v)

    def put(self, index: int, value) -> None:
        check(lib().ec_buf_put(self._h, index, C.byref(_into_value(value)._v)))

    def extend(self, values) -> None:
        """Extend<C> (src/buffer.rs:205-221)."""
        a = _host(values)
        check(lib().ec_buf_extend_host(self._h, int(CellType.of(a)), a.ctypes.data_as(C.c_void_p), a.size))

    def convert(self, cell_type: CellType) -> "CellBuffer":
        h = C.c_void_p()
        check(lib().ec_buf_convert(self._h, int(cell_type), C.byref(h)))
        return CellBuffer._take(h)

    def min_max(self):
        mn, mx = Value(), Value()
        check(lib().ec_buf_min_max(self._h, None, C.byref(mn), C.byref(mx)))
        return CellValue._wrap(mn), CellValue._wrap(mx)

    def statistics(self) -> "Statistics":
        """count / min / max / mean / population stddev — an extension (the reference stops at min_max)."""
        return _statistics(self._h, None)

    def to_vec(self, cell_type: CellType | None = None, out: np.ndarray | None = None) -> np.ndarray:
        """to_vec::<T>() (src/buffer.rs:175-185): convert on the device, then one D2H copy."""
        src = self if cell_type is None or cell_type == self.cell_type() else self.convert(cell_type)
        ct = src.cell_type() if cell_type is None else cell_type
        n = src.len()
        if out is None:
            out = np.empty(n, dtype=_DTYPES[int(ct)])
        assert out.dtype == _DTYPES[int(ct)] and out.size >= n and out.flags.c_contiguous
        check(lib().ec_buf_to_host(src._h, out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out[:n]

    def clone(self) -> "CellBuffer":
        h = C.c_void_p()
        check(lib().ec_buf_clone(self._h, C.byref(h)))
        return CellBuffer._take(h)

    def view(self, offset: int, len_: int) -> "CellBuffer":
        """Cells [offset, offset + len) as a buffer sharing this allocation (a row strip); offset on a 32-byte boundary."""
        h = C.c_void_p()
        check(lib().ec_buf_view(self._h, offset, len_, C.byref(h)))
        return CellBuffer._take(h)

    def __iter__(self):
        """IntoIterator for &CellBuffer (src/buffer.rs:278-305): yields CellValues from a host copy."""
        ct = self.cell_type()
        for x in self.to_vec():
            yield CellValue(ct, x)

    # -- std::ops (src/buffer.rs:321-371) --------------------------------------------------------
    def _bin(self, op: int, rhs) -> "CellBuffer":
        h = C.c_void_p()
        if isinstance(rhs, CellBuffer):
            check(lib().ec_buf_binary(op, self._h, rhs._h, C.byref(h)))
        else:
            check(lib().ec_buf_scalar(op, self._h, C.byref(_into_value(rhs)._v), C.byref(h)))
        return CellBuffer._take(h)

    def __add__(self, r): return self._bin(0, r)
    def __sub__(self, r): return self._bin(1, r)
    def __mul__(self, r): return self._bin(2, r)
    def __truediv__(self, r): return self._bin(3, r)

    def __neg__(self) -> "CellBuffer":
        h = C.c_void_p()
        check(lib().ec_buf_neg(self._h, C.byref(h)))
        return CellBuffer._take(h)

    # fused chains (same results as the unfused operator chain, one pass over HBM)
    def normalized_difference(self, other: "CellBuffer") -> "CellBuffer":
        """`(&self - &other) / (&self + &other)`"""
        h = C.c_void_p()
        check(lib().ec_buf_normalized_difference(self._h, other._h, C.byref(h)))
        return CellBuffer._take(h)

    def binary_scalar(self, op1: int, other: "CellBuffer", op2: int, scalar) -> "CellBuffer":
        """`(self op1 other) op2 scalar`"""
        h = C.c_void_p()
        check(lib().ec_buf_binary_scalar(op1, self._h, other._h, op2, C.byref(_into_value(scalar)._v), C.byref(h)))
        return CellBuffer._take(h)

    # -- Ord / Eq (src/buffer.rs:373-436) ----------------------------------------------------------
    def cmp(self, other: "CellBuffer") -> int:
        o = C.c_int()
        check(lib().ec_buf_cmp(self._h, other._h, C.byref(o)))
        return o.value

    def __eq__(self, other):
        if not isinstance(other, CellBuffer):
            return NotImplemented
        return self.cmp(other) == 0

    def __lt__(self, o): return self.cmp(o) < 0
    def __le__(self, o): return self.cmp(o) <= 0
    def __gt__(self, o): return self.cmp(o) > 0
    def __ge__(self, o): return self.cmp(o) >= 0
    __hash__ = None

    def __repr__(self):
        """Debug (src/buffer.rs:188-203) with Elided (src/lib.rs:166-194)."""
        n = self.len()  # at most ten cells leave the device: a 4 GB raster in a pytest failure message must not be downloaded
        v = self.to_vec() if n <= 10 else [self.get(i).value() for i in list(range(5)) + list(range(n - 5, n))]
        items = [repr(x.item()) for x in v]
        body = ", ".join(items) if n <= 10 else ", ".join(items[:5]) + ", ... " + ", ".join(items[5:])
        return f"{self.cell_type()}CellBuffer({body})"


# ---------------------------------------------------------------------------------------------
# Mask — src/masked/mask.rs (packed bits in HBM instead of Vec<bool>)
# ---------------------------------------------------------------------------------------------
class Mask:
    __slots__ = ("_h",)

    def __init__(self, handle):
        self._h = handle

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib._lib is not None:
            _lib._lib.ec_mask_free(h)

    @staticmethod
    def _take(h) -> "Mask":
        return Mask(h.value)

    @staticmethod
    def new(values) -> "Mask":
        a = np.ascontiguousarray(np.asarray(values, dtype=bool).reshape(-1)).view(np.uint8)
        h = C.c_void_p()
        check(lib().ec_mask_from_bools(a.ctypes.data_as(C.c_void_p), a.size, C.byref(h)))
        check(lib().ec_synchronize())
        return Mask._take(h)

    @staticmethod
    def fill(len_: int, value: bool) -> "Mask":
        h = C.c_void_p()
        check(lib().ec_mask_fill(len_, int(bool(value)), C.byref(h)))
        return Mask._take(h)

    @staticmethod
    def fill_via(len_: int, f) -> "Mask":
        return Mask.new([bool(f(i)) for i in range(len_)])

    def len(self) -> int:
        return lib().ec_mask_len(self._h)

    __len__ = len

    def is_empty(self) -> bool:
        return self.len() == 0

    def put(self, index: int, value: bool) -> None:
        check(lib().ec_mask_put(self._h, index, int(bool(value))))

    def get(self, index: int) -> bool:
        o = C.c_int()
        check(lib().ec_mask_get(self._h, index, C.byref(o)))
        return bool(o.value)

    __getitem__ = get
    __setitem__ = put

    def extend(self, values) -> None:
        a = np.ascontiguousarray(np.asarray(list(values), dtype=bool)).view(np.uint8)
        check(lib().ec_mask_extend_host(self._h, a.ctypes.data_as(C.c_void_p), a.size))

    def all(self, value: bool) -> bool:
        o = C.c_int()
        check(lib().ec_mask_all(self._h, int(bool(value)), C.byref(o)))
        return bool(o.value)

    def counts(self):
        d, nd = C.c_size_t(), C.c_size_t()
        check(lib().ec_mask_counts(self._h, C.byref(d), C.byref(nd)))
        return d.value, nd.value

    def to_vec(self) -> np.ndarray:
        out = np.empty(self.len(), dtype=np.uint8)
        check(lib().ec_mask_to_bools(self._h, out.ctypes.data_as(C.c_void_p), out.size))
        return out.view(bool)

    def __iter__(self):
        return iter(self.to_vec().tolist())

    def clone(self) -> "Mask":
        h = C.c_void_p()
        check(lib().ec_mask_clone(self._h, C.byref(h)))
        return Mask._take(h)

    def slice(self, offset: int, len_: int) -> "Mask":
        """Validity bits of cells [offset, offset + len) as a mask of its own (a row strip; offset on a 128-cell boundary)."""
        h = C.c_void_p()
        check(lib().ec_mask_slice(self._h, offset, len_, C.byref(h)))
        return Mask._take(h)

    def _op(self, fn, other=None) -> "Mask":
        h = C.c_void_p()
        check(fn(self._h, C.byref(h)) if other is None else fn(self._h, other._h, C.byref(h)))
        return Mask._take(h)

    def __invert__(self): return self._op(lib().ec_mask_not)
    def __and__(self, o): return self._op(lib().ec_mask_and, o)
    def __or__(self, o): return self._op(lib().ec_mask_or, o)

    def cmp(self, other: "Mask") -> int:
        o = C.c_int()
        check(lib().ec_mask_cmp(self._h, other._h, C.byref(o)))
        return o.value

    def __eq__(self, other):
        if not isinstance(other, Mask):
            return NotImplemented
        return self.cmp(other) == 0

    def __lt__(self, o): return self.cmp(o) < 0
    __hash__ = None

    def __repr__(self):
        n = self.len()
        bits = self.to_vec() if n <= 10 else [self.get(i) for i in list(range(5)) + list(range(n - 5, n))]
        v = ["true" if b else "false" for b in bits]
        body = ", ".join(v) if n <= 10 else ", ".join(v[:5]) + ", ... " + ", ".join(v[5:])
        return f"Mask({body})"


# ---------------------------------------------------------------------------------------------
# NoData<T> — src/masked/nodata.rs
# ---------------------------------------------------------------------------------------------
class NoData:
    """``NoData.none(T)``, ``NoData.default(T)``, ``NoData.new(T, v)`` for NoData::<T>::{None, Default, Value(v)}."""
    NONE, DEFAULT, VALUE = 0, 1, 2
    __slots__ = ("kind", "ct", "_value")

    def __init__(self, kind: int, ct: CellType, value=None):
        self.kind, self.ct = kind, CellType(ct)
        self._value = CellValue(self.ct, value) if kind == NoData.VALUE else None

    @staticmethod
    def none(ct): return NoData(NoData.NONE, ct)
    @staticmethod
    def default(ct): return NoData(NoData.DEFAULT, ct)
    @staticmethod
    def new(ct, value): return NoData(NoData.VALUE, ct, value)

    def _ptr(self):
        return C.byref(self._value._v) if self._value is not None else None

    def value(self):
        out, has = Value(), C.c_int()
        check(lib().ec_nodata_value(self.kind, int(self.ct), self._ptr(), C.byref(out), C.byref(has)))
        return CellValue._wrap(out).value() if has.value else None

    def is_(self, v) -> bool:
        """NoData::is (src/masked/nodata.rs:42-49): total-order equality with the sentinel."""
        s = self.value()
        return False if s is None else CellValue(self.ct, s) == _into_value(v)


# ---------------------------------------------------------------------------------------------
# MaskedCellBuffer — src/masked/masked_buffer.rs
# ---------------------------------------------------------------------------------------------
class MaskedCellBuffer:
    __slots__ = ("_buf", "_mask")

    def __init__(self, buffer: CellBuffer, mask: Mask):
        assert buffer.len() == mask.len(), "Mask and buffer must have the same length."
        self._buf, self._mask = buffer, mask

    new = None  # set below (constructor alias)

    @staticmethod
    def from_vec(data) -> "MaskedCellBuffer":
        b = CellBuffer.from_vec(data)
        return MaskedCellBuffer(b, Mask.fill(b.len(), True))

    @staticmethod
    def from_vec_with_nodata(data, nodata: NoData) -> "MaskedCellBuffer":
        """from_vec_with_nodata (src/masked/masked_buffer.rs:62-71)."""
        b = CellBuffer.from_vec(data)
        return MaskedCellBuffer.from_buffer_with_nodata(b, nodata)

    @staticmethod
    def from_iter(items, ct: CellType | None = None) -> "MaskedCellBuffer":
        """FromIterator<C> (src/masked/masked_buffer.rs:257-261: all cells valid) and FromIterator<(C, bool)> (:263-278):
        plain cells, or (cell, valid) pairs. The cell type is `ct`, else numpy scalars keep theirs and Python numbers
        follow Rust's literal defaults (int -> i32, float -> f64); an empty iterator of pairs gives `ct` (default Int32) cells."""
        items = list(items)
        paired = bool(items) and isinstance(items[0], tuple)
        cells = [p[0] for p in items] if paired else items
        if ct is not None:
            dt = _DTYPES[int(ct)]
        elif cells:
            v0 = cells[0]
            dt = v0.dtype if isinstance(v0, np.generic) else (np.int32 if isinstance(v0, int) else np.float64)
        else:
            dt = np.int32
        buf = CellBuffer.from_vec(np.array(cells, dtype=dt)) if cells else CellBuffer.with_defaults(0, CellType.of(np.dtype(dt)))
        return MaskedCellBuffer(buf, Mask.new([bool(p[1]) for p in items]) if paired else Mask.fill(len(cells), True))

    @staticmethod
    def from_buffer_with_nodata(b: CellBuffer, nodata: NoData) -> "MaskedCellBuffer":
        h = C.c_void_p()
        check(lib().ec_mask_from_nodata(b._h, nodata.kind, nodata._ptr(), C.byref(h)))
        return MaskedCellBuffer(b, Mask._take(h))

    @staticmethod
    def with_defaults(len_, ct): return MaskedCellBuffer(CellBuffer.with_defaults(len_, ct), Mask.fill(len_, True))
    @staticmethod
    def fill(len_, value): return MaskedCellBuffer(CellBuffer.fill(len_, value), Mask.fill(len_, True))
    @staticmethod
    def fill_via(len_, f, ct): return MaskedCellBuffer(CellBuffer.fill_via(len_, f, ct), Mask.fill(len_, True))

    @staticmethod
    def fill_with_mask_via(len_: int, mv, ct: CellType) -> "MaskedCellBuffer":
        pairs = [mv(i) for i in range(len_)]
        return MaskedCellBuffer(CellBuffer.from_vec(np.array([p[0] for p in pairs], dtype=_DTYPES[int(ct)])),
                                Mask.new([p[1] for p in pairs]))

    def buffer(self) -> CellBuffer: return self._buf
    def buffer_mut(self) -> CellBuffer: return self._buf
    def mask(self) -> Mask: return self._mask
    def mask_mut(self) -> Mask: return self._mask
    def len(self): return self._buf.len()
    __len__ = len
    def is_empty(self): return self.len() == 0
    def cell_type(self): return self._buf.cell_type()
    def get(self, index): return self._buf.get(index)
    def put(self, index, value): self._buf.put(index, value)

    def get_masked(self, index: int):
        return self._buf.get(index) if self._mask.get(index) else None

    def get_with_mask(self, index: int):
        return self._buf.get(index), self._mask.get(index)

    def put_with_mask(self, index: int, value, mask: bool) -> None:
        self._buf.put(index, value)
        self._mask.put(index, mask)

    def extend(self, pairs) -> None:
        """Extend<(C, bool)> (src/masked/masked_buffer.rs:274-281). Python numbers follow Rust's literal defaults
        (int -> i32, float -> f64); numpy scalars keep their type."""
        pairs = list(pairs)
        if not pairs:
            return
        v0 = pairs[0][0]
        dt = v0.dtype if isinstance(v0, np.generic) else (np.int32 if isinstance(v0, int) else np.float64)
        self._buf.extend(np.array([p[0] for p in pairs], dtype=dt))
        self._mask.extend([p[1] for p in pairs])

    def counts(self):
        return self._mask.counts()

    def convert(self, cell_type: CellType) -> "MaskedCellBuffer":
        return MaskedCellBuffer(self._buf.convert(cell_type), self._mask.clone())

    def view(self, offset: int, len_: int) -> "MaskedCellBuffer":
        """A row strip of a resident masked raster: the buffer half shares the allocation (CellBuffer.view), the mask
        half is a slice; offset on a 128-cell boundary (ec_row_strip)."""
        return MaskedCellBuffer(self._buf.view(offset, len_), self._mask.slice(offset, len_))

    def min_max(self):
        mn, mx = Value(), Value()
        check(lib().ec_buf_min_max(self._buf._h, self._mask._h, C.byref(mn), C.byref(mx)))
        return CellValue._wrap(mn), CellValue._wrap(mx)

    def statistics(self) -> "Statistics":
        """statistics of the valid cells (extension, see CellBuffer.statistics)"""
        return _statistics(self._buf._h, self._mask._h)

    def to_vec(self, cell_type=None):
        return self._buf.to_vec(cell_type)

    def to_vec_with_nodata(self, no_data: NoData, out: np.ndarray | None = None) -> np.ndarray:
        """to_vec_with_nodata::<T> (src/masked/masked_buffer.rs:137-152): convert + fill fused, one D2H."""
        h = C.c_void_p()
        check(lib().ec_buf_fill_nodata(self._buf._h, self._mask._h, int(no_data.ct), no_data.kind, no_data._ptr(), C.byref(h)))
        filled = CellBuffer._take(h)
        if filled.len() == 0:
            return np.empty(0, dtype=no_data.ct.dtype)
        return filled.to_vec(out=out)

    def __iter__(self):
        return iter(zip(self._buf, self._mask))

    def _bin(self, op: int, rhs) -> "MaskedCellBuffer":
        if isinstance(rhs, MaskedCellBuffer):
            hb, hm = C.c_void_p(), C.c_void_p()
            check(lib().ec_masked_binary(op, self._buf._h, self._mask._h, rhs._buf._h, rhs._mask._h, C.byref(hb), C.byref(hm)))
            return MaskedCellBuffer(CellBuffer._take(hb), Mask._take(hm))
        return MaskedCellBuffer(self._buf._bin(op, rhs), self._mask.clone())

    def __add__(self, r): return self._bin(0, r)
    def __sub__(self, r): return self._bin(1, r)
    def __mul__(self, r): return self._bin(2, r)
    def __truediv__(self, r): return self._bin(3, r)
    def __neg__(self): return MaskedCellBuffer(-self._buf, self._mask.clone())

    def cmp(self, other) -> int:
        """derive(PartialOrd) over (CellBuffer, Mask)."""
        c = self._buf.cmp(other._buf)
        return c if c else self._mask.cmp(other._mask)

    def __eq__(self, other):
        if isinstance(other, CellBuffer):
            other = MaskedCellBuffer(other, Mask.fill(other.len(), True))
        if not isinstance(other, MaskedCellBuffer):
            return NotImplemented
        return self.cmp(other) == 0

    __hash__ = None

    def __repr__(self):
        return f"{self.cell_type()}MaskedCellBuffer({self._buf!r}, {self._mask!r})"


MaskedCellBuffer.new = staticmethod(lambda buffer, mask: MaskedCellBuffer(buffer, mask))


# ---------------------------------------------------------------------------------------------
# serde wire format: the derives of src/ctype.rs:15, src/value.rs:16, src/buffer.rs:51,
# src/masked/mask.rs:11, src/masked/masked_buffer.rs:40, src/masked/nodata.rs:8 with serde's default
# representations — externally tagged enums, newtype struct = inner value, tuple struct = sequence. The reference
# holds no test for it (parity unpinned); this follows serde_json's documented behaviour, non-finite floats as null.
# ---------------------------------------------------------------------------------------------
def _jnum(x):
    x = x.item() if hasattr(x, "item") else x
    if isinstance(x, float) and (x != x or x in (float("inf"), float("-inf"))):
        return None
    return x


def to_serde(obj):
    """A JSON-ready python structure in serde's default layout."""
    if isinstance(obj, CellType):
        return str(obj)
    if isinstance(obj, CellValue):
        return {str(obj.cell_type()): _jnum(obj.value())}
    if isinstance(obj, CellBuffer):
        return {str(obj.cell_type()): [_jnum(x) for x in obj.to_vec()]}
    if isinstance(obj, Mask):
        return [bool(b) for b in obj.to_vec()]
    if isinstance(obj, MaskedCellBuffer):
        return [to_serde(obj.buffer()), to_serde(obj.mask())]
    if isinstance(obj, NoData):
        return {NoData.NONE: "None", NoData.DEFAULT: "Default"}.get(obj.kind) or {"Value": _jnum(obj.value())}
    raise TypeError(type(obj).__name__)


def from_serde(kind, data):
    """Inverse of to_serde for kind in (CellType, CellValue, CellBuffer, Mask, MaskedCellBuffer)."""
    if kind is CellType:
        return CellType.from_str(data)
    if kind is CellValue:
        ((name, v),) = data.items()
        return CellValue(CellType.from_str(name), float("nan") if v is None else v)
    if kind is CellBuffer:
        ((name, vals),) = data.items()
        ct = CellType.from_str(name)
        return CellBuffer.from_vec(np.array([float("nan") if v is None else v for v in vals], dtype=ct.dtype))
    if kind is Mask:
        return Mask.new(data)
    if kind is MaskedCellBuffer:
        return MaskedCellBuffer(from_serde(CellBuffer, data[0]), from_serde(Mask, data[1]))
    raise TypeError(kind)
