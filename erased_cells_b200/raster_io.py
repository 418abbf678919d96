"""Ingest / egress staging around the compute path: the stand-in for the reference's GDAL
adapter (``RasterBandEx::read_cells`` / ``read_cells_masked``, src/gdal/rasterband.rs:81-126) for the file format of
its own fixtures — baseline TIFF, one band, uncompressed strips, GDAL_NODATA tag. libgdal is not available here, so
only that subset is read; what matters for the path is what happens after the bytes are in host memory:

    pixels -> pinned staging -> asynchronous H2D on the upload stream -> NoData mask built on the device

and the GDAL nodata conversion ``f64 -> NoData<T>`` of src/gdal/mod.rs:47-70 (value-checked ``to_<p>()``).
"""
from __future__ import annotations

import struct

import numpy as np

from .api import CellBuffer, CellType, CellValue, MaskedCellBuffer, NoData

# GdalDataType -> CellType for the reduced set of src/gdal/mod.rs:14-26 (TIFF SampleFormat, BitsPerSample)
_TIFF_TYPES = {(1, 8): CellType.UInt8, (1, 16): CellType.UInt16, (1, 32): CellType.UInt32, (2, 16): CellType.Int16,
               (2, 32): CellType.Int32, (3, 32): CellType.Float32, (3, 64): CellType.Float64}


class UnsupportedCellTypeError(Exception):
    """Error::UnsupportedCellTypeError (src/error.rs:16-17)"""


class NoDataConversionError(Exception):
    """Error::NoDataConversionError (src/error.rs:22-23)"""


def nodata_from_gdal(value: float | None, ct: CellType) -> NoData:
    """`impl TryFrom<GdalND> for NoData<T>` (src/gdal/mod.rs:47-70): None -> NoData::None, else `nd.to_<p>()`."""
    if value is None:
        return NoData.none(ct)
    v = CellValue(CellType.Float64, float(value)).to_prim(ct)
    if v is None:
        raise NoDataConversionError(f"Unable to convert {value} into NoData<{ct.dtype.name}>::Value")
    return NoData.new(ct, v.value())


def read_tiff(path: str):
    """(pixels as a 2-D numpy array, GDAL_NODATA as float or None) of a baseline single-band strip TIFF."""
    b = open(path, "rb").read()
    bo = {b"II": "<", b"MM": ">"}[b[:2]]
    if struct.unpack(bo + "H", b[2:4])[0] != 42:
        raise ValueError("not a TIFF")
    (ifd,) = struct.unpack(bo + "I", b[4:8])
    (n,) = struct.unpack(bo + "H", b[ifd:ifd + 2])
    tsz = {1: 1, 2: 1, 3: 2, 4: 4, 5: 8, 12: 8, 16: 8}
    tfmt = {1: "B", 2: "c", 3: "H", 4: "I", 12: "d", 16: "Q"}
    tags = {}
    for i in range(n):
        e = b[ifd + 2 + 12 * i: ifd + 14 + 12 * i]
        tag, typ, cnt = struct.unpack(bo + "HHI", e[:8])
        size = tsz.get(typ, 1) * cnt
        data = e[8:8 + size] if size <= 4 else b[struct.unpack(bo + "I", e[8:12])[0]:][:size]
        if typ in tfmt:
            tags[tag] = struct.unpack(bo + tfmt[typ] * cnt, data)
    if tags.get(259, (1,)) != (1,) or tags.get(277, (1,)) != (1,):
        raise UnsupportedCellTypeError("only uncompressed single-band TIFFs")
    key = (tags.get(339, (1,))[0], tags[258][0])
    if key not in _TIFF_TYPES:
        raise UnsupportedCellTypeError(f"sample format {key}")
    ct = _TIFF_TYPES[key]
    width, height = tags[256][0], tags[257][0]
    raw = b"".join(b[o:o + c] for o, c in zip(tags[273], tags[279]))
    px = np.frombuffer(raw, dtype=ct.dtype.newbyteorder(bo)).astype(ct.dtype).reshape(height, width)
    nodata = None
    if 42113 in tags:
        txt = b"".join(tags[42113]).split(b"\0")[0].decode().strip()
        nodata = float(txt) if txt else None
    return px, nodata


def write_tiff(path: str, px: np.ndarray, nodata: float | None = None, rows_per_strip: int = 32) -> None:
    """Little-endian baseline strip TIFF with an optional GDAL_NODATA tag (to build fixtures)."""
    ct = CellType.of(px)
    fmt = {v: k for k, v in _TIFF_TYPES.items()}[ct]
    h, w = px.shape
    strips = [np.ascontiguousarray(px[r:r + rows_per_strip]).astype(ct.dtype.newbyteorder("<")).tobytes() for r in range(0, h, rows_per_strip)]
    nd = (repr(nodata).rstrip("0").rstrip(".") if nodata is not None and float(nodata).is_integer() else repr(nodata)).encode() + b"\0" if nodata is not None else None
    entries = [(256, 4, 1, w), (257, 4, 1, h), (258, 3, 1, fmt[1]), (259, 3, 1, 1), (262, 3, 1, 1), (277, 3, 1, 1), (278, 4, 1, rows_per_strip),
               (339, 3, 1, fmt[0])]
    ns = len(strips)
    n_entries = len(entries) + 2 + (1 if nd else 0)
    ifd_off = 8
    extra_off = ifd_off + 2 + 12 * n_entries + 4
    offs_off, cnts_off = extra_off, extra_off + 4 * ns
    nd_off = cnts_off + 4 * ns
    data_off = nd_off + (len(nd) if nd else 0)
    data_off += data_off % 2
    offs, pos = [], data_off
    for s in strips:
        offs.append(pos)
        pos += len(s)
    entries += [(273, 4, ns, offs_off if ns > 1 else offs[0]), (279, 4, ns, cnts_off if ns > 1 else len(strips[0]))]
    if nd:
        entries.append((42113, 2, len(nd), nd_off if len(nd) > 4 else int.from_bytes(nd.ljust(4, b"\0"), "little")))
    entries.sort()
    out = bytearray(b"II*\0" + struct.pack("<I", ifd_off) + struct.pack("<H", n_entries))
    for tag, typ, cnt, val in entries:
        out += struct.pack("<HHI", tag, typ, cnt) + (struct.pack("<H", val) + b"\0\0" if typ == 3 and cnt == 1 else struct.pack("<I", val))
    out += struct.pack("<I", 0)
    out += struct.pack(f"<{ns}I", *offs) + struct.pack(f"<{ns}I", *[len(s) for s in strips])
    if nd:
        out += nd
    out += b"\0" * (data_off - len(out))
    for s in strips:
        out += s
    open(path, "wb").write(bytes(out))


class Ingest:
    """A raster arriving chunk by chunk (ec_ingest_*): the reader fills pinned staging buffers in turn; the upload of one
    chunk, the reader's work on the next and the NoData compare of the previous one overlap. `nodata=None` with
    `masked=False` is RasterBandEx::read_cells, a NoData with `masked=True` is read_cells_masked."""

    def __init__(self, ct: CellType, len_: int, nodata: NoData | None = None, masked: bool = False, chunk_cells: int = 0):
        import ctypes as C

        from ._lib import check, lib
        self._C, self._check, self._lib = C, check, lib()
        self.ct, self.len, self.masked = CellType(ct), len_, masked
        nd = nodata if nodata is not None else NoData.none(self.ct)
        h = C.c_void_p()
        check(self._lib.ec_ingest_begin(int(self.ct), len_, nd.kind, nd._ptr(), int(masked), chunk_cells, C.byref(h)))
        self._h = h

    def next_buffer(self):
        """a numpy view of the next pinned staging buffer (None once every cell is in); fill a prefix, then submit(n)"""
        C = self._C
        p, cap = C.c_void_p(), C.c_size_t()
        self._check(self._lib.ec_ingest_next_buffer(self._h, C.byref(p), C.byref(cap)))
        if not cap.value:
            return None
        raw = (C.c_uint8 * (cap.value * self.ct.size_of())).from_address(p.value)
        return np.frombuffer(raw, dtype=self.ct.dtype)

    def submit(self, n_cells: int) -> None:
        self._check(self._lib.ec_ingest_submit(self._h, n_cells))

    def finish(self):
        C = self._C
        b, m = C.c_void_p(), C.c_void_p()
        self._check(self._lib.ec_ingest_finish(self._h, C.byref(b), C.byref(m) if self.masked else None))
        self._h = None
        buf = CellBuffer._take(b)
        if not self.masked:
            return buf
        from .api import Mask
        return MaskedCellBuffer(buf, Mask._take(m))

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.ec_ingest_abort(self._h)


def ingest(cells: np.ndarray, nodata: NoData | None = None, masked: bool = False, chunk_cells: int = 0):
    """Upload a host array (pageable is fine) through the chunked pipeline: each chunk is copied into pinned staging by the
    host while the previous one crosses PCIe and the one before gets its mask built."""
    a = np.ascontiguousarray(cells).reshape(-1)
    g = Ingest(CellType.of(a), a.size, nodata, masked, chunk_cells)
    pos = 0
    while True:
        buf = g.next_buffer()
        if buf is None:
            break
        n = min(buf.size, a.size - pos)
        buf[:n] = a[pos:pos + n]
        g.submit(n)
        pos += n
    return g.finish()


def _tiff_layout(path: str):
    """(cell type, width, height, [(file offset, byte count)] of the strips, byte order, GDAL_NODATA) without reading pixels"""
    with open(path, "rb") as f:
        b = f.read(1 << 20)  # header + IFD + strip tables of the fixtures live in the first MiB
    bo = {b"II": "<", b"MM": ">"}[b[:2]]
    (ifd,) = struct.unpack(bo + "I", b[4:8])
    (n,) = struct.unpack(bo + "H", b[ifd:ifd + 2])
    tsz = {1: 1, 2: 1, 3: 2, 4: 4, 5: 8, 12: 8, 16: 8}
    tfmt = {1: "B", 2: "c", 3: "H", 4: "I", 12: "d", 16: "Q"}
    tags = {}
    for i in range(n):
        e = b[ifd + 2 + 12 * i: ifd + 14 + 12 * i]
        tag, typ, cnt = struct.unpack(bo + "HHI", e[:8])
        size = tsz.get(typ, 1) * cnt
        data = e[8:8 + size] if size <= 4 else b[struct.unpack(bo + "I", e[8:12])[0]:][:size]
        if typ in tfmt and len(data) == size:
            tags[tag] = struct.unpack(bo + tfmt[typ] * cnt, data)
    if tags.get(259, (1,)) != (1,) or tags.get(277, (1,)) != (1,):
        raise UnsupportedCellTypeError("only uncompressed single-band TIFFs")
    key = (tags.get(339, (1,))[0], tags[258][0])
    if key not in _TIFF_TYPES:
        raise UnsupportedCellTypeError(f"sample format {key}")
    nodata = None
    if 42113 in tags:
        txt = b"".join(tags[42113]).split(b"\0")[0].decode().strip()
        nodata = float(txt) if txt else None
    return _TIFF_TYPES[key], tags[256][0], tags[257][0], list(zip(tags[273], tags[279])), bo, nodata


def _read_streaming(path: str, masked: bool, chunk_cells: int):
    """the TIFF's strips straight into the pinned staging buffers of a chunked ingest: the file read of chunk k + 1 overlaps
    the PCIe copy of chunk k and the mask kernel of chunk k - 1 (no assembled host copy of the band)"""
    ct, width, height, strips, bo, nodata = _tiff_layout(path)
    nd = nodata_from_gdal(nodata, ct) if masked else None
    g = Ingest(ct, width * height, nd, masked, chunk_cells)
    sz = ct.size_of()
    with open(path, "rb", buffering=0) as f:
        si, so = 0, 0  # current strip, bytes of it already consumed
        while True:
            buf = g.next_buffer()
            if buf is None:
                break
            raw, filled = buf.view(np.uint8), 0
            while filled < raw.size and si < len(strips):
                off, cnt = strips[si]
                take = min(cnt - so, raw.size - filled)
                f.seek(off + so)
                got = f.readinto(memoryview(raw[filled:filled + take]))
                if got != take:
                    raise ValueError("truncated TIFF")
                filled += take
                so += take
                if so == cnt:
                    si, so = si + 1, 0
            n = filled // sz
            if bo == ">" and sz > 1:
                buf[:n] = buf[:n].byteswap()
            g.submit(n)
    return g.finish()


def read_cells(path: str, wait: bool = False, chunk_cells: int = 0) -> CellBuffer:
    """RasterBandEx::read_cells (src/gdal/rasterband.rs:82-103): the band as a device CellBuffer; nodata is ignored.
    Strips are streamed through pinned staging (Ingest); consumers are stream-ordered after the uploads."""
    b = _read_streaming(path, False, chunk_cells)
    return b.wait() if wait else b


def read_cells_masked(path: str, wait: bool = False, chunk_cells: int = 0) -> MaskedCellBuffer:
    """RasterBandEx::read_cells_masked (src/gdal/rasterband.rs:104-126): GDAL nodata -> NoData<T> -> mask built on the device
    chunk by chunk while the next chunk is still on its way."""
    m = _read_streaming(path, True, chunk_cells)
    if wait:
        m.buffer().wait()
    return m
