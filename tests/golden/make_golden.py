"""Regenerates tests/golden/landsat_elkton.npz from the reference's TIFF fixtures.

Run in the build container only (it reads /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py
The three files are the Landsat-8 fixtures of /root/reference/testkit/data used by the reference's
NDVI tests (src/gdal/rasterband.rs:138-191): single band, UInt16, uncompressed, stripped,
GDAL_NODATA tag "0". They are decoded with the minimal baseline-TIFF reader below (no GDAL here).
"""
import os
import struct

import numpy as np

SRC = "/root/reference/testkit/data"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "landsat_elkton.npz")


def read_tiff_u16(path):
    b = open(path, "rb").read()
    bo = {b"II": "<", b"MM": ">"}[b[:2]]
    assert struct.unpack(bo + "H", b[2:4])[0] == 42
    (ifd,) = struct.unpack(bo + "I", b[4:8])
    (n,) = struct.unpack(bo + "H", b[ifd:ifd + 2])
    tags = {}
    tsz = {1: 1, 2: 1, 3: 2, 4: 4, 5: 8, 12: 8, 16: 8}
    tfmt = {1: "B", 2: "c", 3: "H", 4: "I", 12: "d", 16: "Q"}
    for i in range(n):
        e = b[ifd + 2 + 12 * i: ifd + 14 + 12 * i]
        tag, typ, cnt = struct.unpack(bo + "HHI", e[:8])
        size = tsz[typ] * cnt
        data = e[8:8 + size] if size <= 4 else b[struct.unpack(bo + "I", e[8:12])[0]:][:size]
        if typ in tfmt:
            tags[tag] = struct.unpack(bo + tfmt[typ] * cnt, data)
    width, height = tags[256][0], tags[257][0]
    assert tags[258] == (16,) and tags[259] == (1,) and tags.get(277, (1,)) == (1,)  # u16, uncompressed, 1 band
    assert tags.get(339, (1,)) == (1,)  # unsigned integer samples
    offs, cnts = tags[273], tags[279]
    raw = b"".join(b[o:o + c] for o, c in zip(offs, cnts))
    a = np.frombuffer(raw, dtype=bo + "u2").astype(np.uint16).reshape(height, width)
    nodata = b"".join(tags[42113]).split(b"\0")[0].decode() if 42113 in tags else None
    return a, nodata


if __name__ == "__main__":
    red, nd_r = read_tiff_u16(os.path.join(SRC, "L8-Elkton-VA-B4.tiff"))
    nir, nd_n = read_tiff_u16(os.path.join(SRC, "L8-Elkton-VA-B5.tiff"))
    nir_nd, nd_nn = read_tiff_u16(os.path.join(SRC, "L8-Elkton-VA-B5-nd.tiff"))
    print("shape", red.shape, "GDAL_NODATA", nd_r, nd_n, nd_nn)
    np.savez_compressed(OUT, red=red, nir=nir, nir_nd=nir_nd, gdal_nodata=np.array([float(nd_nn)]))
    ndvi = (nir.astype("f8") - red.astype("f8")) / (nir.astype("f8") + red.astype("f8"))
    print("ndvi min", float(ndvi.min()).hex(), "max", float(ndvi.max()).hex(), "nir_nd zeros", int((nir_nd == 0).sum()))
    print("wrote", OUT, os.path.getsize(OUT), "bytes")
