"""Ingest staging and wire format: the reference's two GDAL tests
(src/gdal/rasterband.rs:138-191) end to end from TIFF files — rebuilt from the committed pixel fixtures with the same
layout as testkit/data (u16, uncompressed strips, GDAL_NODATA "0") — through read_cells / read_cells_masked."""
import json

import numpy as np
import pytest

import erased_cells_b200 as ec
from erased_cells_b200 import CellBuffer, CellType, CellValue, Mask, MaskedCellBuffer, NoData, raster_io


def test_tiff_roundtrip_and_nodata_conversion(tmp_path):
    """CPU: the reader/writer pair and the f64 -> NoData<T> conversion of src/gdal/mod.rs:47-70."""
    for dt in ("u1", "u2", "u4", "i2", "i4", "f4", "f8"):
        px = (np.arange(37 * 19).reshape(37, 19) % 251).astype(dt)
        p = str(tmp_path / f"t_{dt}.tiff")
        raster_io.write_tiff(p, px, nodata=0.0, rows_per_strip=8)
        got, nd = raster_io.read_tiff(p)
        assert got.dtype == px.dtype and np.array_equal(got, px) and nd == 0.0
    assert raster_io.nodata_from_gdal(None, CellType.UInt16).kind == NoData.NONE
    assert raster_io.nodata_from_gdal(0.0, CellType.UInt16).value() == 0
    assert raster_io.nodata_from_gdal(-9999.0, CellType.Float32).value() == np.float32(-9999.0)
    assert raster_io.nodata_from_gdal(3.7, CellType.Int16).value() == 3          # to_i16 truncates in-range floats
    with pytest.raises(raster_io.NoDataConversionError):
        raster_io.nodata_from_gdal(-1.0, CellType.UInt8)                          # out of range: None -> error
    with pytest.raises(raster_io.NoDataConversionError):
        raster_io.nodata_from_gdal(float("nan"), CellType.Int32)


@pytest.mark.gpu
def test_read_cells_ndvi(landsat, tmp_path):  # src/gdal/rasterband.rs:138-163
    for name in ("red", "nir"):
        raster_io.write_tiff(str(tmp_path / f"{name}.tiff"), landsat[name], nodata=0.0, rows_per_strip=24)
    red, nir = raster_io.read_cells(str(tmp_path / "red.tiff")), raster_io.read_cells(str(tmp_path / "nir.tiff"))
    assert red.cell_type() == CellType.UInt16 and red.len() == 169 * 186
    ndvi = (nir - red) / (nir + red)
    mn, mx = ndvi.min_max()
    assert mn.to_f64() - -0.1248899911993 < 1e-8 and mx.to_f64() - 0.66998345719859 < 1e-8
    assert float(mn.value()).hex() == "-0x1.ff8ca5bcc77dcp-4" and float(mx.value()).hex() == "0x1.5708125b0ed28p-1"


@pytest.mark.gpu
def test_read_cells_masked_ndvi(landsat, tmp_path):  # src/gdal/rasterband.rs:166-191
    raster_io.write_tiff(str(tmp_path / "red.tiff"), landsat["red"], nodata=0.0)
    raster_io.write_tiff(str(tmp_path / "nir_nd.tiff"), landsat["nir_nd"], nodata=0.0)
    red, nir = raster_io.read_cells_masked(str(tmp_path / "red.tiff")), raster_io.read_cells_masked(str(tmp_path / "nir_nd.tiff"))
    nir_data, nir_nodata = nir.counts()
    with ec.lazy():
        ndvi = (nir - red) / (nir + red)
    assert ndvi.counts() == (nir_data, nir_nodata) == (31430, 4)
    mn, mx = ndvi.min_max()
    assert float(mn.value()).hex() == "-0x1.ff8ca5bcc77dcp-4" and float(mx.value()).hex() == "0x1.5708125b0ed28p-1"
    # examples/gdal.rs: counts add up to the raster size
    assert nir_data + nir_nodata == 169 * 186


@pytest.mark.gpu
def test_serde_wire_format():
    b = CellBuffer.from_vec(np.array([1, 2, 3], dtype=np.uint8))
    assert json.dumps(ec.to_serde(b)) == '{"UInt8": [1, 2, 3]}'
    assert json.dumps(ec.to_serde(CellValue(CellType.Float32, 1.5))) == '{"Float32": 1.5}'
    assert ec.to_serde(CellType.Int16) == "Int16"
    m = MaskedCellBuffer(CellBuffer.from_vec(np.array([0.5, np.nan])), Mask.new([True, False]))
    wire = json.dumps(ec.to_serde(m))
    assert wire == '[{"Float64": [0.5, null]}, [true, false]]'
    back = ec.from_serde(MaskedCellBuffer, json.loads(wire))
    assert back.cell_type() == CellType.Float64 and back.mask() == m.mask() and back.get(0) == CellValue.new(0.5)
    assert ec.from_serde(CellBuffer, json.loads('{"UInt16": [7, 8]}')) == CellBuffer.from_vec(np.array([7, 8], np.uint16))
    assert ec.to_serde(NoData.default(CellType.UInt8)) == "Default" and ec.to_serde(NoData.new(CellType.Int16, 3)) == {"Value": 3}


@pytest.mark.gpu
def test_chunked_ingest_matches_one_shot_upload(orc, tmp_path):
    """ec_ingest_*: a band arriving in chunks (ragged last chunk, chunk sizes from one tile to many, every cell type,
    NoData None / Default / Value) gives the same buffer, mask and counts as from_vec / from_vec_with_nodata, and the same
    mask as the oracle; the streaming TIFF reader agrees with the whole-file reader."""
    rng = np.random.default_rng(11)
    for ct in CellType:
        n = 3 * 40960 + 777
        a = np.frombuffer(rng.bytes(n * ct.size_of()), dtype=ct.dtype).copy()
        sentinel = a[5]
        a[rng.integers(0, n, 500)] = sentinel
        for chunk in (128, 4096, 40960, 0):
            for nd, okind, oval in ((NoData.none(ct), orc.ND_NONE, None), (NoData.default(ct), orc.ND_DEFAULT, None),
                                    (NoData.new(ct, sentinel), orc.ND_VALUE, orc.value(int(ct), sentinel))):
                got = raster_io.ingest(a, nd, masked=True, chunk_cells=chunk)
                assert np.array_equal(got.buffer().to_vec().view(np.uint8), a.view(np.uint8))
                want = orc.mask_from_nodata(a, okind, oval)
                assert np.array_equal(got.mask().to_vec(), want), (ct, chunk, okind)
                assert got.counts() == orc.mask_counts(want)
        plain = raster_io.ingest(a, chunk_cells=1024)
        assert plain == CellBuffer.from_vec(a)
    assert raster_io.ingest(np.array([], dtype=np.int16), NoData.default(CellType.Int16), masked=True).len() == 0
    # the streaming TIFF reader: strips of 7 rows, chunks that straddle strips, big-endian-free fixture layout
    px = rng.integers(0, 65535, (301, 257)).astype(np.uint16)
    px[rng.random(px.shape) < 0.01] = 0
    p = str(tmp_path / "band.tiff")
    raster_io.write_tiff(p, px, nodata=0.0, rows_per_strip=7)
    whole, nd = raster_io.read_tiff(p)
    m = raster_io.read_cells_masked(p, chunk_cells=1024)
    assert np.array_equal(m.buffer().to_vec(), whole.reshape(-1)) and m.counts() == (int((px != 0).sum()), int((px == 0).sum()))
    assert raster_io.read_cells(p, wait=True) == CellBuffer.from_vec(px.reshape(-1))


@pytest.mark.gpu
def test_pageable_from_vec_and_to_vec_go_through_the_staging_threads():
    """from_vec / to_vec on a Vec<T> (src/buffer.rs:60-66, :175-188; numpy arrays are pageable memory like a Vec): copies
    of >= 16 MiB run through the staged path (8 MiB chunks, a pool of host threads); ragged sizes, every thread count,
    the driver's own path (0) and pinned memory must all give the same bits."""
    import torch
    from erased_cells_b200 import synth
    ec.set_host_copy_threads(1)
    try:
        for ct, n in ((CellType.UInt8, (16 << 20) + 1), (CellType.Float64, (3 << 20) + 77), (CellType.Int16, (20 << 20) - 3)):
            h = synth.host(ct, n, 0xC0B1 + int(ct))
            pinned = torch.empty(h.nbytes, dtype=torch.uint8).pin_memory().numpy().view(h.dtype)
            pinned[:] = h
            want = CellBuffer.from_vec(pinned)  # pinned memory: always the direct copy
            for threads in (0, 1, 2, 5, 12):
                ec.set_host_copy_threads(threads)
                b = CellBuffer.from_vec(h)
                assert b == want, (ct, threads)
                out = b.to_vec()
                assert out.dtype == h.dtype and np.array_equal(out.view(np.uint8), h.view(np.uint8)), (ct, threads)
                a = CellBuffer.from_vec(h, wait=False).wait()
                assert a == want, (ct, threads, "async")
                sink = np.zeros(n + 5, dtype=h.dtype)
                assert np.array_equal(b.to_vec(out=sink).view(np.uint8), h.view(np.uint8)) and not sink[n:].any()
        # several caller threads at once (ctypes drops the GIL): the transfers take turns on the copy pool chunk by chunk
        from concurrent.futures import ThreadPoolExecutor
        ec.set_host_copy_threads(4)
        arrays = [synth.host(CellType.UInt16, (9 << 20) + 1000 * k, 0xAB00 + k) for k in range(4)]

        def round_trip(a):
            for _ in range(3):
                if not np.array_equal(CellBuffer.from_vec(a).to_vec(), a):
                    return False
            return True

        with ThreadPoolExecutor(4) as pool:
            assert all(pool.map(round_trip, arrays))
        bools = synth.host(CellType.UInt8, (17 << 20) + 9, 0xB001) > 100  # a Vec<bool> of 17 Mi cells (src/masked/mask.rs:120-131)
        for threads in (0, 3, 12):
            ec.set_host_copy_threads(threads)
            m = Mask.new(bools)
            assert m.counts() == (int(bools.sum()), int((~bools).sum())) and np.array_equal(m.to_vec(), bools), threads
        assert ec.set_host_copy_threads(7) == 12
    finally:
        ec.set_host_copy_threads(-1)  # back to the default
