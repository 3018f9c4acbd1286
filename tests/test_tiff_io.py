"""CPU: the TIFF container the reference's stereo stage hands height rasters over in
(5-out-F.tif, disparity.py:213-224 / utils.py:37-51) -- reader, writer, stacking."""
import importlib
import struct

import numpy as np
import pytest

tio = importlib.import_module("3d-point-cloud-multiday-imagery_b200.tiff_io")


def _raster(h, w, seed=0):
    rs = np.random.RandomState(seed)
    out = np.zeros((3, h, w), dtype=np.float32)
    out[0] = rs.uniform(-150, 150, size=(h, w))
    out[2] = rs.rand(h, w) > 0.2
    return out


@pytest.mark.parametrize("kw", [
    {}, {"planar": True}, {"rows_per_strip": 1}, {"rows_per_strip": 7, "planar": True}, {"tile": 16},
    {"tile": 16, "planar": True}, {"big": True}, {"byteorder": ">"}, {"big": True, "byteorder": ">", "tile": 32},
])
def test_roundtrip_layouts(tmp_path, kw):
    bands = _raster(37, 53, seed=len(kw))
    p = str(tmp_path / "r.tif")
    tio.write_tiff(p, bands, **kw)
    a = tio.read_tiff(p)
    assert a.shape == (37, 53, 3) and a.dtype == np.float32
    np.testing.assert_array_equal(np.moveaxis(a, 2, 0), bands)


def test_default_layout_is_zero_copy_and_gdal_like(tmp_path):
    bands = _raster(64, 100)
    p = str(tmp_path / "5-out-F.tif")
    tio.write_tiff(p, bands)
    a = tio.read_tiff(p)
    base = a
    while getattr(base, "base", None) is not None and not isinstance(base, np.memmap):
        base = base.base
    assert isinstance(base, np.memmap)  # pixel-interleaved strips back to back: no copy
    raw = open(p, "rb").read()
    assert raw[:4] == b"II*\x00"
    # first pixel right after the 8-byte header: (height, 0, defined) as little-endian floats
    assert struct.unpack_from("<3f", raw, 8) == (bands[0, 0, 0], 0.0, bands[2, 0, 0])


def test_hand_built_single_band_uint16():
    # a 2x2 uint16 image assembled byte by byte (TIFF 6.0, section 2)
    import os
    import tempfile

    pix = struct.pack("<4H", 1, 2, 3, 65535)
    ifd_off = 8 + len(pix)
    ents = [(256, 3, 1, 2), (257, 3, 1, 2), (258, 3, 1, 16), (259, 3, 1, 1), (262, 3, 1, 1), (273, 4, 1, 8),
            (277, 3, 1, 1), (278, 3, 1, 2), (279, 4, 1, len(pix))]
    ifd = struct.pack("<H", len(ents))
    for tag, typ, cnt, val in ents:
        ifd += struct.pack("<HHI", tag, typ, cnt) + (struct.pack("<H", val) + b"\0\0" if typ == 3 else struct.pack("<I", val))
    ifd += struct.pack("<I", 0)
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "h.tif")
        open(p, "wb").write(b"II" + struct.pack("<HI", 42, ifd_off) + pix + ifd)
        a = tio.read_tiff(p)
    assert a.dtype == np.uint16 and a.tolist() == [[1, 2], [3, 65535]]


def test_rejects_compressed_and_garbage(tmp_path):
    p = str(tmp_path / "x.tif")
    open(p, "wb").write(b"not a tiff at all")
    with pytest.raises(tio.TiffError):
        tio.read_tiff(p)
    tio.write_tiff(p, _raster(4, 4))
    raw = bytearray(open(p, "rb").read())
    i = raw.find(struct.pack("<HHI", 259, 3, 1))
    raw[i + 8:i + 10] = struct.pack("<H", 5)  # LZW
    open(p, "wb").write(raw)
    with pytest.raises(tio.TiffError, match="compressed"):
        tio.read_tiff(p)


def test_stacking_pads_with_invalid_pixels(tmp_path):
    a, b = _raster(10, 12, 1), _raster(8, 15, 2)
    pa, pb = str(tmp_path / "a.tif"), str(tmp_path / "b.tif")
    tio.write_tiff(pa, a)
    tio.write_tiff(pb, b, planar=True)
    st = tio.load_height_rasters([pa, pb])
    assert st.shape == (2, 10, 15, 3) and st.dtype == np.float32
    np.testing.assert_array_equal(st[0, :, :12, 0], a[0])
    np.testing.assert_array_equal(st[1, :8, :, 2], b[2])
    assert np.isnan(st[0, :, 12:, 0]).all() and (st[0, :, 12:, 2] == 0).all()
    assert np.isnan(st[1, 8:, :, 0]).all() and (st[1, 8:, :, 2] == 0).all()
    with pytest.raises(tio.TiffError):
        tio.write_tiff(pa, a[:2])
        tio.load_height_rasters([pa])


# ---- files written by INDEPENDENT TIFF writers (OpenCV's libtiff, Pillow), and our writer read
# ---- back by them: the container code is not only checked against itself ----
def test_reads_opencv_written_float32_rasters(tmp_path):
    cv2 = pytest.importorskip("cv2")
    rs = np.random.RandomState(5)
    # 3-band Float32, the sample layout of the reference's 5-out-F.tif (disparity.py:213-224)
    a = rs.uniform(-150, 150, size=(45, 70, 3)).astype(np.float32)
    a[..., 2] = rs.rand(45, 70) > 0.3
    a[3, 4, 0] = np.nan
    p = str(tmp_path / "cv3.tif")
    assert cv2.imwrite(p, a, [cv2.IMWRITE_TIFF_COMPRESSION, 1])  # 1 = none (GDAL's GTiff default)
    got = tio.read_tiff(p)
    assert got.shape == (45, 70, 3) and got.dtype == np.float32
    # OpenCV keeps channels in BGR order in memory and writes RGB: the file holds them reversed
    np.testing.assert_array_equal(got, a[..., ::-1])
    np.testing.assert_array_equal(got[..., ::-1], cv2.imread(p, cv2.IMREAD_UNCHANGED))
    # strips of a few rows each (libtiff's default strip size), not one block
    for rows in (1, 7):
        assert cv2.imwrite(p, a, [cv2.IMWRITE_TIFF_COMPRESSION, 1, cv2.IMWRITE_TIFF_ROWSPERSTRIP, rows])
        np.testing.assert_array_equal(tio.read_tiff(p), a[..., ::-1])
    # single band float32 / uint16 / uint8
    for arr in (a[..., 0].copy(), (rs.rand(33, 21) * 65535).astype(np.uint16), (rs.rand(9, 300) * 255).astype(np.uint8)):
        assert cv2.imwrite(p, arr, [cv2.IMWRITE_TIFF_COMPRESSION, 1])
        got = tio.read_tiff(p)
        assert got.dtype == arr.dtype
        np.testing.assert_array_equal(got, arr)
    # a compressed file (LZW; OpenCV only compresses integer samples) must be refused, not misread
    assert cv2.imwrite(p, (rs.rand(64, 64) * 255).astype(np.uint8), [cv2.IMWRITE_TIFF_COMPRESSION, 5])
    with pytest.raises(tio.TiffError, match="compressed"):
        tio.read_tiff(p)


def test_reads_pillow_written_rasters(tmp_path):
    Image = pytest.importorskip("PIL.Image")
    rs = np.random.RandomState(6)
    p = str(tmp_path / "pil.tif")
    f32 = rs.uniform(-150, 150, size=(40, 61)).astype(np.float32)
    Image.fromarray(f32).save(p)  # mode "F": one Float32 band, uncompressed strips
    got = tio.read_tiff(p)
    assert got.dtype == np.float32
    np.testing.assert_array_equal(got, f32)
    u16 = (rs.rand(25, 33) * 65535).astype(np.uint16)
    Image.fromarray(u16).save(p)  # mode "I;16"
    np.testing.assert_array_equal(tio.read_tiff(p), u16)
    rgb = (rs.rand(19, 27, 3) * 255).astype(np.uint8)
    Image.fromarray(rgb).save(p)  # chunky 8-bit RGB
    np.testing.assert_array_equal(tio.read_tiff(p), rgb)
    Image.fromarray(f32).save(p, tiffinfo={278: 3})  # RowsPerStrip = 3
    np.testing.assert_array_equal(tio.read_tiff(p), f32)
    Image.fromarray(f32).save(p, compression="tiff_lzw")
    with pytest.raises(tio.TiffError, match="compressed"):
        tio.read_tiff(p)


@pytest.mark.parametrize("kw", [{}, {"planar": True}, {"rows_per_strip": 5}, {"tile": 16}, {"big": True},
                                {"byteorder": ">"}])
def test_our_writer_is_readable_by_opencv_and_pillow(tmp_path, kw):
    """The GPU tests feed the device path with files from tio.write_tiff: an independent reader
    (libtiff through OpenCV; Pillow for single-band files) must see the same pixels."""
    cv2 = pytest.importorskip("cv2")
    Image = pytest.importorskip("PIL.Image")
    bands = _raster(37, 53, seed=3)
    p = str(tmp_path / "w.tif")
    tio.write_tiff(p, bands, **kw)
    back = cv2.imread(p, cv2.IMREAD_UNCHANGED)
    assert back is not None and back.shape == (37, 53, 3) and back.dtype == np.float32
    if not kw.get("planar"):  # OpenCV's decoder ignores PlanarConfiguration=2 (reads the planes as chunky pixels)
        np.testing.assert_array_equal(back[..., ::-1], np.moveaxis(bands, 0, 2))  # OpenCV returns BGR
    tio.write_tiff(p, bands[0], **kw)
    np.testing.assert_array_equal(np.asarray(Image.open(p)), bands[0])
    np.testing.assert_array_equal(cv2.imread(p, cv2.IMREAD_UNCHANGED), bands[0])
