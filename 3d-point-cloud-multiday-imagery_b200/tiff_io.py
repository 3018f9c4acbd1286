"""Minimal TIFF reader / writer for the height rasters on either side of the fused path.

The reference hands the per-pair height map from the stereo stage to the point-cloud stage as
``5-out-F.tif`` (``members/rafael/disparity/disparity.py:213-224``, written through
``utils.save_tiff_file``, ``utils.py:45-51``): a GDAL GTiff with three Float32 bands --
band 0 = height (``-disparity / 16``), band 1 unused, band 2 = ``final_defined`` -- and reads
rasters back with ``gdal.Open(...).ReadAsArray()`` (``utils.py:37-42``).  GDAL is not part of
this image, so this module restates the container format itself (TIFF 6.0 baseline / BigTIFF,
uncompressed, strips or tiles, chunky or planar samples), which is all GDAL's GTiff driver
produces with the default creation options the reference uses.

``read_tiff`` returns the samples pixel-interleaved ``[H, W, S]`` -- the layout
``mdkm_unproject(..., MDKM_HM_F32_GTIFF3, ...)`` consumes directly on the device, so a default
GDAL file (pixel-interleaved strips stored back to back) is passed on without a host-side
copy of the pixels.
"""
from __future__ import annotations

import struct
from typing import Dict, List, Sequence, Tuple

import numpy as np

_TYPES = {1: ("B", 1), 2: ("c", 1), 3: ("H", 2), 4: ("I", 4), 5: ("II", 8), 6: ("b", 1), 7: ("B", 1),
          8: ("h", 2), 9: ("i", 4), 10: ("ii", 8), 11: ("f", 4), 12: ("d", 8), 16: ("Q", 8), 17: ("q", 8), 18: ("Q", 8)}

# tags
IMAGE_WIDTH, IMAGE_LENGTH, BITS_PER_SAMPLE, COMPRESSION, PHOTOMETRIC = 256, 257, 258, 259, 262
STRIP_OFFSETS, SAMPLES_PER_PIXEL, ROWS_PER_STRIP, STRIP_BYTE_COUNTS = 273, 277, 278, 279
PLANAR_CONFIG, TILE_WIDTH, TILE_LENGTH, TILE_OFFSETS, TILE_BYTE_COUNTS = 284, 322, 323, 324, 325
EXTRA_SAMPLES, SAMPLE_FORMAT = 338, 339


class TiffError(ValueError):
    pass


def _read_ifd(buf, bo: str) -> Dict[int, Tuple]:
    magic = struct.unpack_from(bo + "H", buf, 2)[0]
    if magic == 42:
        big, off = False, struct.unpack_from(bo + "I", buf, 4)[0]
        n = struct.unpack_from(bo + "H", buf, off)[0]
        ent0, ent_sz, cnt_fmt, inl = off + 2, 12, "I", 4
    elif magic == 43:
        big, off = True, struct.unpack_from(bo + "Q", buf, 8)[0]
        n = struct.unpack_from(bo + "Q", buf, off)[0]
        ent0, ent_sz, cnt_fmt, inl = off + 8, 20, "Q", 8
    else:
        raise TiffError("not a TIFF file (bad magic)")
    tags = {}
    for i in range(n):
        e = ent0 + i * ent_sz
        tag, typ = struct.unpack_from(bo + "HH", buf, e)
        count = struct.unpack_from(bo + cnt_fmt, buf, e + 4)[0]
        if typ not in _TYPES:
            continue
        fmt, size = _TYPES[typ]
        voff = e + 4 + (8 if big else 4)
        if size * count > inl:
            voff = struct.unpack_from(bo + ("Q" if big else "I"), buf, voff)[0]
        if typ == 2:
            tags[tag] = (bytes(buf[voff:voff + count]),)
        else:
            tags[tag] = struct.unpack_from(bo + fmt * count, buf, voff)
    return tags


def _dtype(tags, bo) -> np.dtype:
    bits = set(tags.get(BITS_PER_SAMPLE, (1,)))
    fmt = set(tags.get(SAMPLE_FORMAT, (1,)))
    if len(bits) != 1 or len(fmt) != 1:
        raise TiffError("samples of different types are not supported")
    b, f = bits.pop(), fmt.pop()
    kind = {1: "u", 2: "i", 3: "f"}.get(f)
    if kind is None or b not in (8, 16, 32, 64) or (kind == "f" and b < 32):
        raise TiffError(f"unsupported sample format {f} / {b} bits")
    return np.dtype(("<" if bo == "<" else ">") + kind + str(b // 8))


def read_tiff(path: str, mmap: bool = True) -> np.ndarray:
    """Read an uncompressed TIFF / BigTIFF.  Returns ``[H, W, S]`` (``[H, W]`` for one sample) in
    the file's sample type, native byte order.  Zero-copy (``np.memmap``) when the file stores
    pixel-interleaved strips back to back, which is GDAL's default layout."""
    raw = np.memmap(path, dtype=np.uint8, mode="r") if mmap else np.fromfile(path, dtype=np.uint8)
    buf = memoryview(raw)
    if len(buf) < 8:
        raise TiffError("file too short")
    bo = {b"II": "<", b"MM": ">"}.get(bytes(buf[:2]))
    if bo is None:
        raise TiffError("not a TIFF file (byte order mark)")
    tags = _read_ifd(buf, bo)
    if tags.get(COMPRESSION, (1,))[0] != 1:
        raise TiffError("compressed TIFFs are not supported (GDAL's default GTiff is uncompressed)")
    W, H = int(tags[IMAGE_WIDTH][0]), int(tags[IMAGE_LENGTH][0])
    S = int(tags.get(SAMPLES_PER_PIXEL, (1,))[0])
    planar = int(tags.get(PLANAR_CONFIG, (1,))[0])
    dt = _dtype(tags, bo)
    isz = dt.itemsize
    native = dt.newbyteorder("=")

    def block(offset, count):  # a view of the (memory-mapped) file, never a copy
        return raw[offset:offset + count * isz].view(dt)

    def finish(a):
        a = a if a.dtype == native else a.astype(native)
        return a[:, :, 0] if S == 1 else a

    if TILE_OFFSETS in tags:
        tw, th = int(tags[TILE_WIDTH][0]), int(tags[TILE_LENGTH][0])
        offs = tags[TILE_OFFSETS]
        tx, ty = (W + tw - 1) // tw, (H + th - 1) // th
        out = np.empty((H, W, S), dtype=native)
        planes = S if planar == 2 else 1
        spp = 1 if planar == 2 else S
        for pl in range(planes):
            for j in range(ty):
                for i in range(tx):
                    o = offs[(pl * ty + j) * tx + i]
                    t = block(o, tw * th * spp).reshape(th, tw, spp)
                    h, w = min(th, H - j * th), min(tw, W - i * tw)
                    if planar == 2:
                        out[j * th:j * th + h, i * tw:i * tw + w, pl] = t[:h, :w, 0]
                    else:
                        out[j * th:j * th + h, i * tw:i * tw + w, :] = t[:h, :w, :]
        return finish(out)

    offs = tags[STRIP_OFFSETS]
    rps = min(int(tags.get(ROWS_PER_STRIP, (H,))[0]), H)
    n_strips = (H + rps - 1) // rps
    if planar == 1:
        row_bytes = W * S * isz
        contiguous = all(offs[s + 1] - offs[s] == rps * row_bytes for s in range(n_strips - 1))
        if contiguous:
            a = block(offs[0], H * W * S).reshape(H, W, S)
            return finish(a)
        out = np.empty((H, W, S), dtype=native)
        for s in range(n_strips):
            r0 = s * rps
            r = min(rps, H - r0)
            out[r0:r0 + r] = block(offs[s], r * W * S).reshape(r, W, S)
        return finish(out)
    out = np.empty((H, W, S), dtype=native)
    for pl in range(S):
        for s in range(n_strips):
            r0 = s * rps
            r = min(rps, H - r0)
            out[r0:r0 + r, :, pl] = block(offs[pl * n_strips + s], r * W).reshape(r, W)
    return finish(out)


def write_tiff(path: str, bands: np.ndarray, planar: bool = False, rows_per_strip: int | None = None,
               tile: int | None = None, big: bool = False, byteorder: str = "<") -> None:
    """Write ``bands[S, H, W]`` (the argument order of the reference's ``save_tiff_file``,
    ``utils.py:45-51``) as an uncompressed TIFF.  Defaults mirror GDAL's GTiff driver:
    pixel-interleaved samples, strips of about 8 KB."""
    a = np.asarray(bands)
    if a.ndim == 2:
        a = a[None]
    S, H, W = a.shape
    kind = {"f": 3, "u": 1, "i": 2}[a.dtype.kind]
    isz = a.dtype.itemsize
    dt = a.dtype.newbyteorder(byteorder)
    chunky = np.ascontiguousarray(np.moveaxis(a, 0, 2)).astype(dt, copy=False)
    blocks: List[bytes] = []
    if tile:
        tx, ty = (W + tile - 1) // tile, (H + tile - 1) // tile
        for pl in range(S if planar else 1):
            for j in range(ty):
                for i in range(tx):
                    t = np.zeros((tile, tile, 1 if planar else S), dtype=dt)
                    src = chunky[j * tile:(j + 1) * tile, i * tile:(i + 1) * tile]
                    t[:src.shape[0], :src.shape[1]] = src[:, :, pl:pl + 1] if planar else src
                    blocks.append(t.tobytes())
    else:
        if rows_per_strip is None:
            rows_per_strip = max(1, 8192 // max(1, W * isz * (1 if planar else S)))
        rows_per_strip = min(rows_per_strip, H)
        for pl in range(S if planar else 1):
            for r0 in range(0, H, rows_per_strip):
                src = chunky[r0:r0 + rows_per_strip]
                blocks.append((np.ascontiguousarray(src[:, :, pl]) if planar else src).tobytes())
    bo = byteorder
    hdr = 16 if big else 8
    offsets, pos = [], hdr
    for b in blocks:
        offsets.append(pos)
        pos += len(b)
    pos += pos & 1
    off_t, off_f = (16, "Q") if big else (4, "I")
    entries = [
        (IMAGE_WIDTH, 4, (W,)), (IMAGE_LENGTH, 4, (H,)), (BITS_PER_SAMPLE, 3, (isz * 8,) * S),
        (COMPRESSION, 3, (1,)), (PHOTOMETRIC, 3, (1,)), (SAMPLES_PER_PIXEL, 3, (S,)),
        (PLANAR_CONFIG, 3, (2 if planar else 1,)), (SAMPLE_FORMAT, 3, (kind,) * S),
    ]
    if S > 1:  # MINISBLACK has one colour channel: the other bands are "unspecified" extra samples (as GDAL writes)
        entries.append((EXTRA_SAMPLES, 3, (0,) * (S - 1)))
    if tile:
        entries += [(TILE_WIDTH, 3, (tile,)), (TILE_LENGTH, 3, (tile,)), (TILE_OFFSETS, off_t, tuple(offsets)),
                    (TILE_BYTE_COUNTS, off_t, tuple(len(b) for b in blocks))]
    else:
        entries += [(ROWS_PER_STRIP, 3, (rows_per_strip,)), (STRIP_OFFSETS, off_t, tuple(offsets)),
                    (STRIP_BYTE_COUNTS, off_t, tuple(len(b) for b in blocks))]
    entries.sort()
    ent_sz, inl = (20, 8) if big else (12, 4)
    ifd_off = pos
    ifd_len = (8 if big else 2) + len(entries) * ent_sz + (8 if big else 4)
    extra_off = ifd_off + ifd_len
    ifd = struct.pack(bo + ("Q" if big else "H"), len(entries))
    extra = b""
    for tag, typ, vals in entries:
        fmt, size = _TYPES[typ]
        data = struct.pack(bo + fmt * len(vals), *vals)
        ifd += struct.pack(bo + "HH" + ("Q" if big else "I"), tag, typ, len(vals))
        if len(data) <= inl:
            ifd += data.ljust(inl, b"\0")
        else:
            ifd += struct.pack(bo + off_f, extra_off + len(extra))
            extra += data + (b"\0" if len(data) & 1 else b"")
    ifd += struct.pack(bo + ("Q" if big else "I"), 0)
    with open(path, "wb") as f:
        f.write(b"II" if bo == "<" else b"MM")
        if big:
            f.write(struct.pack(bo + "HHHQ", 43, 8, 0, ifd_off))
        else:
            f.write(struct.pack(bo + "HI", 42, ifd_off))
        for b in blocks:
            f.write(b)
        f.write(b"\0" * (ifd_off - f.tell()))
        f.write(ifd)
        f.write(extra)


def load_height_rasters(paths: Sequence[str]) -> np.ndarray:
    """Stack the reference's per-pair ``5-out-F.tif`` rasters into ``[D, H, W, 3]`` float32
    (pixel-interleaved: height, unused, final_defined).  Every pair lives in its own rectified
    frame (``disparity.py:192-204``), so the rasters may differ in size: smaller ones are padded
    at the bottom / right with (NaN, 0, 0), i.e. invalid pixels."""
    rasters = []
    for p in paths:
        a = read_tiff(p)
        if a.ndim != 3 or a.shape[2] != 3 or a.dtype != np.float32:
            raise TiffError(f"{p}: expected a 3-band Float32 raster (disparity.py:213-224), got {a.shape} {a.dtype}")
        rasters.append(a)
    if not rasters:
        raise ValueError("no rasters given")
    H = max(a.shape[0] for a in rasters)
    W = max(a.shape[1] for a in rasters)
    if len(rasters) == 1:
        return rasters[0][None]
    out = np.zeros((len(rasters), H, W, 3), dtype=np.float32)
    out[..., 0] = np.nan
    for d, a in enumerate(rasters):
        out[d, :a.shape[0], :a.shape[1]] = a
    return out
