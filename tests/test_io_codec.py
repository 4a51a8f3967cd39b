"""libb2sio (include/b2sio.h): the native TIFF / .raw tile codec against Pillow / OpenCV-written files, plus the
symbols the header declares.  Host-only: runs without a GPU."""
import ctypes
import os
import re
from pathlib import Path

import numpy as np
import pytest

from pystripe import _io, raw

ROOT = Path(__file__).resolve().parents[1]


def _tile(h, w, dtype, seed=0):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w]
    a = (np.sin(x / 37.0) + np.cos(y / 23.0)) * 1000 + 2000 + rng.normal(0, 30, (h, w))
    if np.dtype(dtype) == np.uint8:
        a = a / 20
    return a.astype(dtype)


def test_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "b2sio.h").read_text()
    declared = set(re.findall(r"\b(b2sio_[a-z_]+)\s*\(", header))
    assert declared == set(_io.EXPORTS), declared ^ set(_io.EXPORTS)
    lib = ctypes.CDLL(str(_io.LIB_PATH))
    for name in declared:
        assert hasattr(lib, name), name
    assert _io.lib().b2sio_version() == 100


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.float32])
@pytest.mark.parametrize("shape", [(37, 53), (300, 411)])
def test_reads_what_pillow_writes(tmp_path, dtype, shape):
    from PIL import Image
    img = _tile(*shape, dtype)
    for comp in (None, "tiff_lzw", "tiff_adobe_deflate", "tiff_deflate"):
        p = tmp_path / "a.tif"
        im = Image.fromarray(img)
        im.save(p, format="TIFF", **({"compression": comp} if comp else {}))
        got_shape, got_dtype, info = _io.probe(p)
        assert got_shape == shape and got_dtype == np.dtype(dtype)
        assert np.array_equal(_io.read(p), img), (comp, info.compression)
        assert np.array_equal(_io.read(p, threads=1), img)


def test_reads_lzw_with_horizontal_predictor(tmp_path):
    cv2 = pytest.importorskip("cv2")
    img = _tile(300, 411, np.uint16)
    p = tmp_path / "c.tif"
    cv2.imwrite(str(p), img)                       # OpenCV: LZW + predictor 2 for 16-bit samples
    assert _io.probe(p)[2].compression == 5
    assert np.array_equal(_io.read(p), img)


def test_reads_big_endian_and_tiled_and_bigtiff(tmp_path):
    """hand-built files: MM byte order with a horizontal predictor, a tiled layout, a BigTIFF directory."""
    import struct
    import zlib
    img = _tile(70, 90, np.uint16, seed=3)
    h, w = img.shape

    def classic(order, entries, blobs):
        # entries: (tag, type, count, value-or-blob-index); blobs placed after the header, IFD last
        e = "<" if order == "II" else ">"
        body = b""
        offs = []
        for b in blobs:
            offs.append(8 + len(body))
            body += b + (b"\0" if len(b) & 1 else b"")
        ifd = 8 + len(body)
        out = order.encode() + struct.pack(e + "HI", 42, ifd) + body + struct.pack(e + "H", len(entries))
        for tag, typ, count, val in entries:
            if isinstance(val, tuple):
                val = offs[val[0]]
            if typ == 3 and count == 1:
                out += struct.pack(e + "HHIHH", tag, typ, count, val, 0)
            else:
                out += struct.pack(e + "HHII", tag, typ, count, val)
        return out + struct.pack(e + "I", 0)

    # big-endian, deflate, predictor 2, two strips
    be = img.astype(">u2")
    diff = be.astype(np.uint16).copy()
    diff[:, 1:] = (img[:, 1:] - img[:, :-1]).astype(np.uint16)
    diff = diff.astype(">u2")
    s0, s1 = zlib.compress(diff[:40].tobytes()), zlib.compress(diff[40:].tobytes())
    offsets_blob_index, counts_blob_index = 2, 3
    blobs = [s0, s1, b"", b""]
    # offsets of the strips are known only after layout: strips come first, so compute them here
    o0, o1 = 8, 8 + len(s0) + (len(s0) & 1)
    blobs[2] = struct.pack(">II", o0, o1)
    blobs[3] = struct.pack(">II", len(s0), len(s1))
    ents = [(256, 3, 1, w), (257, 3, 1, h), (258, 3, 1, 16), (259, 3, 1, 8), (262, 3, 1, 1),
            (273, 4, 2, (offsets_blob_index,)), (277, 3, 1, 1), (278, 3, 1, 40), (279, 4, 2, (counts_blob_index,)),
            (317, 3, 1, 2)]
    p = tmp_path / "be.tif"
    p.write_bytes(classic("MM", ents, blobs))
    info = _io.probe(p)[2]
    assert info.big_endian == 1 and info.compression == 8 and info.n_chunks == 2
    assert np.array_equal(_io.read(p), img)

    # little-endian, tiled 32 x 48, uncompressed
    th, tw = 32, 48
    tiles = []
    for ty in range(0, h, th):
        for tx in range(0, w, tw):
            t = np.zeros((th, tw), np.uint16)
            blk = img[ty:ty + th, tx:tx + tw]
            t[:blk.shape[0], :blk.shape[1]] = blk
            tiles.append(t.tobytes())
    n = len(tiles)
    pos, toffs = 8, []
    for t in tiles:
        toffs.append(pos)
        pos += len(t)
    blobs = tiles + [struct.pack(f"<{n}I", *toffs), struct.pack(f"<{n}I", *[len(t) for t in tiles])]
    ents = [(256, 3, 1, w), (257, 3, 1, h), (258, 3, 1, 16), (259, 3, 1, 1), (262, 3, 1, 1), (277, 3, 1, 1),
            (322, 3, 1, tw), (323, 3, 1, th), (324, 4, n, (n,)), (325, 4, n, (n + 1,))]
    p = tmp_path / "tiled.tif"
    p.write_bytes(classic("II", ents, blobs))
    assert _io.probe(p)[2].tiled == 1
    assert np.array_equal(_io.read(p), img)

    # BigTIFF, one stored strip
    data = img.tobytes()
    ents8 = [(256, 4, 1, w), (257, 4, 1, h), (258, 3, 1, 16), (259, 3, 1, 1), (273, 16, 1, 16), (277, 3, 1, 1),
             (278, 4, 1, h), (279, 16, 1, len(data))]
    ifd = 16 + len(data)
    out = b"II" + struct.pack("<HHHQ", 43, 8, 0, ifd) + data + struct.pack("<Q", len(ents8))
    for tag, typ, count, val in ents8:
        out += struct.pack("<HHQQ", tag, typ, count, val)
    out += struct.pack("<Q", 0)
    p = tmp_path / "big.tif"
    p.write_bytes(out)
    assert np.array_equal(_io.read(p), img)


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.float32])
def test_writer_round_trips_through_pillow(tmp_path, dtype):
    from PIL import Image
    img = _tile(513, 300, dtype, seed=1)
    for comp in (None, ("ADOBE_DEFLATE", 1), ("ADOBE_DEFLATE", 6), "DEFLATE"):
        p = tmp_path / "b.tif"
        _io.write_tiff(p, img, comp)
        with Image.open(p) as im:
            assert np.array_equal(np.array(im), img), comp
        assert np.array_equal(_io.read(p), img)
        assert os.stat(p).st_mode & 0o777 == 0o777            # the reference chmods its output (core.py:311-314)
        assert not list(tmp_path.glob("*.b2s~"))
    assert _io.can_write(img, ("LZMA", 1)) is False and _io.can_write(img, ("ADOBE_DEFLATE", 1)) is True


def test_raw_tiles_both_byte_orders(tmp_path):
    img = _tile(120, 77, np.uint16, seed=2)
    raw.raw_imsave(tmp_path / "t.raw", img)
    assert np.array_equal(_io.read(tmp_path / "t.raw"), img)
    _io.write_raw(tmp_path / "u.raw", img)
    assert np.array_equal(np.asarray(raw.raw_imread(tmp_path / "u.raw")), img)
    with open(tmp_path / "be.raw", "wb") as f:
        np.array([77, 120], dtype=">u4").tofile(f)
        img.astype(">u2").tofile(f)
    assert _io.probe(tmp_path / "be.raw")[2].big_endian == 1
    assert np.array_equal(_io.read(tmp_path / "be.raw"), img)
    # the reference's reader returns a big-endian memmap for such a tile: the GPU path converts it at the boundary
    from pystripe import core
    arr, _ = core._as_supported(raw.raw_imread(tmp_path / "be.raw"))
    assert arr.dtype == np.uint16 and arr.dtype.isnative and np.array_equal(arr, img)


def test_reads_greyscale_png(tmp_path):
    """8 / 16-bit greyscale PNG (the reference reads them through imageio, core.py:209-210): every scanline filter, written
    by Pillow and by OpenCV; colour files are refused (the caller's general reader takes them)."""
    from PIL import Image
    cv2 = pytest.importorskip("cv2")
    for dtype in (np.uint8, np.uint16):
        img = _tile(211, 307, dtype, seed=6)
        p = tmp_path / "a.png"
        Image.fromarray(img).save(p, format="PNG")                 # Pillow picks filters per row (adaptive)
        shape, dt, info = _io.probe(p)
        assert shape == img.shape and dt == np.dtype(dtype) and info.n_chunks == 1
        assert np.array_equal(_io.read(p), img)
        for strategy in (cv2.IMWRITE_PNG_STRATEGY_DEFAULT, cv2.IMWRITE_PNG_STRATEGY_FILTERED, cv2.IMWRITE_PNG_STRATEGY_RLE):
            cv2.imwrite(str(p), img, [cv2.IMWRITE_PNG_STRATEGY, strategy, cv2.IMWRITE_PNG_COMPRESSION, 3])
            assert np.array_equal(_io.read(p), img)
    # hand-built file that uses each filter type once per five rows, IDAT split over several chunks
    import struct
    import zlib
    img = _tile(23, 31, np.uint16, seed=7)
    be = img.astype(">u2").view(np.uint8).reshape(23, 62).astype(np.int32)
    rows = []
    for y in range(23):
        ft = y % 5
        cur = be[y]
        up = be[y - 1] if y else np.zeros(62, np.int32)
        left = np.concatenate([np.zeros(2, np.int32), cur[:-2]])
        upleft = np.concatenate([np.zeros(2, np.int32), up[:-2]])
        if ft == 0:
            enc = cur
        elif ft == 1:
            enc = cur - left
        elif ft == 2:
            enc = cur - up
        elif ft == 3:
            enc = cur - ((left + up) >> 1)
        else:
            pp = left + up - upleft
            pa, pb, pc = abs(pp - left), abs(pp - up), abs(pp - upleft)
            pred = np.where((pa <= pb) & (pa <= pc), left, np.where(pb <= pc, up, upleft))
            enc = cur - pred
        rows.append(bytes([ft]) + (enc & 0xff).astype(np.uint8).tobytes())
    z = zlib.compress(b"".join(rows), 6)

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d))
    png = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", 31, 23, 16, 0, 0, 0, 0)) + \
        chunk(b"IDAT", z[:40]) + chunk(b"tEXt", b"k\0v") + chunk(b"IDAT", z[40:]) + chunk(b"IEND", b"")
    p = tmp_path / "filters.png"
    p.write_bytes(png)
    assert np.array_equal(_io.read(p), img)
    Image.fromarray(np.dstack([_tile(20, 20, np.uint8)] * 3)).save(tmp_path / "rgb.png")
    with pytest.raises(_io.CodecError) as e:
        _io.read(tmp_path / "rgb.png")
    assert e.value.code == _io.ERR_UNSUPPORTED


def test_batch_reports_per_file_status(tmp_path):
    imgs = np.stack([_tile(64, 80, np.uint16, seed=s) for s in range(5)])
    paths = [tmp_path / f"s{i}.tif" for i in range(5)]
    assert _io.write_tiff_batch(paths, imgs, ("ADOBE_DEFLATE", 1), threads=3) == [0] * 5
    _io.write_tiff(paths[1], _tile(64, 81, np.uint16))           # another shape
    paths[3].write_bytes(b"not a tiff at all")
    missing = tmp_path / "missing.tif"
    out = np.full((6, 64, 80), 7, np.uint16)
    st = _io.read_batch(paths + [missing], out, threads=4)
    assert st == [0, _io.ERR_SHAPE, 0, _io.ERR_FORMAT, 0, _io.ERR_IO]
    for i in (0, 2, 4):
        assert np.array_equal(out[i], imgs[i])
    for i in (1, 3, 5):
        assert (out[i] == 7).all()                               # failed files leave their plane untouched
    with pytest.raises(_io.CodecError) as e:
        _io.read(missing)
    assert e.value.code == _io.ERR_IO


def test_zstd_tiff_round_trip(tmp_path):
    """compression=('ZSTD', level) is the reference GUI's other output option (TIFF tag 50000); libzstd is bound at run time."""
    img = _tile(400, 333, np.uint16, seed=4)
    p = tmp_path / "z.tif"
    if not _io.can_write(img, ("ZSTD", 1)):
        pytest.skip("no libzstd")
    try:
        _io.write_tiff(p, img, ("ZSTD", 1))
    except _io.CodecError as e:
        if e.code == _io.ERR_UNSUPPORTED:
            pytest.skip("libzstd could not be loaded")
        raise
    assert _io.probe(p)[2].compression == 50000
    assert np.array_equal(_io.read(p), img) and np.array_equal(_io.read(p, threads=1), img)
    assert p.stat().st_size < img.nbytes
