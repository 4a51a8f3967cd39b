"""ctypes binding of libb2sio.so (include/b2sio.h): the native multithreaded TIFF / .raw tile codec.

Replaces the per-file tifffile / Pillow calls of the reference (pystripe/core.py:200-334) with batch decodes straight
into a caller-owned (pinned) buffer.  Files outside the codec's subset return a status the caller can fall back on
(Pillow), as the reference itself falls back from tifffile to Pillow (core.py:212-224).
"""
import ctypes as C
import os
from pathlib import Path

import numpy as np

_PKG = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("B2SIO_LIB", _PKG.parent / "lib" / "libb2sio.so"))

OK, ERR_IO, ERR_FORMAT, ERR_UNSUPPORTED, ERR_SHAPE, ERR_INVALID = 0, -1, -2, -3, -4, -5
EXPORTS = ("b2sio_version", "b2sio_last_error", "b2sio_probe", "b2sio_read", "b2sio_read_batch", "b2sio_write_tiff",
           "b2sio_write_tiff_batch", "b2sio_write_tiff_strips_batch", "b2sio_write_raw")
_CODES = {np.dtype(np.uint8): 0, np.dtype(np.uint16): 1, np.dtype(np.float32): 2}
_DTYPES = {0: np.uint8, 1: np.uint16, 2: np.float32}


class Info(C.Structure):
    _fields_ = [("height", C.c_int32), ("width", C.c_int32), ("dtype", C.c_int32), ("compression", C.c_int32),
                ("big_endian", C.c_int32), ("tiled", C.c_int32), ("n_chunks", C.c_int64)]


class CodecError(OSError):
    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise OSError(f"{LIB_PATH} not found — build it with `make -C image-preprocessing-pipeline_b200/csrc`")
        L = C.CDLL(str(LIB_PATH))
        vp, i32, cp = C.c_void_p, C.c_int32, C.c_char_p
        L.b2sio_version.restype = C.c_int
        L.b2sio_last_error.restype = cp
        L.b2sio_probe.argtypes = [cp, C.POINTER(Info)]
        L.b2sio_read.argtypes = [cp, vp, i32, i32, i32, C.c_int]
        L.b2sio_read_batch.argtypes = [C.POINTER(cp), C.c_int, vp, C.c_size_t, i32, i32, i32, C.c_int, C.POINTER(i32)]
        L.b2sio_write_tiff.argtypes = [cp, vp, i32, i32, i32, C.c_int, C.c_int]
        L.b2sio_write_tiff_batch.argtypes = [C.POINTER(cp), C.c_int, vp, C.c_size_t, i32, i32, i32, C.c_int, C.c_int,
                                             C.POINTER(i32)]
        L.b2sio_write_tiff_strips_batch.argtypes = [C.POINTER(cp), C.c_int, vp, vp, vp, i32, i32, i32, i32, i32, i32, C.c_int,
                                                    C.POINTER(i32)]
        L.b2sio_write_raw.argtypes = [cp, vp, i32, i32]
        _lib = L
    return _lib


def _err(code):
    return CodecError(code, (lib().b2sio_last_error() or b"").decode(errors="replace"))


def default_threads() -> int:
    return max(1, min(16, (os.cpu_count() or 2)))


def probe(path):
    """(shape, numpy dtype, Info) of a .tif / .tiff / .raw file without decoding it."""
    info = Info()
    rc = lib().b2sio_probe(os.fsencode(path), C.byref(info))
    if rc:
        raise _err(rc)
    return (info.height, info.width), np.dtype(_DTYPES[info.dtype]), info


def read(path, out=None, threads=None):
    """decode one file (native byte order); `out` may be a preallocated C-contiguous (pinned) array of the file's shape."""
    if out is None:
        shape, dtype, _ = probe(path)
        out = np.empty(shape, dtype)
    if not out.flags.c_contiguous or out.ndim != 2:
        raise ValueError("out must be a C-contiguous 2-D array")
    rc = lib().b2sio_read(os.fsencode(path), out.ctypes.data, out.shape[0], out.shape[1], _CODES[out.dtype],
                          int(threads or default_threads()))
    if rc:
        raise _err(rc)
    return out


def read_batch(paths, out, threads=None):
    """decode len(paths) files into out[i] (out: (n, H, W) C-contiguous).  Returns the per-file status list (0 = ok)."""
    n = len(paths)
    if out.ndim != 3 or out.shape[0] < n or not out.flags.c_contiguous:
        raise ValueError("out must be a C-contiguous (n, H, W) array holding at least len(paths) planes")
    arr = (C.c_char_p * n)(*[os.fsencode(p) for p in paths])
    status = (C.c_int32 * max(n, 1))()
    lib().b2sio_read_batch(arr, n, out.ctypes.data, out.strides[0], out.shape[1], out.shape[2], _CODES[out.dtype],
                           int(threads or default_threads()), status)
    return list(status[:n])


def _level(compression):
    """the reference's compression argument (('ADOBE_DEFLATE', 1), None, ...) -> deflate level (0 = stored), or None when
    the codec does not write that scheme."""
    if not compression:
        return 0
    name, level = (compression[0], compression[1] if len(compression) > 1 else 6) if isinstance(compression, (tuple, list)) \
        else (compression, 6)
    name = str(name).upper()
    if name in ("NONE", "1"):
        return 0
    if name in ("ADOBE_DEFLATE", "DEFLATE", "ZLIB", "8"):
        level = 6 if level is None else int(level)
        return 0 if level <= 0 else min(level, 9)
    if name in ("ZSTD", "50000"):                      # the reference's other option: compression=('ZSTD', 1)
        level = 1 if level is None else int(level)
        return 100 + max(1, min(level, 22))
    return None


def can_write(img, compression) -> bool:
    return (isinstance(img, np.ndarray) and img.ndim == 2 and img.dtype in _CODES and _level(compression) is not None
            and img.dtype.isnative)


def write_tiff(path, img, compression=("ADOBE_DEFLATE", 1), threads=None):
    img = np.ascontiguousarray(img)
    rc = lib().b2sio_write_tiff(os.fsencode(path), img.ctypes.data, img.shape[0], img.shape[1], _CODES[img.dtype],
                                _level(compression), int(threads or default_threads()))
    if rc:
        raise _err(rc)


def write_tiff_batch(paths, planes, compression=("ADOBE_DEFLATE", 1), threads=None):
    """planes: (n, H, W) C-contiguous; returns the per-file status list."""
    n = len(paths)
    if planes.ndim != 3 or planes.shape[0] < n or not planes.flags.c_contiguous:
        raise ValueError("planes must be a C-contiguous (n, H, W) array")
    arr = (C.c_char_p * n)(*[os.fsencode(p) for p in paths])
    status = (C.c_int32 * max(n, 1))()
    lib().b2sio_write_tiff_batch(arr, n, planes.ctypes.data, planes.strides[0], planes.shape[1], planes.shape[2],
                                 _CODES[planes.dtype], _level(compression), int(threads or default_threads()), status)
    return list(status[:n])


def is_deflate_level_1(compression) -> bool:
    """the reference's default output scheme, compression=('ADOBE_DEFLATE', 1) (or 'DEFLATE' / 'ZLIB' at level <= 1): the one the
    GPU encoder (b2s_deflate_strips) stands in for."""
    if not isinstance(compression, (tuple, list)) or len(compression) < 2 or compression[1] is None:
        return False
    return str(compression[0]).upper() in ("ADOBE_DEFLATE", "DEFLATE", "ZLIB", "8") and int(compression[1]) == 1


def write_tiff_strips_batch(paths, data, strip_offsets, strip_sizes, rows_per_strip, shape, dtype, compression_tag=8, threads=None):
    """TIFF files from strips compressed elsewhere: `data` uint8 array holding the streams, strip_offsets (n, strips) uint64 and
    strip_sizes (n, strips) uint32 per file; paths[i] None skips slot i.  Returns the per-file status list."""
    n = len(paths)
    strip_offsets = np.ascontiguousarray(strip_offsets, dtype=np.uint64).reshape(n, -1)
    strip_sizes = np.ascontiguousarray(strip_sizes, dtype=np.uint32).reshape(n, -1)
    arr = (C.c_char_p * max(n, 1))(*[None if p is None else os.fsencode(p) for p in paths])
    status = (C.c_int32 * max(n, 1))()
    lib().b2sio_write_tiff_strips_batch(arr, n, data.ctypes.data, strip_offsets.ctypes.data, strip_sizes.ctypes.data,
                                        strip_sizes.shape[1], int(rows_per_strip), int(shape[0]), int(shape[1]),
                                        _CODES[np.dtype(dtype)], int(compression_tag), int(threads or default_threads()), status)
    return list(status[:n])


def write_raw(path, img):
    img = np.ascontiguousarray(img, dtype=np.uint16)
    rc = lib().b2sio_write_raw(os.fsencode(path), img.ctypes.data, img.shape[0], img.shape[1])
    if rc:
        raise _err(rc)
