"""ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/README.md).

A minimal stand-in for the `pywt` module surface that /root/reference/pystripe/core.py imports
(`from pywt import wavedec2, waverec2, Wavelet, dwt_max_level`, core.py:38), backed by the C
restatement in oracle/pywt_c.c.  PARITY UNPINNED: PyWavelets itself is not available offline.

Restates (PyWavelets 1.x):
  pywt/_multilevel.py  wavedec2 / waverec2 (level rule, list layout, trim-by-one rule)
  pywt/_multidim.py    dwt2 -> dwtn (axis -2 first, then -1; keys aa/da/ad/dd), idwt2 -> idwtn (axis -1 first)
  pywt/_extensions/_dwt.pyx  dwt_axis / idwt_axis (float32 stays float32, everything else -> float64)
"""
import ctypes
import json
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None


def build(force: bool = False) -> Path:
    so = _HERE / "_build" / "liboracle.so"
    if force or not so.exists() or so.stat().st_mtime < max((_HERE / f).stat().st_mtime for f in ("pywt_c.c", "pocketfft_c.c")):
        subprocess.check_call(["make", "-s", "-C", str(_HERE)])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(str(build()))
        _LIB.orc_dwt_max_level.restype = ctypes.c_int
        _LIB.orc_dwt_max_level.argtypes = [ctypes.c_size_t, ctypes.c_size_t]
        for name in ("orc_idwt_axis_f32", "orc_idwt_axis_f64"):
            getattr(_LIB, name).restype = ctypes.c_int
    return _LIB


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data)


_TABLES = None


def _tables():
    global _TABLES
    if _TABLES is None:
        _TABLES = json.loads((_HERE / "wavelet_tables.json").read_text())
    return _TABLES


class Wavelet:
    """pywt.Wavelet subset: orthogonal families, filters derived from dec_lo (pywt wavelets.c)."""

    def __init__(self, name):
        if isinstance(name, Wavelet):
            name = name.name
        t = _tables()
        if name not in t:
            raise ValueError(f"Unknown wavelet name '{name}', check wavelist() for the list of available builtin wavelets.")
        self.name = name
        dec_lo = np.asarray(t[name], dtype=np.float64)
        F = dec_lo.size
        rec_lo = dec_lo[::-1].copy()
        rec_hi = np.array([(-1) ** k * rec_lo[F - 1 - k] for k in range(F)], dtype=np.float64)
        dec_hi = rec_hi[::-1].copy()
        self.dec_lo, self.dec_hi, self.rec_lo, self.rec_hi = dec_lo, dec_hi, rec_lo, rec_hi
        self.dec_len = self.rec_len = F


def dwt_max_level(data_len, filter_len):
    if isinstance(filter_len, Wavelet):
        filter_len = filter_len.dec_len
    elif isinstance(filter_len, str):
        filter_len = Wavelet(filter_len).dec_len
    return lib().orc_dwt_max_level(int(data_len), int(filter_len))


def _as_float(data):
    data = np.asarray(data)
    if data.dtype != np.float32:  # pywt: float32 -> float32, anything else real -> float64
        data = data.astype(np.float64)
    return np.ascontiguousarray(data)


def dwt_axis(data, wav, axis):
    """axis is -2 or -1 of a 2-D array. returns (cA, cD)."""
    data = _as_float(data)
    ny, nx = data.shape
    F = wav.dec_len
    f32 = data.dtype == np.float32
    lo = np.ascontiguousarray(wav.dec_lo.astype(data.dtype))
    hi = np.ascontiguousarray(wav.dec_hi.astype(data.dtype))
    if axis in (-1, 1):
        shape = (ny, (nx + F - 1) // 2)
        ax = 1
    else:
        shape = ((ny + F - 1) // 2, nx)
        ax = 0
    ca = np.empty(shape, data.dtype)
    cd = np.empty(shape, data.dtype)
    scratch = np.empty(3 * max(ny, nx) + F + 8, data.dtype)
    fn = lib().orc_dwt_axis_f32 if f32 else lib().orc_dwt_axis_f64
    fn(_ptr(data), ctypes.c_size_t(ny), ctypes.c_size_t(nx), ctypes.c_int(ax), _ptr(lo), _ptr(hi),
       ctypes.c_size_t(F), _ptr(ca), _ptr(cd), _ptr(scratch))
    return ca, cd


def idwt_axis(ca, cd, wav, axis):
    ca = _as_float(ca)
    cd = _as_float(cd)
    if ca.dtype != cd.dtype:  # pywt idwtn: mixed precision -> promote
        ca = ca.astype(np.float64)
        cd = cd.astype(np.float64)
    if ca.shape != cd.shape:
        raise ValueError("Coefficients arrays must have the same size.")
    ny, nx = ca.shape
    F = wav.rec_len
    f32 = ca.dtype == np.float32
    lo = np.ascontiguousarray(wav.rec_lo.astype(ca.dtype))
    hi = np.ascontiguousarray(wav.rec_hi.astype(ca.dtype))
    if axis in (-1, 1):
        shape = (ny, 2 * nx - F + 2)
        ax = 1
    else:
        shape = (2 * ny - F + 2, nx)
        ax = 0
    out = np.empty(shape, ca.dtype)
    scratch = np.empty(4 * max(ny, nx, shape[0], shape[1]) + 8, ca.dtype)
    fn = lib().orc_idwt_axis_f32 if f32 else lib().orc_idwt_axis_f64
    rc = fn(_ptr(ca), _ptr(cd), ctypes.c_size_t(ny), ctypes.c_size_t(nx), ctypes.c_int(ax), _ptr(lo), _ptr(hi),
            ctypes.c_size_t(F), _ptr(out), _ptr(scratch))
    if rc != 0:
        raise RuntimeError("C inverse wavelet transform failed")
    return out


def dwt2(data, wavelet, mode="symmetric", axes=(-2, -1)):
    assert mode == "symmetric" and tuple(axes) == (-2, -1)
    wav = wavelet if isinstance(wavelet, Wavelet) else Wavelet(wavelet)
    a, d = dwt_axis(data, wav, -2)          # dwtn: first axis of `axes`
    aa, ad = dwt_axis(a, wav, -1)
    da, dd = dwt_axis(d, wav, -1)
    return aa, (da, ad, dd)                 # cA, (cH, cV, cD)


def idwt2(coeffs, wavelet, mode="symmetric", axes=(-2, -1)):
    assert mode == "symmetric" and tuple(axes) == (-2, -1)
    wav = wavelet if isinstance(wavelet, Wavelet) else Wavelet(wavelet)
    aa, (da, ad, dd) = coeffs
    a = idwt_axis(aa, ad, wav, -1)          # idwtn: last axis first
    d = idwt_axis(da, dd, wav, -1)
    return idwt_axis(a, d, wav, -2)


def wavedec2(data, wavelet, mode="symmetric", level=None, axes=(-2, -1)):
    data = np.asarray(data)
    if data.ndim != 2:
        raise ValueError("Expected input data to have two dimensions (oracle shim is 2-D only).")
    wav = wavelet if isinstance(wavelet, Wavelet) else Wavelet(wavelet)
    max_level = min(dwt_max_level(s, wav.dec_len) for s in data.shape)
    if level is None:
        level = max_level
    elif level < 0:
        raise ValueError("Level value of %d is too low . Minimum level is 0." % level)
    coeffs = []
    a = data
    for _ in range(level):
        a, ds = dwt2(a, wav, mode, axes)
        coeffs.append(ds)
    coeffs.append(a)
    coeffs.reverse()
    return coeffs


def waverec2(coeffs, wavelet, mode="symmetric", axes=(-2, -1)):
    if not isinstance(coeffs, (list, tuple)) or len(coeffs) < 1:
        raise ValueError("Coefficient list too short (minimum 1 array required).")
    wav = wavelet if isinstance(wavelet, Wavelet) else Wavelet(wavelet)
    a, ds = coeffs[0], coeffs[1:]
    a = np.asarray(a)
    for d in ds:
        d = tuple(np.asarray(c) for c in d)
        d_shape = d[0].shape
        idxs = tuple(slice(None, -1 if a_len == d_len + 1 else None) for a_len, d_len in zip(a.shape, d_shape))
        a = idwt2((a[idxs], d), wav, mode, axes)
    return a
