"""pystripe.core — drop-in module surface of the reference's pystripe/core.py, backed by libb200stripe.so (sm_100a).

Same names, keyword arguments, defaults and error behaviour as the reference for the per-plane destripe /
enhancement path (reference lines cited per function, paths relative to the reference root).  The arithmetic runs
on the GPU through the C ABI in include/b200stripe.h; there is no CPU fallback — without the library or without
a B200 the calls raise.  Small host-side helpers that other reference scripts import from this module
(process_images.py:37-40, parallel_image_processor.py:29-30, convert.py:18-19) are provided as plain numpy.

Extensions over the reference (additive): `filter_streaks` / `process_img` also accept (Z, H, W) stacks and CUDA
torch tensors (zero-copy, current stream); `process_stack` is the batched entry point batch_filter itself uses.
"""
import os
import re
import sys
import threading
from collections import OrderedDict
from concurrent.futures import ThreadPoolExecutor
from math import ceil, exp, log, sqrt
from multiprocessing import Process, Queue
from pathlib import Path
from queue import Empty
from time import sleep, time
from typing import Callable, Iterator, List, Tuple, Union

import numpy as np
from numpy import max as np_max, mean as np_mean, median as np_median, min as np_min
from numpy import ndarray, float32, uint8, uint16, zeros

from . import _io, _native
from ._util import PrintColors, date_time_now
from ._wavelet_tables import DEC_LO
from .lightsheet_correct import correct_lightsheet, prctl  # noqa: F401  (re-exported, process_images.py:39)
from .raw import raw_imread

try:  # the reference imports these from torch.cuda (core.py:55-58); keep the names importable
    from torch.cuda import device_count as cuda_device_count
    from torch.cuda import get_device_properties as cuda_get_device_properties
    from torch.cuda import is_available as cuda_is_available_for_pt
except Exception:  # pragma: no cover
    def cuda_device_count():
        return 0

    def cuda_get_device_properties(i):
        raise RuntimeError("torch is not available")

    def cuda_is_available_for_pt():
        return False

SUPPORTED_EXTENSIONS = ('.png', '.tif', '.tiff', '.raw', '.dcimg')
NUM_RETRIES: int = 40
USE_NUMEXPR: bool = False       # kept for callers that read the flag (process_images.py:40)
USE_PYTORCH = False             # reference value on Linux (core.py:85-91); the GPU path here is not torch
USE_JAX = False
CUDA_IS_AVAILABLE_FOR_PT = False
REFERENCE_QUIRKS = False        # True: reproduce as-written behaviour (Gaussian discarded, uint16 /= flat TypeError)
MAX_BATCH = int(os.environ.get("B200STRIPE_MAX_BATCH", "8"))
EXACT = os.environ.get("B200STRIPE_EXACT", "1") != "0"


# --------------------------------------------------------------------------------------------------------------
# host helpers imported by other reference scripts
# --------------------------------------------------------------------------------------------------------------
def is_uniform_1d(arr: ndarray):
    """core.py:94-103."""
    if len(arr) <= 0:
        return None
    return bool((arr == arr[0]).all())


def is_uniform_2d(arr: ndarray):
    """core.py:106-121."""
    if len(arr) <= 0:
        return None
    return bool((arr == arr[0, 0]).all())


def is_uniform_3d(arr: ndarray):
    """core.py:124-139."""
    if len(arr) <= 0:
        return None
    return bool((arr == arr[0, 0, 0]).all())


def min_max_1d(arr: ndarray):
    """core.py:142-164."""
    if len(arr) <= 0:
        return None, None
    return arr.min(), arr.max()


def min_max_2d(arr: ndarray):
    """core.py:167-177."""
    if len(arr) <= 0:
        return None, None
    return arr.min(), arr.max()


def expm1_jit(img, dtype=float32):
    """core.py:180-187 (host helper; the hot path evaluates expm1 on the GPU)."""
    return np.expm1(img).astype(dtype)


def log1p_jit(img, dtype=float32):
    """core.py:190-197 (host helper; the hot path evaluates log1p on the GPU)."""
    return np.log1p(img, dtype=dtype)


def convert_to_16bit_fun(img: ndarray):
    """core.py:397-399."""
    np.clip(img, 0, 65535, out=img)
    return img.astype(uint16)


def convert_to_8bit_fun(img: ndarray, bit_shift_to_right: int = 8):
    """core.py:402-423 (host helper for callers; process_img does this on the GPU)."""
    if img is None or img.dtype in ('uint8', uint8):
        return img
    elif img.dtype not in ('uint16', uint16):
        img = convert_to_16bit_fun(img)
    if bit_shift_to_right is None:
        bit_shift_to_right = 8
    if 0 <= bit_shift_to_right < 9:
        lower_bound = 2 ** bit_shift_to_right
        img = np.where((0 < img) & (img < lower_bound), 1, img >> bit_shift_to_right)
    else:
        print("right shift should be between 0 and 8")
        raise RuntimeError
    np.clip(img, 0, 255, out=img)
    return img.astype(uint8)


def np_notch(length: int, sigma: float) -> ndarray:
    """core.py:637-667."""
    if length <= 0:
        raise ValueError('np_notch: length must be positive')
    if sigma <= 0:
        raise ValueError('np_notch: sigma must be positive')
    g = np.arange(length, dtype=float32)
    g **= 2
    g /= -float32(2) * sigma ** 2
    return float32(1) - np.exp(g)


def notch_rise_point(sigma, rise: float):
    """core.py:670-678."""
    return int(sqrt(-2 * sigma ** 2 * log(1 - rise)) + .5) // 2 * 2


def calculate_pad_size(shape: tuple, sigma, rise: float = 0.5):
    """core.py:681-698 (the same rule is evaluated inside the library for the plan geometry)."""
    if sigma == 0:
        return 0
    x = shape[1] + 1
    y = shape[0] + 1
    c = 5e14
    sqrt_xyc = sqrt(x ** 2 - 2 * x * y + y ** 2 + 4 * c)
    rise = min(round(1 - exp((x + y - sqrt_xyc) / (4 * sigma ** 2)), 2) - 0.01, rise)
    return notch_rise_point(sigma, rise)


def np_gaussian_filter(shape: tuple, sigma: float, axis: int) -> ndarray:
    """core.py:701-722."""
    g = np_notch(length=shape[axis], sigma=sigma)
    if axis == -2:
        g = np.reshape(g, (shape[axis], 1))
    return np.broadcast_to(g, shape)


def max_level(min_len, wavelet):
    """core.py:466-472: pywt.dwt_max_level(min_len, Wavelet(wavelet).dec_len)."""
    f = len(_dec_lo(wavelet))
    if f <= 1 or min_len < f - 1:
        return 0
    return int(min_len // (f - 1)).bit_length() - 1


def calculate_down_sampled_size(tile_size, down_sample):
    """core.py:1162-1170."""
    if isinstance(down_sample, (int, float)):
        tile_size = [ceil(size / down_sample) for size in tile_size]
    elif isinstance(down_sample, (tuple, list)):
        tile_size = list(tile_size)
        for idx, factor in enumerate(down_sample):
            if factor is not None:
                tile_size[idx] = ceil(tile_size[idx] / factor)
    return tile_size


def normalize_flat(flat):
    """core.py:2047-2049."""
    flat_float = flat.astype(float32)
    return flat_float / flat_float.max()


def _unsupported(name):
    def f(*a, **k):
        raise NotImplementedError(f"{name} is outside the B200 hot path (SURVEY.md §2: never enabled by a caller)")
    f.__name__ = name
    return f


hist_match = _unsupported("hist_match")                    # core.py:426
correct_bleaching = _unsupported("correct_bleaching")      # core.py:501
otsu_threshold = _unsupported("otsu_threshold")            # core.py:562
foreground_fraction = _unsupported("foreground_fraction")  # core.py:586


# --------------------------------------------------------------------------------------------------------------
# plan cache
# --------------------------------------------------------------------------------------------------------------
def _dec_lo(wavelet):
    if hasattr(wavelet, "dec_lo"):
        return tuple(float(v) for v in wavelet.dec_lo)
    name = str(wavelet)
    if name not in DEC_LO:
        raise ValueError(f"Unknown wavelet name '{name}', check wavelist() for the list of available builtin wavelets.")
    return DEC_LO[name]


def wavelist():
    return sorted(DEC_LO)


_PAD_ALL = ('constant', 'edge', 'linear_ramp', 'maximum', 'mean', 'median', 'minimum', 'reflect', 'symmetric',
            'wrap', 'empty')
_plans = OrderedDict()            # LRU: most recently used last
_plans_lock = threading.Lock()
_MAX_CACHED_PLANS = 16


_tls = threading.local()


class use_device:
    """`with use_device(i):` — GPU used for host (numpy) inputs on this thread."""

    def __init__(self, device: int):
        self.device = int(device)

    def __enter__(self):
        self.prev = getattr(_tls, "device", None)
        _tls.device = self.device
        return self

    def __exit__(self, *exc):
        _tls.device = self.prev


def _device_of(x) -> int:
    if _native._is_torch(x):
        if not x.is_cuda:
            raise TypeError("torch tensors must live on a CUDA device (pass numpy arrays for host data)")
        return x.device.index or 0
    dev = getattr(_tls, "device", None)
    if dev is not None:
        return dev
    return int(os.environ.get("B200STRIPE_DEVICE", os.environ.get("LOCAL_RANK", "0")))


_flat_ids = {}


def _flat_fingerprint(flat):
    """(shape, CRC32 of the values): two flat-field arrays with the same content share a plan (batch_filter normalises its
    flat argument into a new array on every call).  The checksum is computed once per array object."""
    import weakref
    import zlib
    ent = _flat_ids.get(id(flat))
    if ent is not None and ent[0]() is flat:
        return ent[1]
    a = flat.detach().cpu().numpy() if _native._is_torch(flat) else np.asarray(flat)
    fp = (tuple(a.shape), str(a.dtype), zlib.crc32(np.ascontiguousarray(a).view(np.uint8)))
    try:
        if len(_flat_ids) > 64:
            _flat_ids.clear()
        _flat_ids[id(flat)] = (weakref.ref(flat), fp)
    except TypeError:
        pass
    return fp


def _get_plan(device, shape, in_code, *, process, sigma, level, wavelet, threshold, padding_mode, bidirectional,
              log1p, flat=None, gaussian=False, down_sample=None, down_sample_method='max', dark=0, lightsheet=False,
              artifact_length=150, background_window_size=200, percentile=0.25, lightsheet_vs_background=2.0,
              convert_to_16bit=False, convert_to_8bit=False, bit_shift_to_right=8, rotate=0, flip=False,
              out_code=None, max_batch=None, stop_after=0, exact=None, new_size=None, bleach=None, pad_constant=0.0, aa=(None, None),
              mask=None, _acquire=False):
    """Cached plan for one (device, shape, dtype, parameter set).  `_acquire=True` marks the plan in use until
    `_release_plan` (eviction and the out-of-memory retry only close idle plans); a plan serialises its own runs
    with `plan.lock`, so two threads that ask for the same parameters share tables and workspace safely."""
    if not isinstance(sigma, (tuple, list)):
        sigma = (sigma,) * 2
    s1, s2 = float(sigma[0]), float(sigma[1])
    destripe = not (s1 == 0 and s2 == 0)
    mode = padding_mode.lower() if isinstance(padding_mode, str) else padding_mode
    if destripe:
        if mode not in _PAD_ALL:                                     # core.py:1088-1099
            print(f"{PrintColors.FAIL}Unsupported padding mode: {padding_mode}{PrintColors.ENDC}")
            raise RuntimeError(f"Unsupported padding mode: {padding_mode}")
        if mode not in _native.PAD_MODES:
            raise NotImplementedError(f"padding_mode='{mode}' is not implemented on the GPU path "
                                      f"(available: {sorted(_native.PAD_MODES)})")
    ds = None
    if down_sample is not None:
        if isinstance(down_sample, (int, float)):
            down_sample = (down_sample, down_sample)
        ds = tuple(1 if d is None else int(d) for d in down_sample)
        if ds == (1, 1):
            ds = None
    method = str(down_sample_method).lower()
    if ds is not None and method not in _native.DS_METHODS:          # core.py:1288-1298
        print(f"{PrintColors.FAIL}unsupported down-sampling method: {down_sample_method}{PrintColors.ENDC}")
        raise RuntimeError(f"unsupported down-sampling method: {down_sample_method}")
    taps = _dec_lo(wavelet) if destripe else None
    flat_key = None
    if flat is not None:
        flat_key = _flat_fingerprint(flat)
    key = (device, tuple(shape), in_code, process, s1, s2, int(level), taps, mode if destripe else None,
           bool(bidirectional), bool(log1p), threshold is not None and threshold <= 0, flat_key, bool(gaussian), ds,
           method, float(dark or 0), bool(lightsheet), artifact_length, background_window_size, percentile,
           lightsheet_vs_background, bool(convert_to_16bit), bool(convert_to_8bit), bit_shift_to_right, rotate,
           bool(flip), out_code, max_batch, stop_after, exact, REFERENCE_QUIRKS, new_size, bleach, tuple(None if a is None else a[0] for a in aa),
           float(pad_constant) if (destripe and mode == 'constant') else 0.0, mask)
    with _plans_lock:
        plan = _plans.get(key)
        if plan is not None:
            _plans.move_to_end(key)
            if _acquire:
                plan._users += 1
            return plan
        p = _native.default_params()
        p.height, p.width = int(shape[0]), int(shape[1])
        p.in_dtype = in_code
        p.out_dtype = in_code if out_code is None else out_code
        p.sigma1, p.sigma2 = s1, s2
        p.threshold_nonpositive = int(threshold is not None and threshold <= 0)
        p.level = int(level)
        p.pad_mode = _native.PAD_MODES.get(mode, 0) if destripe else 0
        p.bidirectional = int(bool(bidirectional))
        p.log1p = int(bool(log1p))
        p.process_img = int(process)
        p.has_flat = int(flat is not None)
        p.gaussian = int(bool(gaussian))
        if ds is not None:
            p.down_sample_y, p.down_sample_x = ds
            p.down_sample_method = _native.DS_METHODS[method]
        p.dark = float(dark or 0)
        p.lightsheet = int(bool(lightsheet))
        p.artifact_length = int(artifact_length)
        p.background_window_size = int(background_window_size)
        p.percentile = float(percentile)
        p.lightsheet_vs_background = float(lightsheet_vs_background)
        p.convert_to_16bit = int(bool(convert_to_16bit))
        p.convert_to_8bit = int(bool(convert_to_8bit))
        p.bit_shift_to_right = 8 if bit_shift_to_right is None else int(bit_shift_to_right)
        p.rotate = int(rotate or 0)
        p.flip_upside_down = int(bool(flip))
        p.reference_quirks = int(REFERENCE_QUIRKS)
        if new_size is not None:
            p.new_height, p.new_width = int(new_size[0]), int(new_size[1])
        if bleach is not None:
            (p.bleach_b0, p.bleach_b1, p.bleach_a1, p.bleach_zi,
             p.bleach_clip_min, p.bleach_clip_med, p.bleach_clip_max, p.bleach, p.bleach_per_plane) = bleach
        p.pad_constant = float(pad_constant)
        if mask is not None:
            p.mask, (p.mask_threshold, p.mask_close, p.mask_open, p.mask_per_plane) = 1, mask
        p.aa_radius_y = 0 if aa[0] is None else int(aa[0][0])
        p.aa_radius_x = 0 if aa[1] is None else int(aa[1][0])
        p.max_batch = int(max_batch or MAX_BATCH)
        p.debug_stop_after = int(stop_after)
        p.exact = int(EXACT if exact is None else exact)
        try:
            plan = _native.Plan(_native.context(device), p, dec_lo=taps, flat=flat)
        except MemoryError:
            # cached plans own workspace (gigabytes each for whole stitched slices): give back what is idle, try once more
            _evict_idle(keep=0)
            plan = _native.Plan(_native.context(device), p, dec_lo=taps, flat=flat)
        plan._flat_ref = flat  # keep id(flat) stable while the plan is cached
        plan._users = 1 if _acquire else 0
        _upload_numpy_notch_tables(plan, s1, s2)
        for ax in range(2):
            if aa[ax] is not None:
                plan.set_aa_weights(ax, aa[ax][1])
        _plans[key] = plan
        _evict_idle(keep=_MAX_CACHED_PLANS, protect=plan)   # plans own GPU workspace: keep the cache small
        return plan


def _evict_idle(keep, protect=None):
    """close least-recently-used plans nobody is running until at most `keep` remain (caller holds _plans_lock)."""
    for k in list(_plans.keys()):
        if len(_plans) <= keep:
            break
        old = _plans[k]
        if old is protect or getattr(old, "_users", 0) > 0 or not old.lock.acquire(blocking=False):
            continue                                  # another thread is inside b2s_run with it
        try:
            del _plans[k]
            old.close()
        finally:
            old.lock.release()


def _release_plan(plan):
    with _plans_lock:
        plan._users = max(0, getattr(plan, "_users", 0) - 1)


def _upload_numpy_notch_tables(plan, s1, s2):
    """np_notch (core.py:637-667) is numpy arithmetic in the reference; numpy's float32 exp is not libm's expf, so the
    tables are evaluated here, by numpy, exactly as np_filter_coefficient (core.py:749-754) does, and handed to the
    library.  sigma of a level = (length of the OTHER sub-band axis) * sigma / (padded image rows or cols)."""
    info = plan.info
    if info.n_passes == 0:
        return
    sigmas = (s1,) if info.n_passes == 1 else (s1, s2)
    f64 = plan.wants_notch_matrix

    def response(n, g):
        """the float64 path (integer pixels, no log1p): np_filter_coefficient as a matrix — row k = irfft(rfft(e_k) * g),
        evaluated by scipy.fftpack in float64 exactly as core.py:749-754 would on a float64 coefficient array."""
        from scipy.fftpack import irfft, rfft
        spec = rfft(np.eye(n, dtype=np.float64), axis=-1)
        spec *= g
        return irfft(spec, axis=-1)
    for pi, sg in enumerate(sigmas):
        for lvl in range(1, info.levels + 1):
            rows, cols = info.level_rows[lvl - 1], info.level_cols[lvl - 1]
            g0 = np_notch(cols, rows * (sg / info.padded_height))
            if f64:
                plan.set_notch_matrix(pi, lvl, 0, response(cols, g0))
            else:
                plan.set_notch(pi, lvl, 0, g0)
            if plan.params.bidirectional:
                g1 = np_notch(rows, cols * (sg / info.padded_width))
                if f64:
                    plan.set_notch_matrix(pi, lvl, 1, response(rows, g1))
                else:
                    plan.set_notch(pi, lvl, 1, g1)


def clear_plan_cache():
    """close every idle cached plan (plans another thread is running stay until it is done)."""
    with _plans_lock:
        _evict_idle(keep=0)


def pinned_empty(shape, dtype, device: int = None) -> ndarray:
    """numpy array over page-locked host memory on the GPU's NUMA node (b2s_host_alloc).  Stacks built in such an array
    travel to the GPU without the staging copy ordinary (pageable) arrays need; results of host calls are returned in
    arrays of the same kind.  The memory goes back to a recycling pool when the last view of the array is dropped."""
    dev = device if device is not None else _device_of(None)
    return _native.context(dev).pooled_empty(tuple(shape), dtype)


def _run_with_levels(plan, arr, bleach, clip_min, clip_med, clip_max, padding_mode, skip_uniform, mask=None):
    """run the plan; when bleach clip levels or the mask threshold are per plane (multi-Otsu), compute and upload them first,
    atomically with the run (the plan may be shared with another thread)."""
    per_plane_bleach = bleach is not None and bool(bleach[-1])
    per_plane_mask = mask is not None and bool(mask[-1])
    if not per_plane_bleach and not per_plane_mask:
        return _run(plan, arr)
    if _code_of(arr) == _native.F32:
        raise NotImplementedError("clip levels left to threshold_multiotsu need integer pixels (uint8 / uint16): the "
                                  "levels come from the exact intensity histogram")
    constant = isinstance(padding_mode, str) and padding_mode.lower() == 'constant'
    if constant and not per_plane_bleach and clip_min is None and plan.info.n_passes > 0:
        raise NotImplementedError("enable_masking without clip levels and without bleach correction makes the reference pad "
                                  "with log1p(multi-Otsu clip_min) per image (core.py:1066-1077, 1101-1105): not implemented "
                                  "for padding_mode='constant'")
    levels, pads, meds = _otsu_levels(arr, clip_min, clip_med, clip_max, constant, skip_uniform, check=per_plane_bleach)
    with plan.lock:
        if per_plane_bleach:
            plan.set_bleach_levels(levels, pads)
        if per_plane_mask:
            plan.set_mask_thresholds(meds)
        return _run(plan, arr)


def _run(plan, img):
    if _native._is_torch(img):
        return plan.run_torch(img)
    return plan.run_host(img)


def _as_supported(img):
    """(array for the GPU, restore-dtype or None)."""
    if _native._is_torch(img):
        import torch
        if img.dtype in (torch.uint8, torch.uint16, torch.float32):
            return img, None
        if img.dtype == torch.float64:
            return img.float(), torch.float64
        raise TypeError(f"unsupported tensor dtype {img.dtype}")
    img = np.asarray(img)
    if not img.dtype.isnative:                     # big-endian .raw tiles (raw.py:33-38): same values, native order
        img = img.astype(img.dtype.newbyteorder('='))
    if img.dtype in (np.uint8, np.uint16, np.float32):
        return img, None
    if img.dtype == np.float64:
        return img.astype(np.float32), np.float64
    raise TypeError(f"unsupported dtype {img.dtype}: the GPU path takes uint8, uint16, float32 or float64 planes")


def _code_of(img):
    if _native._is_torch(img):
        import torch
        return {torch.uint8: _native.U8, torch.uint16: _native.U16, torch.float32: _native.F32}[img.dtype]
    return _native.np_dtype_code(img.dtype)


# --------------------------------------------------------------------------------------------------------------
# filter_streaks  (core.py:982-1159)
# --------------------------------------------------------------------------------------------------------------
def _bleach_plan_args(frequency, clip_min, clip_med, clip_max, max_method, enable_masking):
    """Host side of correct_bleaching (core.py:501-559) and butter_lowpass_filter (core.py:493-499): the reference's
    argument checks, the first-order Butterworth section and its steady-state start value from scipy.signal (filter
    design, a handful of scalars), and the clip levels in the precision numpy.clip compares them in.  Returns
    (bleach tuple for the plan | None, constant-padding value)."""
    pad_constant = 0.0
    if clip_min is not None:                                            # core.py:1101-1105
        pad_constant = float(np.float32(np.log1p(clip_min)))
    if frequency is None:
        return None, pad_constant
    assert isinstance(frequency, (float, float32, np.float64)) and frequency > 0     # core.py:521
    from scipy.signal import butter, sosfilt_zi
    sos = butter(1, frequency, output='sos')                            # core.py:495: [[b0, b1, 0, 1, a1, 0]]
    zi = sosfilt_zi(sos)
    assert sos.shape == (1, 6) and sos[0, 2] == 0 and sos[0, 5] == 0 and sos[0, 3] == 1 and zi[0, 1] == 0
    section = (float(sos[0, 0]), float(sos[0, 1]), float(sos[0, 4]), float(zi[0, 0]))
    method = 2 if max_method else 1
    if clip_min is None or clip_med is None or clip_max is None:
        # core.py:1066-1077: the missing levels come from threshold_multiotsu(log1p(img)), per image -> per-plane device data
        return section + (0.0, 0.0, 0.0, method, 1), pad_constant
    return section + _clip_levels(clip_min, clip_med, clip_max) + (method, 0), pad_constant


def _clip_levels(clip_min, clip_med, clip_max):
    """correct_bleaching's argument checks (core.py:522-531) and the three levels in the precision numpy.clip compares them."""
    ok = (float, float32, np.float64)
    assert isinstance(clip_min, ok) and clip_min >= 0
    assert isinstance(clip_med, ok) and clip_med > clip_min
    assert isinstance(clip_max, ok) and clip_max > clip_min
    assert clip_max > clip_med
    clip_min_lb = np.log1p(1)                                           # numpy float64 scalar (core.py:529-531)
    if clip_min < clip_min_lb:
        clip_min = clip_min_lb

    def as_clip_sees(v):      # numpy.clip(float32 array, bound): a Python float is weak (-> float32), numpy.float64 is not
        return float(v) if isinstance(v, np.float64) else float(np.float32(v))
    return as_clip_sees(clip_min), float(np.float32(clip_med)), as_clip_sees(clip_max)


def _otsu_levels(arr, clip_min, clip_med, clip_max, constant_padding: bool, skip_uniform: bool, check: bool = True):
    """Per-plane clip levels when some are left to multi-Otsu (core.py:1066-1077): exact per-plane histograms on the GPU
    (b2s_histogram), thresholds on the host from the bins (pystripe/stack_stats.py), then the same checks as explicit levels
    (`check`: correct_bleaching's, when the levels feed it).
    Returns ([n, 3] float64 levels, [n] float32 constant-padding values or None, [n] float64 clip_med as get_img_mask sees it)."""
    from . import stack_stats
    a3 = arr if arr.ndim == 3 else arr[None]
    h = stack_stats.histogram(a3, per_plane=True)
    h = h.cpu().numpy() if _native._is_torch(h) else h
    levels = np.empty((h.shape[0], 3), np.float64)
    pads = np.zeros(h.shape[0], np.float32) if constant_padding else None
    meds = np.zeros(h.shape[0], np.float64)
    for z in range(h.shape[0]):
        if skip_uniform and np.count_nonzero(h[z]) <= 1:               # process_img returns zeros for such a plane (core.py:1232)
            levels[z] = (1.0, 2.0, 3.0)
            continue
        lb, mb, ub = stack_stats.threshold_multiotsu_from_histogram(h[z], classes=4)
        cmin = lb if clip_min is None else clip_min
        cmed = mb if clip_med is None else clip_med
        cmax = ub if clip_max is None else clip_max
        meds[z] = _mask_threshold(cmed, np.float32)
        levels[z] = _clip_levels(cmin, cmed, cmax) if check else (cmin, cmed, cmax)
        if constant_padding:
            pads[z] = np.float32(np.log1p(cmin))                       # core.py:1101-1105
    return levels, pads, meds


def _mask_threshold(threshold, image_dtype) -> float:
    """`img > threshold` (core.py:479) as numpy evaluates it: a Python scalar is weak (rounded to float32 against the float32
    log image), a numpy scalar promotes with the image dtype; against an integer image every comparison is exact.  The
    kernel compares in double, which holds all of these exactly."""
    if np.dtype(image_dtype) == np.float32:
        if isinstance(threshold, np.generic) and np.result_type(np.float32, threshold.dtype) != np.float32:
            return float(threshold)
        return float(np.float32(threshold))
    return float(threshold)


def filter_streaks(
        img,
        sigma: Tuple[int, int] = (250, 250),
        level: int = 0,
        wavelet: str = 'db9',
        crossover: float = 10,
        threshold: float = None,
        padding_mode: str = "wrap",
        bidirectional: bool = False,
        gpu_semaphore: Queue = None,
        bleach_correction_frequency: float = None,
        bleach_correction_max_method: bool = False,
        bleach_correction_clip_min: Union[float, int] = None,
        bleach_correction_clip_med: Union[float, int] = None,
        bleach_correction_clip_max: Union[float, int] = None,
        log1p_normalization_needed: bool = True,
        enable_masking: bool = False,
        close_steps: int = 50,
        open_steps: int = 500,
        verbose: bool = False
):
    """Filter horizontal streaks with the wavelet-FFT notch (reference core.py:982-1159) on the GPU.

    img: (H, W) or (Z, H, W); numpy array (host round trip) or CUDA torch tensor (zero-copy, current stream).
    `gpu_semaphore`, `crossover` are accepted and ignored (the thresholded dual-band variant is unreachable in the
    reference, core.py:1113-1117).  Bleach correction (core.py:501-559, both methods, explicit or multi-Otsu clip levels) and
    masking (get_img_mask, core.py:475-489) run on the GPU.
    """
    if not isinstance(sigma, (tuple, list)):
        sigma = (sigma,) * 2
    if sigma[0] == sigma[1] == 0 and bleach_correction_frequency is None:
        return img                                                      # core.py:1058-1059
    bleach, pad_constant = _bleach_plan_args(bleach_correction_frequency, bleach_correction_clip_min,
                                             bleach_correction_clip_med, bleach_correction_clip_max,
                                             bleach_correction_max_method, enable_masking)
    arr, restore = _as_supported(img)
    mask = None
    if enable_masking and close_steps is not None and open_steps is not None:        # core.py:1079-1080
        if bleach_correction_clip_med is None and not log1p_normalization_needed:
            raise NotImplementedError("enable_masking with the threshold left to threshold_multiotsu on an image that is not "
                                      "log-normalised (skimage's exact integer histogram path) is not implemented")
        seen = np.float32 if (log1p_normalization_needed or _code_of(arr) == _native.F32) else np.uint16
        mask = (0.0 if bleach_correction_clip_med is None else _mask_threshold(bleach_correction_clip_med, seen),
                int(close_steps), int(open_steps), int(bleach_correction_clip_med is None))
    plan = _get_plan(_device_of(arr), arr.shape[-2:], _code_of(arr), process=0, sigma=sigma, level=level,
                     wavelet=wavelet, threshold=threshold, padding_mode=padding_mode, bidirectional=bidirectional,
                     log1p=log1p_normalization_needed, bleach=bleach, pad_constant=pad_constant, mask=mask, _acquire=True)
    try:
        out = _run_with_levels(plan, arr, bleach, bleach_correction_clip_min, bleach_correction_clip_med,
                               bleach_correction_clip_max, padding_mode, skip_uniform=False, mask=mask)
    finally:
        _release_plan(plan)
    if verbose:
        print(f"de-striping applied: sigma={sigma}, level={level}, wavelet={wavelet}, crossover={crossover}, "
              f"threshold={threshold}, bidirectional={bidirectional}.")
    if restore is not None:
        out = out.to(restore) if _native._is_torch(out) else out.astype(restore)
    return out


# --------------------------------------------------------------------------------------------------------------
# process_img  (core.py:1190-1381)
# --------------------------------------------------------------------------------------------------------------
def _resize_target(shape, tile_size, down_sample, new_size):
    """core.py:1356-1359: `resize(img, new_size, preserve_range=True, anti_aliasing=tile_size < new_size)` unless the
    (down-sampled) tile_size equals new_size.  Returns (the (rows, cols) the GPU plan resizes to or None, the
    anti-aliasing Gaussian per axis as [(radius, weights) | None, ...])."""
    if new_size is None:
        return None, (None, None)
    new_size = tuple(int(v) for v in new_size)
    ts, work = tuple(int(v) for v in tile_size), tuple(shape)
    if down_sample is not None:
        ts = tuple(calculate_down_sampled_size(ts, down_sample))        # core.py:1300
        work = tuple(calculate_down_sampled_size(work, down_sample))
    if ts == new_size:
        return None, (None, None)
    if work == new_size or (ts < new_size) != (work < new_size):
        raise NotImplementedError("new_size with a tile_size that differs from the image shape is not implemented")
    for n_in, n_out in zip(work, new_size):                             # the shape scipy.ndimage.zoom derives from the factors
        if int(round(n_in * (1 / np.divide(n_in, n_out)))) != n_out:
            raise NotImplementedError(f"new_size {new_size}: skimage's zoom factor rounds to another output shape")
    aa = [None, None]
    if ts < new_size:                                                   # anti_aliasing=True (tuple comparison, core.py:1357)
        # skimage.transform.resize: sigma = max(0, (in / out - 1) / 2); scipy.ndimage.gaussian_filter skips an axis
        # with sigma <= 1e-15 and builds _gaussian_kernel1d(sigma, 0, int(4 sigma + 0.5)) — numpy arithmetic, done here
        factors = np.divide(work, new_size)
        sigmas = np.maximum(0, (factors - 1) / 2)
        for ax in range(2):
            sd = float(sigmas[ax])
            if sd > 1e-15:
                radius = int(4.0 * sd + 0.5)
                if radius > 0:                                          # radius 0: the kernel is [1.0], x * 1.0 == x
                    x = np.arange(-radius, radius + 1)
                    phi = np.exp(-0.5 / (sd * sd) * x ** 2)
                    aa[ax] = (radius, (phi / phi.sum())[::-1].copy())
    return new_size, tuple(aa)


def process_img(
        img,
        flat: ndarray = None,
        gaussian_filter_2d: bool = False,
        down_sample: Tuple[int, int] = None,
        down_sample_method: str = 'max',
        tile_size: Tuple[int, int] = None,
        new_size: Tuple[int, int] = None,
        exclude_dark_edges_set_them_to_zero: bool = False,
        sigma: Tuple[int, int] = (0, 0),
        level: int = 0,
        wavelet: str = 'coif15',
        crossover: float = 10,
        threshold: float = None,
        padding_mode: str = "wrap",
        bidirectional: bool = False,
        gpu_semaphore: Queue = None,
        bleach_correction_frequency: float = None,
        bleach_correction_clip_min: Union[float, int] = None,
        bleach_correction_clip_med: Union[float, int] = None,
        bleach_correction_clip_max: Union[float, int] = None,
        bleach_correction_max_method: bool = False,
        log1p_normalization_needed: bool = True,
        dark: float = 0,
        lightsheet: bool = False,
        artifact_length: int = 150,
        background_window_size: int = 200,
        percentile: float = 0.25,
        lightsheet_vs_background: float = 2.0,
        rotate: int = 0,
        flip_upside_down: bool = False,
        convert_to_16bit: bool = False,
        convert_to_8bit: bool = False,
        bit_shift_to_right: int = 8,
        d_type: str = None,
        verbose: bool = False,
        _max_batch: int = None,
):
    """Per-plane enhancement in the reference's order of operations (core.py:1190-1381), fused on the GPU:
    uniform-plane shortcut -> flat -> 5x5 Gaussian -> block down-sample -> filter_streaks -> dark -> lightsheet ->
    8/16-bit conversion -> flip -> rot90.   img: (H, W) or (Z, H, W), numpy or CUDA torch tensor."""
    if exclude_dark_edges_set_them_to_zero:
        raise NotImplementedError("dark-edge exclusion is outside the GPU hot path")
    bleach, pad_constant = _bleach_plan_args(bleach_correction_frequency, bleach_correction_clip_min,
                                             bleach_correction_clip_med, bleach_correction_clip_max,
                                             bleach_correction_max_method, False)
    if not isinstance(sigma, (tuple, list)):
        sigma = (sigma,) * 2
    arr, restore = _as_supported(img)
    shape = tuple(arr.shape[-2:])
    if tile_size is None:
        tile_size = shape
    resize_to, aa = _resize_target(shape, tile_size, down_sample, new_size)
    if d_type is None:
        d_type = np.float64 if restore is not None else (
            _native.CODE_TO_NP[_code_of(arr)])
    d_type = np.dtype(d_type)
    if d_type.kind in "ui":
        if d_type not in (np.uint8, np.uint16):
            raise TypeError(f"d_type {d_type} is not supported on the GPU path (uint8 / uint16 / float)")
        out_code = _native.np_dtype_code(d_type)
    else:
        out_code = _native.F32
    if flat is not None:
        if tuple(tile_size) != tuple(flat.shape):                       # core.py:1248-1254
            print(f"{PrintColors.WARNING}warning: image and flat arrays had different shapes{PrintColors.ENDC}")
            flat = None
        elif REFERENCE_QUIRKS and _code_of(arr) != _native.F32:
            raise TypeError("Cannot cast ufunc 'divide' output from dtype('float32') to dtype('uint16') "
                            "(reference core.py:1250 as written)")
    if not tuple(sigma) > (0, 0):                                       # core.py:1302
        sigma = (0, 0)
    plan = _get_plan(_device_of(arr), shape, _code_of(arr), process=1, sigma=sigma, level=level, wavelet=wavelet,
                     threshold=threshold, padding_mode=padding_mode, bidirectional=bidirectional,
                     log1p=log1p_normalization_needed, flat=flat, gaussian=gaussian_filter_2d, down_sample=down_sample,
                     down_sample_method=down_sample_method, dark=dark, lightsheet=lightsheet,
                     artifact_length=artifact_length, background_window_size=background_window_size,
                     percentile=percentile, lightsheet_vs_background=lightsheet_vs_background,
                     convert_to_16bit=convert_to_16bit, convert_to_8bit=convert_to_8bit,
                     bit_shift_to_right=bit_shift_to_right, rotate=rotate, flip=flip_upside_down, out_code=out_code,
                     max_batch=_max_batch, new_size=resize_to, bleach=bleach, pad_constant=pad_constant, aa=aa,
                     _acquire=True)
    try:
        if bleach is not None and bleach[-1] and (flat is not None or gaussian_filter_2d or down_sample is not None):
            raise NotImplementedError("bleach clip levels left to threshold_multiotsu (core.py:1066-1077) need the integer image "
                                      "that enters filter_streaks: not implemented together with flat / Gaussian / down-sampling")
        out = _run_with_levels(plan, arr, bleach, bleach_correction_clip_min, bleach_correction_clip_med,
                               bleach_correction_clip_max, padding_mode, skip_uniform=True)
    finally:
        _release_plan(plan)
    if out_code == _native.F32 and plan.info.out_dtype == _native.F32 and d_type != np.float32:
        out = out.astype(d_type) if not _native._is_torch(out) else out.double()
    return out


def process_stack(stack, **kwargs):
    """Batched process_img over a (Z, H, W) stack — the call batch_filter makes per group of decoded files."""
    return process_img(stack, **kwargs)


# --------------------------------------------------------------------------------------------------------------
# file I/O (boundary only — SURVEY.md §8f N2).  tifffile is used when present, else Pillow / OpenCV.
# --------------------------------------------------------------------------------------------------------------
def imread_tif_raw_png(path: Path, dtype: str = None, shape: Tuple[int, int] = None):
    """core.py:200-264: retries, returns None when the file cannot be decoded."""
    path = Path(path)
    extension = path.suffix.lower()
    img = None
    attempt = 0
    for attempt in range(NUM_RETRIES):
        try:
            if extension == '.raw':
                img = raw_imread(path, dtype=dtype, shape=shape)
            elif extension in ('.png', '.tif', '.tiff'):
                img = _decode_image(path)
            else:
                print(f"{PrintColors.WARNING}Unsupported file format: {extension}{PrintColors.ENDC}")
        except (OSError, TypeError, PermissionError, ValueError) as e:
            print(f"[Attempt {attempt + 1}]: file: {path.name} Read error: {type(e).__name__} - {e}")
            sleep(0.1)
            continue
        if img is not None:
            break
    if img is None:
        print(f"{PrintColors.FAIL}Failed to load image after {attempt + 1} attempts:\n{path.name}{PrintColors.ENDC}")
    return img


def _decode_image(path: Path):
    if path.suffix.lower() in ('.tif', '.tiff', '.png'):
        try:                                        # native codec (libb2sio): strips / tiles decoded on several threads
            return _io.read(path)
        except _io.CodecError as e:                 # (greyscale PNG only: colour / interlaced files go to Pillow below)
            if e.code == _io.ERR_IO:
                raise OSError(str(e))
        except OSError:
            pass                                    # library not built: fall through to the Python decoders
    if path.suffix.lower() in ('.tif', '.tiff'):
        try:
            import tifffile
            return tifffile.imread(path)
        except ImportError:
            pass
    from PIL import Image
    Image.MAX_IMAGE_PIXELS = None
    with Image.open(path) as im:
        im.load()
        return np.array(im)


def imsave_tif(path: Path, img: ndarray, compression: Union[Tuple[str, int], None] = ('ADOBE_DEFLATE', 1)) -> bool:
    """core.py:276-334: returns True only when the user interrupted (caller dies with dignity)."""
    path = Path(path)
    for attempt in range(1, NUM_RETRIES):
        try:
            tmp_path = path.with_suffix(".tif")
            _encode_tif(tmp_path, img, compression)
            os.chmod(tmp_path, 0o666 if os.name == 'nt' else 0o777)
            if tmp_path != path:
                tmp_path.rename(path)
            return False
        except KeyboardInterrupt:
            print(f"{PrintColors.WARNING}\ndying from imsave_tif{PrintColors.ENDC}")
            _encode_tif(path, img, compression)
            return True
        except (OSError, TypeError, PermissionError) as inst:
            if attempt == NUM_RETRIES - 1:
                print(f"After {NUM_RETRIES} attempts failed to save the file:\n{path}\n\n{type(inst)}\n{inst.args}\n{inst}\n")
                return False
            sleep(0.1)
    return False


_PIL_COMPRESSION = {'ADOBE_DEFLATE': 'tiff_adobe_deflate', 'DEFLATE': 'tiff_deflate', 'LZW': 'tiff_lzw',
                    'ZSTD': 'zstd', 'NONE': None}


def _encode_tif(path: Path, img: ndarray, compression):
    if _io.can_write(img, compression):
        try:
            _io.write_tiff(path, img, compression)
            return
        except _io.CodecError as e:
            if e.code == _io.ERR_IO:
                raise OSError(str(e))
    try:
        import tifffile
        tifffile.imwrite(path, data=img, compression=compression)
        return
    except ImportError:
        pass
    from PIL import Image
    method = None
    if compression:
        name = compression[0] if isinstance(compression, (tuple, list)) else compression
        method = _PIL_COMPRESSION.get(str(name).upper(), 'tiff_adobe_deflate')
    im = Image.fromarray(np.ascontiguousarray(img))
    if method:
        im.save(path, format="TIFF", compression=method)
    else:
        im.save(path, format="TIFF")


def imread_dcimg(path: Path, z: int):
    """core.py:337-355."""
    from dcimg import DCIMGFile
    with DCIMGFile(path) as arr:
        return arr[z]


def check_dcimg_shape(path: Path):
    from dcimg import DCIMGFile
    with DCIMGFile(path) as arr:
        return arr.shape


def check_dcimg_start(path: Path):
    return int(Path(path).name.split('.')[0])


def glob_re(pattern: str, path: Path) -> Iterator[Path]:
    """core.py:1603-1617: recursive case-insensitive regex search on file names."""
    regexp = re.compile(pattern, re.IGNORECASE)
    for p in os.scandir(path):
        if p.is_file() and regexp.search(p.name):
            yield Path(p.path)
        elif p.is_dir(follow_symlinks=False):
            yield from glob_re(pattern, Path(p.path))


def resize_to_tile(img: ndarray, tile_size: Tuple[int, int]) -> ndarray:
    """`resize(img, tile_size, preserve_range=True, anti_aliasing=True)` of read_filter_save (core.py:1540-1549) on the GPU
    (b2s_resize_aa): Gaussian per shrinking axis, order-1 zoom, clip to the image's own range.  Returns float32 — the
    reference holds float64 here and filter_streaks casts it to float32 on entry."""
    import ctypes as C
    import torch
    from .isotropic import anti_aliasing_kernels
    arr, _ = _as_supported(img)
    tile_size = (int(tile_size[0]), int(tile_size[1]))
    for n_in, n_out in zip(arr.shape, tile_size):   # the shape scipy.ndimage.zoom derives from skimage's factors
        if int(round(n_in * (1 / np.divide(n_in, n_out)))) != n_out:
            raise NotImplementedError(f"tile_size {tile_size}: skimage's zoom factor rounds to another output shape")
    aa = anti_aliasing_kernels(arr.shape, tile_size)
    dev = _device_of(arr)
    with torch.cuda.device(dev):
        t = torch.from_numpy(np.ascontiguousarray(arr)).cuda(dev)
        out = torch.empty(tile_size, dtype=torch.float32, device=t.device)
        ctx = _native.context(dev)
        wy = aa[0][1] if aa[0] else None
        wx = aa[1][1] if aa[1] else None
        ctx.check(_native.lib().b2s_resize_aa(
            ctx._h, C.c_void_p(t.data_ptr()), _code_of(arr), arr.shape[0], arr.shape[1], tile_size[0], tile_size[1],
            C.c_void_p(wy.ctypes.data if wy is not None else None), aa[0][0] if aa[0] else 0,
            C.c_void_p(wx.ctypes.data if wx is not None else None), aa[1][0] if aa[1] else 0,
            C.c_void_p(out.data_ptr()), 1, C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)))
        return out.cpu().numpy()


def get_img_mask(img, threshold, close_steps: int = 50, open_steps: int = 500, flood_fill_flag: int = 4):
    """core.py:475-489 on the GPU (b2s_img_mask): `img > threshold`, cv2 MORPH_CLOSE with ones(close_steps, close_steps),
    MORPH_OPEN with ones(open_steps, open_steps), then the background no corner pixel reaches through 4-connected background
    is added back.  img: (H, W) or (Z, H, W), numpy array or CUDA tensor; returns a bool array / tensor of the same shape."""
    if flood_fill_flag != 4:
        raise NotImplementedError("get_img_mask: only the 4-connected flood fill the reference uses is implemented")
    import ctypes as C
    import torch
    arr, _ = _as_supported(img)
    dev = _device_of(arr)
    with torch.cuda.device(dev):
        t = arr.contiguous() if _native._is_torch(arr) else torch.from_numpy(np.ascontiguousarray(arr)).cuda(dev)
        shape = tuple(t.shape)
        rows, cols = shape[-2:]
        n = 1 if t.dim() == 2 else int(shape[0])
        out = torch.empty(shape, dtype=torch.uint8, device=t.device)
        seen = np.float32 if _code_of(arr) == _native.F32 else np.uint16
        ctx = _native.context(dev)
        ctx.check(_native.lib().b2s_img_mask(ctx._h, C.c_void_p(t.data_ptr()), _code_of(arr), rows, cols, n,
                                             _mask_threshold(threshold, seen), int(close_steps), int(open_steps),
                                             C.c_void_p(out.data_ptr()), C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)))
        out = out.bool()
        return out if _native._is_torch(arr) else out.cpu().numpy()


# --------------------------------------------------------------------------------------------------------------
# read_filter_save  (core.py:1384-1600)
# --------------------------------------------------------------------------------------------------------------
_PROCESS_KEYS = ('flat', 'gaussian_filter_2d', 'sigma', 'level', 'wavelet', 'crossover', 'threshold', 'padding_mode',
                 'bidirectional', 'bleach_correction_frequency', 'bleach_correction_max_method',
                 'bleach_correction_clip_min', 'bleach_correction_clip_med', 'bleach_correction_clip_max', 'dark',
                 'lightsheet', 'artifact_length', 'background_window_size', 'percentile', 'lightsheet_vs_background',
                 'convert_to_16bit', 'convert_to_8bit', 'bit_shift_to_right', 'down_sample', 'down_sample_method',
                 'new_size', 'rotate', 'flip_upside_down')


def read_filter_save(
        input_file: Path = None,
        output_file: Path = None,
        z_idx: int = None,
        continue_process: bool = False,
        d_type: str = None,
        tile_size: Tuple[int, int] = None,
        print_input_file_names: bool = False,
        compression: Tuple[str, int] = ('ADOBE_DEFLATE', 1),
        flat: ndarray = None,
        gaussian_filter_2d: bool = False,
        sigma: Tuple[int, int] = (0, 0),
        level: int = 0,
        wavelet: str = 'coif15',
        crossover: float = 10,
        threshold: float = None,
        padding_mode: str = "reflect",
        bidirectional: bool = False,
        gpu_semaphore: Queue = None,
        bleach_correction_frequency: float = None,
        bleach_correction_max_method: bool = True,
        bleach_correction_clip_min: Union[float, int] = None,
        bleach_correction_clip_med: Union[float, int] = None,
        bleach_correction_clip_max: Union[float, int] = None,
        dark: float = 0,
        lightsheet: bool = False,
        artifact_length: int = 150,
        background_window_size: int = 200,
        percentile: float = 0.25,
        lightsheet_vs_background: float = 2.0,
        convert_to_16bit: bool = False,
        convert_to_8bit: bool = True,
        bit_shift_to_right: int = 8,
        down_sample: Tuple[int, int] = None,
        down_sample_method: str = 'max',
        new_size: Tuple[int, int] = None,
        rotate: int = 0,
        flip_upside_down: bool = False,
):
    """One file in, one TIFF out (reference core.py:1384-1600).  Failures are reported and swallowed like the
    reference does (:1594-1600)."""
    try:
        input_file, output_file = Path(input_file), Path(output_file)
        if continue_process and output_file.exists():
            return
        if print_input_file_names:
            print(f"\n{input_file}")
        img = _read_one(input_file, z_idx, d_type, tile_size, output_file)
        if img is None:
            return
        if d_type is None:
            d_type = img.dtype
        if tile_size is not None and img.shape != tuple(tile_size):              # core.py:1540-1549
            print(f"{PrintColors.WARNING}\nwarning: input tile had a different shape. resizing:\n"
                  f"\tinput_file: {input_file} -> \n\t\tinput shape = {img.shape}\n\t\tnew shape   = {tile_size}\n"
                  f"{PrintColors.ENDC}")
            img = resize_to_tile(img, tile_size)
        output_file.parent.mkdir(parents=True, exist_ok=True)
        kw = {k: v for k, v in locals().items() if k in _PROCESS_KEYS}
        img = process_img(np.ascontiguousarray(img), tile_size=img.shape, d_type=d_type, **kw)
        imsave_tif(output_file, img, compression=compression)
    except (OSError, IndexError, TypeError, RuntimeError) as inst:
        print(f"{PrintColors.WARNING}warning: read_filter_save function failed:\n{type(inst)}\n{inst.args}\n{inst}"
              f"\nPossible damaged input file: {input_file}{PrintColors.ENDC}")


def _read_one(input_file, z_idx, d_type, tile_size, output_file):
    """core.py:1515-1540: decode, or substitute a zero tile when shape and dtype are known."""
    if z_idx is None:
        img = imread_tif_raw_png(input_file, dtype=d_type, shape=tile_size)
    else:
        img = imread_dcimg(input_file, z_idx)
    if img is None and d_type is not None and tile_size is not None:
        print(f"{PrintColors.WARNING}\nimread function returned None. Possible damaged input file:\n\t{input_file}."
              f"\n\toutput file is set to a dummy zeros tile of shape {tile_size} and type {d_type}, instead:"
              f"\n\t{output_file}{PrintColors.ENDC}")
        img = zeros(dtype=d_type, shape=tile_size)
    elif img is None:
        print(f"{PrintColors.WARNING}\nimread function returned None. Possible damaged input file:\n\t{input_file}."
              f"\n\toutput file could be replaced with a dummy tile of zeros if shape and d_type were provided."
              f"{PrintColors.ENDC}")
    return img


# --------------------------------------------------------------------------------------------------------------
# job farm classes kept for API compatibility (process_images.py:37 imports them)
# --------------------------------------------------------------------------------------------------------------
class MultiProcessQueueRunner(Process):
    """core.py:1687-1771: a worker that drains a queue of kwargs dicts through `fun`.  batch_filter no longer uses
    it (the GPU scheduler below replaces the process farm); other reference scripts still can."""

    def __init__(self, progress_queue: Queue, args_queue: Queue, gpu_semaphore: Queue = None, gpu: int = None,
                 fun: Callable = read_filter_save, timeout: float = None, replace_timeout_with_dummy: bool = True):
        Process.__init__(self)
        self.daemon = False
        self.progress_queue = progress_queue
        self.args_queue = args_queue
        self.gpu_semaphore = gpu_semaphore
        self.gpu = gpu
        self.timeout = timeout
        self.die = False
        self.function = fun
        self.replace_timeout_with_dummy = replace_timeout_with_dummy

    def run(self):
        """core.py:1704-1771: per-item timeout that adapts upwards (max(t, 0.9 t + 0.3 elapsed)); an item that times out
        gets a zeros tile of `new_size or tile_size` as its output unless one exists already."""
        from concurrent.futures import ProcessPoolExecutor, ThreadPoolExecutor as _TPE, TimeoutError as _Timeout
        from concurrent.futures.process import BrokenProcessPool
        if self.gpu is not None:
            os.environ["B200STRIPE_DEVICE"] = str(self.gpu)
        running_next = True
        timeout = self.timeout
        pool = ProcessPoolExecutor(max_workers=1) if timeout else _TPE(max_workers=1)
        while not self.die:
            try:
                args: dict = self.args_queue.get(block=True, timeout=0.5)
            except Empty:
                break
            if self.gpu_semaphore is not None:
                args.update({"gpu_semaphore": self.gpu_semaphore})
            try:
                start_time = time()
                pool.submit(self.function, **args).result(timeout=timeout)
                if timeout is not None:
                    timeout = max(timeout, 0.9 * timeout + 0.3 * (time() - start_time))
            except (BrokenProcessPool, _Timeout, ValueError) as inst:
                if self.replace_timeout_with_dummy and "output_file" in args:
                    if _save_dummy_tile(args["output_file"], args.get("new_size"), args.get("tile_size"),
                                        args.get("convert_to_8bit"), args.get("input_file"), inst):
                        self.die = True
                else:
                    print(f"{PrintColors.WARNING}\nwarning: timeout reached for processing input file:\n\t"
                          f"{args.get('input_file')}\n\t\nexception instance: {type(inst)}{PrintColors.ENDC}")
                if isinstance(pool, ProcessPoolExecutor):
                    pool.shutdown(wait=False, cancel_futures=True)
                    pool = ProcessPoolExecutor(max_workers=1)
            except KeyboardInterrupt:
                self.die = True
            except Exception as inst:
                print(f"{PrintColors.WARNING}\nwarning: process unexpectedly failed for {args}."
                      f"\nexception instance: {type(inst)}\nexception: {inst}{PrintColors.ENDC}")
            self.progress_queue.put(running_next)
        pool.shutdown(wait=False)
        self.progress_queue.put(not running_next)


def _save_dummy_tile(output_file, new_size, tile_size, convert_to_8bit, input_file, reason) -> bool:
    """core.py:1736-1750: zeros of `new_size or tile_size`, uint8 when convert_to_8bit else uint16, unless the output
    exists.  Returns imsave_tif's "user interrupted" flag."""
    output_file = Path(output_file)
    shape = new_size if new_size else tile_size
    print(f"{PrintColors.WARNING}\nwarning: timeout reached for processing input file:\n\t{input_file}\n\t"
          f"a dummy (zeros) image is saved as output instead:\n\t{output_file}\nexception instance: {type(reason)}"
          f"{PrintColors.ENDC}")
    if shape is None:
        print(f"{PrintColors.WARNING}\tno tile_size / new_size given: the dummy tile cannot be written{PrintColors.ENDC}")
        return False
    if output_file.exists():
        return False
    output_file.parent.mkdir(parents=True, exist_ok=True)
    return imsave_tif(output_file, zeros(shape=tuple(int(v) for v in shape), dtype=uint8 if convert_to_8bit else uint16))


def progress_manager(progress_queue: Queue, workers: int, total: int, desc="PyStripe", unit=" images"):
    """core.py:1774-1803."""
    from tqdm import tqdm
    return_code = 0
    list_of_outputs = []
    print(f"{PrintColors.GREEN}{date_time_now()}: {PrintColors.ENDC}"
          f"using {workers} workers. {total} images need to be processed.", flush=True)
    progress_bar = tqdm(total=total, ascii=True, smoothing=0.01, mininterval=1.0, unit=unit, desc=desc)
    while workers > 0:
        try:
            still_running = progress_queue.get(block=False)
            if isinstance(still_running, bool) and still_running:
                progress_bar.update(1)
            else:
                workers -= 1
                if not isinstance(still_running, bool):
                    list_of_outputs += [still_running]
        except Empty:
            try:
                sleep(0.05)
            except KeyboardInterrupt:
                print(f"\n{PrintColors.WARNING}Terminating processes with dignity!{PrintColors.ENDC}")
                return_code = 1
        except KeyboardInterrupt:
            print(f"\n{PrintColors.WARNING}Terminating processes with dignity!{PrintColors.ENDC}")
            return_code = 1
    progress_bar.close()
    return list_of_outputs if list_of_outputs else return_code


# --------------------------------------------------------------------------------------------------------------
# batch_filter  (core.py:1806-2044) — re-designed scheduling: decode threads -> per-GPU batched plans -> encode threads
# --------------------------------------------------------------------------------------------------------------
def z_shard(n_items: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous Z partition: rank r of `world_size` owns items [lo, hi); the first
    n_items % world_size ranks hold one extra item.  Planes are independent, so this is the only multi-GPU exchange:
    no collective on the data path."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of size {world_size}")
    base, extra = divmod(max(int(n_items), 0), world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _process_group() -> Tuple[int, int, int]:
    """(world_size, rank, local_rank) when launched one process per GPU (torchrun / torch.distributed env), else (1, 0, 0)."""
    try:
        world = int(os.environ.get("WORLD_SIZE", "1"))
        rank = int(os.environ.get("RANK", "0"))
        local = int(os.environ.get("LOCAL_RANK", str(rank)))
    except ValueError:
        return 1, 0, 0
    return (world, rank, local) if world > 1 else (1, 0, 0)


def _visible_gpus() -> List[int]:
    env = os.environ.get("B200STRIPE_DEVICES")
    if env:
        return [int(v) for v in env.split(",") if v.strip() != ""]
    try:
        n = cuda_device_count()
    except Exception:
        n = 0
    return list(range(max(n, 1)))


def batch_filter(
        input_path: Path,
        output_path: Path,
        files_list: List[Path] = None,
        workers: int = None,
        threads_per_gpu: int = 8,
        flat: ndarray = None,
        gaussian_filter_2d: bool = False,
        sigma: Tuple[int, int] = (0, 0),
        level=0,
        wavelet: str = 'db9',
        crossover: int = 10,
        threshold: int = None,
        padding_mode: str = "reflect",
        bidirectional: bool = False,
        bleach_correction_frequency: float = None,
        bleach_correction_max_method: bool = True,
        bleach_correction_clip_min: Union[float, int] = None,
        bleach_correction_clip_med: Union[float, int] = None,
        bleach_correction_clip_max: Union[float, int] = None,
        dark: int = 0,
        z_step: float = None,
        rotate: int = 0,
        flip_upside_down: bool = False,
        lightsheet: bool = False,
        artifact_length: int = 150,
        background_window_size: int = 200,
        percentile: float = .25,
        lightsheet_vs_background: float = 2.0,
        convert_to_16bit: bool = False,
        convert_to_8bit: bool = False,
        bit_shift_to_right: int = 8,
        continue_process: bool = False,
        d_type: str = None,
        tile_size: Tuple[int, int] = None,
        down_sample: Tuple[int, int] = None,
        new_size: Tuple[int, int] = None,
        print_input_file_names: bool = False,
        timeout: float = None,
        compression: Tuple[str, int] = ('ADOBE_DEFLATE', 1)
) -> int:
    """Apply process_img to every image under `input_path`, writing TIFFs under `output_path`
    (reference core.py:1806-2044; same arguments, return code 0 = done, 1 = interrupted).

    Scheduling is re-designed for one box of B200s: `workers` host threads decode files into batches, one feeder
    thread per visible GPU runs a batched plan (Z planes are independent: no collective), encoder threads write the
    results.  `threads_per_gpu` is the number of planes per GPU batch.  `timeout` bounds the decode of one file (it
    adapts upwards like the reference's, core.py:1722-1724); a file that times out gets a zeros tile of `new_size or
    tile_size` as its output (core.py:1736-1750).  Return code: 0 = done, 1 = interrupted, 2 = done but some images
    could not be processed (they are listed).
    """
    from tqdm import tqdm
    input_path = Path(input_path)
    assert input_path.is_dir()
    if convert_to_16bit is True and convert_to_8bit is True:
        print(f"{PrintColors.FAIL}Select 8-bit or 16-bit output format.{PrintColors.ENDC}")
        raise TypeError("Select 8-bit or 16-bit output format.")
    output_path = Path(output_path)
    output_path.mkdir(parents=True, exist_ok=True)
    if sigma is None:
        sigma = (0, 0)
    if workers is None or workers <= 0:
        workers = os.cpu_count() or 1
    if isinstance(flat, (np.ndarray, np.generic)):
        flat = normalize_flat(flat)
    elif isinstance(flat, (Path, str)):
        flat = normalize_flat(imread_tif_raw_png(Path(flat)))
    elif flat is not None:
        print(f"{PrintColors.FAIL}flat argument should be a numpy array or a path to a flat.tif file{PrintColors.ENDC}")
        raise TypeError("flat argument should be a numpy array or a path to a flat.tif file")
    if isinstance(down_sample, tuple) and down_sample == (1, 1):
        down_sample = None
    kw = dict(flat=flat, gaussian_filter_2d=gaussian_filter_2d, sigma=sigma, level=level, wavelet=wavelet,
              crossover=crossover, threshold=threshold, padding_mode=padding_mode, bidirectional=bidirectional,
              bleach_correction_frequency=bleach_correction_frequency,
              bleach_correction_max_method=bleach_correction_max_method,
              bleach_correction_clip_min=bleach_correction_clip_min,
              bleach_correction_clip_med=bleach_correction_clip_med,
              bleach_correction_clip_max=bleach_correction_clip_max, dark=dark, lightsheet=lightsheet,
              artifact_length=artifact_length, background_window_size=background_window_size, percentile=percentile,
              lightsheet_vs_background=lightsheet_vs_background, rotate=rotate, flip_upside_down=flip_upside_down,
              convert_to_16bit=convert_to_16bit, convert_to_8bit=convert_to_8bit,
              bit_shift_to_right=bit_shift_to_right, down_sample=down_sample, new_size=new_size)

    print(f"{PrintColors.GREEN}{date_time_now()}: {PrintColors.ENDC}Scheduling jobs for images in \n\t{input_path}")
    jobs = []  # (input_file, output_file, z_idx)
    if z_step is None:
        files = glob_re(r"\.(?:tiff?|raw|png)$", input_path) if files_list is None else files_list
        for f in files:
            f = Path(f)
            out = output_path / f.relative_to(input_path)
            out = out.parent / (out.name[0:-len(out.suffix)] + '.tif')
            if continue_process and out.exists():
                continue
            jobs.append((f, out, None))
    else:
        files = glob_re(r"\.(?:dcimg)$", input_path) if files_list is None else files_list
        for f in files:
            f = Path(f)
            n = check_dcimg_shape(f)[0]
            start = check_dcimg_start(f)
            for i in range(n):
                out = output_path / f.relative_to(input_path).parent / f'z{start + i * z_step:08.1f}.tif'
                if continue_process and out.exists():
                    continue
                jobs.append((f, out, i))
    # one process per GPU (torchrun): this process takes its contiguous share of the job list on its own device
    world, rank, local_rank = _process_group()
    if world > 1:
        jobs.sort(key=lambda j: (str(j[0]), -1 if j[2] is None else j[2]))   # every rank must see the same order
        lo, hi = z_shard(len(jobs), world, rank)
        jobs = jobs[lo:hi]
    num_images = len(jobs)
    if num_images == 0:
        return 0

    gpus = [local_rank] if world > 1 and not os.environ.get("B200STRIPE_DEVICES") else _visible_gpus()
    gpus = gpus[:max(1, num_images)]
    batch = max(1, int(threads_per_gpu))
    # one pipeline item = `group` files: several plan batches per process_img call, so that the call's own slots overlap
    # H2D / kernels / D2H (a single 8-plane call runs them back to back: 10 vs 19 Gpx/s measured at 2048^2)
    per_gpu = -(-num_images // len(gpus))
    group = int(os.environ.get("B200STRIPE_FILE_GROUP", "0")) or max(batch, min(8 * batch, 64, -(-per_gpu // 6)))
    # writing costs about twice what reading does per thread (new pages of the output files): the encoder is the bottleneck stage
    per_pipe = max(2, workers // len(gpus))
    read_threads = max(1, per_pipe // 3)
    write_threads = per_pipe          # mild over-subscription: the reader blocks on its bounded queue most of the time
    gpu_deflate = _io.is_deflate_level_1(compression) and os.environ.get("B200STRIPE_GPU_DEFLATE", "1") != "0"
    if gpu_deflate:                   # the files receive streams compressed on the GPU: decoding is the host's main job
        read_threads, write_threads = per_pipe, max(2, per_pipe // 2)
    print(f"{PrintColors.GREEN}{date_time_now()}: {PrintColors.ENDC}"
          f"using {workers} decode/encode threads and {len(gpus)} GPU(s). {num_images} images need to be processed.",
          flush=True)
    progress = tqdm(total=num_images, ascii=True, smoothing=0.01, mininterval=1.0, unit=" images", desc="PyStripe")
    # argument errors the reference raises per file are the same for every file: find them before any thread starts
    _bleach_plan_args(bleach_correction_frequency, bleach_correction_clip_min, bleach_correction_clip_med,
                      bleach_correction_clip_max, bleach_correction_max_method, False)
    shared = _BatchShared(jobs=jobs, batch=batch, group=group, progress=progress, kw=kw, d_type=d_type, tile_size=tile_size,
                          compression=compression, timeout=timeout, io_threads=read_threads, write_threads=write_threads,
                          print_input_file_names=print_input_file_names,
                          gpu_deflate=gpu_deflate)
    return_code = 0
    pipelines = [_BatchPipeline(g, shared) for g in gpus]
    try:
        for pl in pipelines:
            pl.start()
        for pl in pipelines:
            pl.join()
    except KeyboardInterrupt:
        print(f"\n{PrintColors.WARNING}Terminating processes with dignity!{PrintColors.ENDC}")
        shared.stop.set()
        for pl in pipelines:
            pl.join(timeout=5.0)
        return_code = 1
    progress.close()
    if shared.stop.is_set():
        return_code = 1
    if shared.failed:
        print(f"{PrintColors.FAIL}batch_filter: {len(shared.failed)} of {num_images} images were NOT processed:{PrintColors.ENDC}")
        for f, why in shared.failed[:20]:
            print(f"\t{f}: {why}")
        return_code = return_code or 2
    return return_code


def _batch_buffer(device, shape, dtype):
    """page-locked, recycled buffer a batch of tiles is decoded into (the plan reads it over PCIe without staging)."""
    return _native.context(device).pooled_empty(shape, dtype)


class _DeflatedPlanes:
    """result planes of one batch as the GPU left them: the zlib streams of their TIFF strips in one page-locked byte array
    (b2s_deflate_strips), with the offset and size of every stream."""

    def __init__(self, data, offsets, sizes, rows_per_strip, shape, dtype):
        self.data, self.offsets, self.sizes, self.rows_per_strip, self.shape, self.dtype = data, offsets, sizes, rows_per_strip, shape, dtype

    def inflate(self, i) -> ndarray:
        """plane i decoded on the host (the writer's fallback when a file cannot be laid out natively)."""
        import zlib
        raw = b"".join(zlib.decompress(self.data[int(o):int(o) + int(n)].tobytes()) for o, n in zip(self.offsets[i], self.sizes[i]))
        return np.frombuffer(raw, dtype=self.dtype).reshape(self.shape)


# the device-resident leg keeps a group's input, result and deflate buffers in HBM (about 3.2 x the group per compute thread):
# groups of whole stitched slices stay on the host-staged path
_DEVICE_GROUP_LIMIT = int(float(os.environ.get("B200STRIPE_DEVICE_GROUP_GB", "4")) * 2 ** 30)


def _to_device(buf, device):
    """the decoded (page-locked) batch as a CUDA tensor (hook: the host-logic tests run without a GPU)."""
    import torch
    return torch.from_numpy(buf).cuda(device, non_blocking=True)


_compute_streams = {}


def _own_stream(device, index=0):
    """context: this compute thread's own CUDA stream on `device` (hook, like _to_device).  The stream objects live as long
    as the process: torch's caching allocator keys its free blocks by stream, so a fresh stream per call would re-allocate
    every batch-sized tensor."""
    import contextlib
    import torch
    with _plans_lock:
        st = _compute_streams.get((device, index))
        if st is None:
            st = _compute_streams[(device, index)] = torch.cuda.Stream(device)
    stack = contextlib.ExitStack()
    stack.enter_context(torch.cuda.device(device))
    stack.enter_context(torch.cuda.stream(st))
    return stack


def _to_host(res):
    return res.cpu().numpy() if _native._is_torch(res) else np.asarray(res)


def _deflatable(res) -> bool:
    if not _native._is_torch(res):
        return False
    import torch
    return res.is_cuda and res.dim() == 3 and res.dtype in (torch.uint8, torch.uint16)


def deflate_rows_per_strip(shape, itemsize) -> int:
    """strips of whole rows, about 64 KB of samples each (one CTA compresses one strip)."""
    return int(max(1, min(shape[0], 65536 // max(1, shape[1] * itemsize))))


def gpu_deflate(planes) -> _DeflatedPlanes:
    """the deflate step of imsave_tif (core.py:275-334, compression=('ADOBE_DEFLATE', 1)) on the GPU: planes is a CUDA tensor
    (n, H, W) of uint8 / uint16; every strip becomes one zlib stream, and only the compressed bytes cross PCIe."""
    import ctypes as C
    import torch
    if planes.dim() == 2:
        planes = planes[None]
    planes = planes.contiguous()
    n, rows, cols = (int(v) for v in planes.shape)
    code = _code_of(planes)
    np_dtype = {_native.U8: np.uint8, _native.U16: np.uint16, _native.F32: np.float32}[code]
    rps = deflate_rows_per_strip((rows, cols), np.dtype(np_dtype).itemsize)
    spp = -(-rows // rps)
    dev = planes.device.index
    ctx = _native.context(dev)
    with torch.cuda.device(dev):
        cap = int(_native.lib().b2s_deflate_bound(code, rows, cols, n, rps))
        d_out = torch.empty(cap, dtype=torch.uint8, device=planes.device)
        d_sizes = torch.empty(n * spp, dtype=torch.int32, device=planes.device)
        d_offs = torch.empty(n * spp + 1, dtype=torch.int64, device=planes.device)
        total = C.c_int64(0)
        ctx.check(_native.lib().b2s_deflate_strips(ctx._h, C.c_void_p(planes.data_ptr()), code, rows, cols, n, rps,
                                                   C.c_void_p(d_out.data_ptr()), cap, C.c_void_p(d_sizes.data_ptr()),
                                                   C.c_void_p(d_offs.data_ptr()), C.byref(total),
                                                   C.c_void_p(torch.cuda.current_stream(planes.device).cuda_stream)))
        # page-locked block in one of eight size classes below the bound: batches of a run land in the same class or two, so
        # the blocks are recycled instead of re-allocated (cudaMallocHost costs ~0.4 s per GB)
        step = max(1 << 21, -(-cap // 8))
        host = ctx.pooled_empty((-(-int(total.value) // step) * step,), np.uint8)[:int(total.value)]
        torch.from_numpy(host).copy_(d_out[:int(total.value)])
        sizes = d_sizes.cpu().numpy().view(np.uint32).reshape(n, spp)
        offs = d_offs[:-1].cpu().numpy().view(np.uint64).reshape(n, spp)
    return _DeflatedPlanes(host, offs, sizes, rps, (rows, cols), np.dtype(np_dtype))


class _BatchShared:
    """state shared by the per-GPU pipelines of one batch_filter call: the job cursor (dynamic Z partition), the progress
    bar, the failure list and the adaptive per-file timeout (core.py:1722-1724)."""

    def __init__(self, **kw):
        self.__dict__.update(kw)
        self.stop = threading.Event()
        self.lock = threading.Lock()
        self.cursor = 0
        self.failed = []

    def claim(self):
        with self.lock:
            lo = self.cursor
            hi = min(lo + self.group, len(self.jobs))
            self.cursor = hi
        return self.jobs[lo:hi]

    def fail(self, job, why):
        with self.lock:
            self.failed.append((job[0], why))
        print(f"{PrintColors.WARNING}warning: processing failed for {job[0]}: {why}{PrintColors.ENDC}")

    def done(self, n=1):
        with self.lock:
            self.progress.update(n)

    def timed(self, elapsed):
        if self.timeout is not None:
            with self.lock:
                self.timeout = max(self.timeout, 0.9 * self.timeout + 0.3 * elapsed)


class _BatchPipeline:
    """One GPU's share of batch_filter: reader -> GPU -> writer, three threads joined by bounded queues, so the decode of
    batch k+1 and the encode of batch k-1 overlap the GPU work of batch k.  A batch is decoded by the native codec
    (pystripe/_io.py, `io_threads` workers) straight into ONE page-locked buffer, which the plan reads over PCIe without a
    staging copy; the result comes back in a page-locked buffer the encoder reads.  Both buffers are recycled through the
    context's pinned pool.  Files the batch path cannot take (another shape or dtype, a format the codec does not cover,
    DCIMG) go one by one through read_filter_save's logic on the same device."""

    def __init__(self, device, shared):
        from queue import Queue as _Q
        self.device, self.sh = device, shared
        self.q_ready, self.q_write = _Q(maxsize=2), _Q(maxsize=2)
        # the device-resident leg (GPU deflate) runs upload -> kernels -> deflate -> download back to back per group: two
        # compute threads on their own CUDA streams overlap one group's copies with the other's kernels
        self.n_compute = max(1, int(os.environ.get("B200STRIPE_COMPUTE_THREADS", "2"))) if getattr(shared, "gpu_deflate", False) else 1
        self._compute_left = self.n_compute
        import functools
        computes = [functools.partial(self._compute, k) for k in range(self.n_compute)]
        for c in computes:
            c.__name__ = "_compute"
        self.threads = [threading.Thread(target=self._guard, args=(f,), daemon=True)
                        for f in [self._reader] + computes + [self._writer]]
        self.pool = ThreadPoolExecutor(max_workers=max(2, shared.write_threads))

    def start(self):
        for t in self.threads:
            t.start()

    def join(self, timeout=None):
        for t in self.threads:
            while t.is_alive():
                t.join(timeout=0.5 if timeout is None else timeout)
                if timeout is not None:
                    break
        self.pool.shutdown(wait=False)

    def _put(self, q, item) -> bool:
        """bounded put that gives up when the pipeline is stopping (a dead consumer must not block its producer for ever)."""
        from queue import Full
        while True:
            try:
                q.put(item, timeout=0.2)
                return True
            except Full:
                if self.sh.stop.is_set():
                    return False

    def _guard(self, stage):
        try:
            stage()
        except BaseException as inst:       # a dying stage must not leave the others blocked on a queue
            self.sh.stop.set()
            print(f"{PrintColors.FAIL}batch_filter: {stage.__name__} on GPU {self.device} died: {type(inst)} {inst}{PrintColors.ENDC}")
            with self.sh.lock:
                self.sh.failed.append((f"<{stage.__name__} of GPU {self.device}>", repr(inst)))
            for q in (self.q_ready, self.q_write):
                try:
                    q.put_nowait(None)
                except Exception:
                    pass

    # ---- stage 1: decode
    def _expected(self, group):
        """(shape, dtype) the batch buffer of this group has: the caller's tile_size / d_type, else the first file's."""
        sh = self.sh
        if sh.tile_size is not None and sh.d_type is not None and np.dtype(sh.d_type).kind in "uf":
            return tuple(int(v) for v in sh.tile_size), np.dtype(sh.d_type)
        for f, _, z in group:
            if z is None and f.suffix.lower() in ('.tif', '.tiff', '.raw', '.png'):
                try:
                    shape, dtype, _ = _io.probe(f)
                    return shape, dtype
                except OSError:
                    continue
        return None, None

    def _reader(self):
        sh = self.sh
        while not sh.stop.is_set():
            group = sh.claim()
            if not group:
                break
            shape, dtype = self._expected(group)
            fast, slow = [], []
            for job in group:
                f, _, z = job
                ok = shape is not None and z is None and f.suffix.lower() in ('.tif', '.tiff', '.raw', '.png') and \
                    np.dtype(dtype) in (np.uint8, np.uint16, np.float32)
                (fast if ok else slow).append(job)
            if fast:
                buf = _batch_buffer(self.device, (len(fast),) + tuple(shape), dtype)
                if sh.print_input_file_names:
                    for f, _, _ in fast:
                        print(f"\n{f}")
                valid = self._decode_fast(fast, buf, slow)
                if any(valid):
                    if not self._put(self.q_ready, (fast, buf, valid)):
                        return
                else:
                    del buf
            for job in slow:                                            # odd files: the reference's per-file logic
                if not self._put(self.q_ready, (job,)):
                    return
        self._put(self.q_ready, None)

    def _decode_fast(self, jobs, buf, slow):
        """decode jobs[i] into buf[i]; returns the per-plane valid flags.  Files of another shape move to `slow`."""
        sh = self.sh
        n = len(jobs)
        valid = [False] * n
        if sh.timeout is None:
            status = _io.read_batch([j[0] for j in jobs], buf, threads=sh.io_threads)
        else:                                                            # per-file timeout (core.py:1719-1724)
            from concurrent.futures import TimeoutError as _Timeout

            def one(i):
                try:
                    _io.read(jobs[i][0], out=buf[i], threads=1)
                    return _io.OK
                except _io.CodecError as e:
                    return e.code
            t0 = time()
            futs = [self.pool.submit(one, i) for i in range(n)]
            status = []
            for i, fu in enumerate(futs):
                try:
                    status.append(fu.result(timeout=max(0.0, sh.timeout - (time() - t0)) if i else sh.timeout))
                except _Timeout as inst:
                    status.append(None)
                    if _save_dummy_tile(jobs[i][1], sh.kw.get("new_size"), sh.tile_size, sh.kw.get("convert_to_8bit"),
                                        jobs[i][0], inst):
                        sh.stop.set()
                    sh.done()
            sh.timed((time() - t0) / max(1, n))
        for i, st in enumerate(status):
            if st == _io.OK:
                valid[i] = True
            elif st is None:
                pass                                                     # timed out: dummy tile written
            elif st == _io.ERR_SHAPE:
                slow.append(jobs[i])                                     # another shape / dtype: per-file path
            else:                                                        # codec does not cover it: Python decoders + retries
                img = _read_one(jobs[i][0], None, sh.d_type, sh.tile_size, jobs[i][1])
                if img is None:
                    sh.done()
                elif img.shape == buf.shape[1:] and img.dtype == buf.dtype:
                    buf[i] = img
                    valid[i] = True
                else:
                    slow.append(jobs[i])
        return valid

    # ---- stage 2: the GPU
    def _compute(self, index=0):
        if self.n_compute > 1:
            with _own_stream(self.device, index):
                return self._compute_loop()
        return self._compute_loop()

    def _get(self, q):
        """blocking get that gives up when the pipeline is stopping (the sentinel of a dead producer may not have fitted the
        bounded queue): returns None then, like the sentinel."""
        while True:
            try:
                return q.get(timeout=0.2)
            except Empty:
                if self.sh.stop.is_set():
                    return None

    def _compute_loop(self):
        sh = self.sh
        try:
            self._compute_items()
        finally:                                                         # also when this thread dies: the last one out tells the writer
            with sh.lock:
                self._compute_left -= 1
                last = self._compute_left == 0
            if last:
                self._put(self.q_write, None)

    def _compute_items(self):
        sh = self.sh
        while True:
            item = self._get(self.q_ready)
            if item is None or sh.stop.is_set():
                if item is None:
                    self._put(self.q_ready, None)                        # the sibling compute threads stop too
                break
            if len(item) == 1:
                self._slow_file(item[0])
                continue
            jobs, buf, valid = item
            try:
                with use_device(self.device):
                    if sh.gpu_deflate and buf.dtype in (np.uint8, np.uint16) and buf.nbytes <= _DEVICE_GROUP_LIMIT:
                        # device-resident batch: the result is deflated where it was computed, D2H carries compressed bytes
                        res = process_img(_to_device(buf, self.device), tile_size=buf.shape[1:],
                                          d_type=sh.d_type if sh.d_type is not None else buf.dtype, _max_batch=sh.batch, **sh.kw)
                        res = gpu_deflate(res) if _deflatable(res) else _to_host(res)
                    else:
                        res = process_img(buf, tile_size=buf.shape[1:], d_type=sh.d_type if sh.d_type is not None else buf.dtype,
                                          _max_batch=sh.batch, **sh.kw)
            except Exception as inst:                                    # every per-batch failure is reported, none is fatal
                for job, ok in zip(jobs, valid):
                    if ok:
                        sh.fail(job, f"{type(inst).__name__}: {inst}")
                        sh.done()
                continue
            del buf
            if not self._put(self.q_write, (jobs, res, valid)):
                return

    def _slow_file(self, job):
        sh = self.sh
        f, out, z = job
        try:
            with use_device(self.device):
                kw = dict(sh.kw)
                read_filter_save(input_file=f, output_file=out, z_idx=z, d_type=sh.d_type, tile_size=sh.tile_size,
                                 print_input_file_names=sh.print_input_file_names, compression=sh.compression, **kw)
            if not out.exists():
                sh.fail(job, "no output was written (see the warning above)")
        except Exception as inst:
            sh.fail(job, f"{type(inst).__name__}: {inst}")
        sh.done()

    # ---- stage 3: encode
    def _writer(self):
        sh = self.sh
        while True:
            item = self._get(self.q_write)
            if item is None:
                break
            jobs, res, valid = item
            if not isinstance(res, _DeflatedPlanes):
                res = np.asarray(res)
            try:
                for d in {j[1].parent for j in jobs}:
                    d.mkdir(parents=True, exist_ok=True)
                todo = [i for i, ok in enumerate(valid) if ok]
                if todo and isinstance(res, _DeflatedPlanes):
                    status = _io.write_tiff_strips_batch([jobs[i][1] if valid[i] else None for i in range(len(jobs))], res.data,
                                                         res.offsets[:len(jobs)], res.sizes[:len(jobs)], res.rows_per_strip,
                                                         res.shape, res.dtype, 8, threads=sh.write_threads)
                    for i in todo:
                        if status[i] and imsave_tif(jobs[i][1], res.inflate(i), compression=sh.compression):
                            sh.stop.set()
                elif todo and res.ndim == 3 and _io.can_write(res[0], sh.compression) and res.flags.c_contiguous:
                    if len(todo) == len(jobs):
                        status = _io.write_tiff_batch([j[1] for j in jobs], res, sh.compression, threads=sh.write_threads)
                    else:
                        futs = [self.pool.submit(_io.write_tiff_batch, [jobs[i][1]], res[i:i + 1], sh.compression, 1) for i in todo]
                        status = [fu.result()[0] for fu in futs]
                    for i, st in zip(todo, status):
                        if st:                                           # retries etc.: the reference's own writer
                            if imsave_tif(jobs[i][1], res[i], compression=sh.compression):
                                sh.stop.set()
                else:
                    futs = [self.pool.submit(imsave_tif, jobs[i][1], res[i], sh.compression) for i in todo]
                    if any(fu.result() for fu in futs):
                        sh.stop.set()
                for i in todo:
                    if not jobs[i][1].exists():
                        sh.fail(jobs[i], "the output file could not be written")
            except Exception as inst:
                for job, ok in zip(jobs, valid):
                    if ok:
                        sh.fail(job, f"write failed: {type(inst).__name__}: {inst}")
            sh.done(sum(1 for ok in valid if ok))
            del res
