"""ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product (image-preprocessing-pipeline_b200/).

CPU restatement (numpy) of the reference's per-plane destripe/enhancement path, each function citing the
reference lines it follows (paths relative to /root/reference).  Third-party arithmetic:
  * PyWavelets  -> oracle/pywt_c.c + oracle/pywt_shim.py        (restated; PARITY UNPINNED, absent offline)
  * scipy.fftpack.rfft/irfft (pocketfft), scipy.ndimage.zoom, cv2.GaussianBlur -> the real libraries (present)
  * log1p / expm1 -> glibc log1pf / expm1f (what numexpr, the reference's default, calls; oracle/pywt_c.c)
Pinned by tests/test_oracle_vs_reference.py, which executes the reference source verbatim (oracle/ref_runner.py)
in the build container, and by the golden vectors under tests/golden/ made from that same verbatim run.

Semantics switch `quirks`:
  quirks=False (default, "intended"): flat-field division is float32 (`img.astype(float32) / flat`) and the 5x5
      Gaussian is applied.  quirks=True ("as written"): uint16 `img /= flat` raises TypeError (core.py:1250) and the
      GaussianBlur result is discarded (core.py:1284), exactly as the shipped code behaves.
"""
import ctypes
from math import ceil, exp, log, sqrt

import numpy as np
from scipy.fftpack import irfft, rfft
from scipy.ndimage import zoom as ndi_zoom

from . import pywt_shim as pywt

PAD_MODES = ('constant', 'edge', 'linear_ramp', 'maximum', 'mean', 'median', 'minimum', 'reflect',
             'symmetric', 'wrap', 'empty')


# ---- libm pins ---------------------------------------------------------------------------------
def log1p_f32(x: np.ndarray) -> np.ndarray:
    """core.py:190-197 log1p_jit (numexpr 'log1p(img)' on float32 -> libm log1pf)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty_like(x)
    pywt.lib().orc_log1pf(ctypes.c_void_p(x.ctypes.data), ctypes.c_void_p(out.ctypes.data), ctypes.c_size_t(x.size))
    return out


def expm1_f32(x: np.ndarray) -> np.ndarray:
    """core.py:180-187 expm1_jit."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty_like(x)
    pywt.lib().orc_expm1f(ctypes.c_void_p(x.ctypes.data), ctypes.c_void_p(out.ctypes.data), ctypes.c_size_t(x.size))
    return out


# ---- scalars -----------------------------------------------------------------------------------
def notch_rise_point(sigma, rise):
    """core.py:670-678."""
    return int(sqrt(-2 * sigma ** 2 * log(1 - rise)) + .5) // 2 * 2


def calculate_pad_size(shape, sigma, rise=0.5):
    """core.py:681-698 (the VRAM cap term uses c = 5e14)."""
    if sigma == 0:
        return 0
    x = shape[1] + 1
    y = shape[0] + 1
    c = 5e14
    root = sqrt(x ** 2 - 2 * x * y + y ** 2 + 4 * c)
    rise = min(round(1 - exp((x + y - root) / (4 * sigma ** 2)), 2) - 0.01, rise)
    return notch_rise_point(sigma, rise)


def np_notch(length: int, sigma: float) -> np.ndarray:
    """core.py:637-667, numpy branch (:663-666): float32 throughout."""
    if length <= 0:
        raise ValueError('np_notch: length must be positive')
    if sigma <= 0:
        raise ValueError('np_notch: sigma must be positive')
    g = np.arange(length, dtype=np.float32)
    g **= 2
    g /= -np.float32(2) * sigma ** 2
    g = np.exp(g)
    return np.float32(1) - g


def np_filter_coefficient(coef: np.ndarray, width_frac: float, axis=-1) -> np.ndarray:
    """core.py:749-754 + :701-722.  sigma uses the OTHER axis' length (coef.shape[axis + 1]); the notch is
    indexed by packed-rfft array position, not by frequency."""
    sigma = coef.shape[axis + 1] * width_frac
    spec = rfft(coef, axis=axis)
    g = np_notch(coef.shape[axis], sigma)
    if axis == -2:
        g = g.reshape(-1, 1)
    spec *= g
    return irfft(spec, axis=axis)


def filter_subband(img, sigma, level, wavelet, axes=-1):
    """core.py:840-940, numpy branch :927-940."""
    level = None if level == 0 else level
    d_type = img.dtype
    rows, cols = img.shape
    if isinstance(axes, int):
        axes = (axes,)
    coeffs = pywt.wavedec2(img, wavelet, mode='symmetric', level=level, axes=(-2, -1))
    for i in range(1, len(coeffs)):
        ch, cv, cd = coeffs[i]
        if -1 in axes:
            ch = np_filter_coefficient(ch, sigma / rows, axis=-1)
        if -2 in axes:
            cv = np_filter_coefficient(cv, sigma / cols, axis=-2)
        coeffs[i] = (ch, cv, cd)
    return pywt.waverec2(coeffs, wavelet, mode='symmetric', axes=(-2, -1)).astype(d_type)


def filter_streak_dual_band(img, sigma1, sigma2, level, wavelet, threshold=None, axes=-1):
    """core.py:943-979 with use_thresholding=False (the only reachable variant)."""
    if (sigma1 > 0 and sigma1 == sigma2) or (threshold is not None and threshold <= 0):
        return filter_subband(img, sigma1, level, wavelet, axes=axes)
    img = filter_subband(img, sigma1, level, wavelet, axes=axes)
    return filter_subband(img, sigma2, level, wavelet, axes=axes)


def padded_geometry(shape, sigma, padding_mode):
    """core.py:1084-1096: returns base_pad, pad_y, pad_x."""
    pad_y, pad_x = [s % 2 for s in shape]
    if isinstance(padding_mode, str):
        padding_mode = padding_mode.lower()
    if padding_mode not in PAD_MODES:
        raise RuntimeError(f"Unsupported padding mode: {padding_mode}")
    base_pad = calculate_pad_size(shape=shape, sigma=max(sigma))
    min_len = 34
    if shape[0] + 2 * base_pad + pad_y < min_len:
        pad_y = min_len - (shape[0] + 2 * base_pad)
    if shape[1] + 2 * base_pad + pad_x < min_len:
        pad_x = min_len - (shape[1] + 2 * base_pad)
    return base_pad, pad_y, pad_x


def butter_lowpass_filter(img, cutoff_frequency, order=1):
    """core.py:493-499 (scipy.signal is the real library here, as in the reference)."""
    from scipy.signal import butter, sosfiltfilt
    d_type = img.dtype
    sos = butter(order, cutoff_frequency, output='sos')
    return sosfiltfilt(sos, img).astype(d_type)


def correct_bleaching(img, frequency, clip_min, clip_med, clip_max, max_method=False):
    """core.py:501-559, numpy branch (float32 throughout; the numexpr expression has the same order)."""
    clip_min_lb = np.log1p(1)
    if clip_min < clip_min_lb:
        clip_min = clip_min_lb
    if max_method:                                                                 # core.py:533-545
        fy, fx = np.max(img, axis=1), np.max(img, axis=0)
        fy[fy == 0] = clip_med
        fx[fx == 0] = clip_med
        fy, fx = np.clip(fy, clip_min, clip_max), np.clip(fx, clip_min, clip_max)
        fy, fx = butter_lowpass_filter(fy, frequency), butter_lowpass_filter(fx, frequency)
        img_filter = np.dot(fy.reshape(len(fy), 1), fx.reshape(1, len(fx)))
    else:
        img_filter = img.copy()
        img_filter[img_filter == 0] = clip_med
        np.clip(img_filter, clip_min, clip_max, out=img_filter)
        img_filter = butter_lowpass_filter(img_filter, frequency)
    img_filter_max = np.max(img_filter)
    img = img / img_filter
    img *= img_filter_max
    return img


def sosfiltfilt_order1_restated(x, sos, zi0):
    """scipy.signal.sosfiltfilt for ONE first-order section [b0, b1, 0, 1, a1, 0] along the last axis of a float32 array,
    restated operation for operation — the sequence csrc/bleach.cu executes (odd extension by 6 samples in float32,
    direct form II transposed in float64 with separate multiplies and adds, start value zi * first sample, the reversed
    second pass, edges dropped).  tests/test_oracle.py pins it bit-for-bit against scipy."""
    b0, b1, a1 = float(sos[0, 0]), float(sos[0, 1]), float(sos[0, 4])
    edge = 6
    x = np.asarray(x, dtype=np.float32)
    left = (np.float32(2) * x[..., :1] - x[..., edge:0:-1]).astype(np.float32)
    right = (np.float32(2) * x[..., -1:] - x[..., -2:-(edge + 2):-1]).astype(np.float32)
    ext = np.concatenate([left, x, right], axis=-1).astype(np.float64)

    def one_pass(sig, start):
        y = np.empty_like(sig)
        z = zi0 * start
        for n in range(sig.shape[-1]):              # vectorised over rows, sequential along the axis
            xn = sig[..., n]
            yn = b0 * xn + z
            z = b1 * xn - a1 * yn
            y[..., n] = yn
        return y
    y = one_pass(ext, ext[..., 0])
    y = one_pass(y[..., ::-1], y[..., -1])[..., ::-1]
    return y[..., edge:-edge]


def _box_morph(mask, k, erode):
    """cv2.dilate / cv2.erode with ones((k, k)) and the default anchor (k // 2, k // 2): the window covers the offsets
    [-(k // 2), k - 1 - k // 2] on both axes for both operations; outside the image dilate sees 0 and erode sees 1
    (cv2.morphologyDefaultBorderValue).  Separable window counts through cumulative sums."""
    lo, hi = -(k // 2), k - 1 - k // 2
    m = mask.astype(np.int64)
    for axis in (1, 0):
        n = m.shape[axis]
        pre = np.concatenate([np.zeros_like(np.take(m, [0], axis=axis)), np.cumsum(m, axis=axis)], axis=axis)
        idx = np.arange(n)
        a, b = np.clip(idx + lo, 0, n), np.clip(idx + hi + 1, 0, n)
        b = np.maximum(a, b)
        ones = np.take(pre, b, axis=axis) - np.take(pre, a, axis=axis)
        length = (b - a).reshape((-1, 1) if axis == 0 else (1, -1))
        m = (ones == length).astype(np.int64) if erode else (ones > 0).astype(np.int64)
    return m.astype(np.uint8)


def get_img_mask(img, threshold, close_steps=50, open_steps=500):
    """core.py:475-489: threshold, MORPH_CLOSE (dilate, erode), MORPH_OPEN (erode, dilate), then the four 4-connected
    floodFill calls from the corners of the inverted mask: background components that hold no corner pixel are holes and
    join the mask.  (cv2 itself is the check: tests/test_oracle.py.)"""
    from scipy import ndimage
    mask = (img > threshold).astype(np.uint8)
    mask = _box_morph(_box_morph(mask, close_steps, False), close_steps, True)
    mask = _box_morph(_box_morph(mask, open_steps, True), open_steps, False).astype(bool)
    lab, _ = ndimage.label(~mask)                       # default structure: 4-connectivity
    corner = {int(lab[y, x]) for y in (0, -1) for x in (0, -1)} - {0}
    holes = (lab > 0) & ~np.isin(lab, list(corner))
    return mask | holes


def filter_streaks(img, sigma=(250, 250), level=0, wavelet='db9', crossover=10, threshold=None,
                   padding_mode="wrap", bidirectional=False, log1p_normalization_needed=True,
                   return_log_domain=False, bleach_correction_frequency=None, bleach_correction_clip_min=None,
                   bleach_correction_clip_med=None, bleach_correction_clip_max=None, bleach_correction_max_method=False,
                   enable_masking=False, close_steps=50, open_steps=500):
    """core.py:982-1159 (multi-Otsu clip levels through the restated threshold_multiotsu below)."""
    if not isinstance(sigma, (tuple, list)):
        sigma = (sigma,) * 2
    s1, s2 = sigma
    if s1 == s2 == 0 and bleach_correction_frequency is None:
        return img
    d_type = img.dtype
    if log1p_normalization_needed:
        img = log1p_f32(img.astype(np.float32))
    if (bleach_correction_frequency is not None and (bleach_correction_clip_min is None or bleach_correction_clip_med is None
                                                     or bleach_correction_clip_max is None)) or \
            (enable_masking and bleach_correction_clip_med is None):                             # core.py:1066-1077
        lb, mb, ub = threshold_multiotsu(img, classes=4)
        if bleach_correction_clip_min is None:
            bleach_correction_clip_min = lb
        if bleach_correction_clip_med is None:
            bleach_correction_clip_med = mb
        if bleach_correction_clip_max is None:
            bleach_correction_clip_max = ub
    if enable_masking and close_steps is not None and open_steps is not None:      # core.py:1079-1080
        img = img.copy() if not log1p_normalization_needed else img
        img *= get_img_mask(img, bleach_correction_clip_med, close_steps=close_steps, open_steps=open_steps)
    if not s1 == s2 == 0:                                                          # core.py:1081 (no padding otherwise)
        shape = img.shape
        base_pad, pad_y, pad_x = padded_geometry(shape, sigma, padding_mode)
        if pad_y > 0 or pad_x > 0 or base_pad > 0:
            mode = padding_mode.lower() if padding_mode else 'reflect'
            if mode == 'constant' and bleach_correction_clip_min is not None:      # core.py:1101-1105
                img = np.pad(img, ((base_pad, base_pad + pad_y), (base_pad, base_pad + pad_x)), mode='constant',
                             constant_values=np.log1p(bleach_correction_clip_min))
            else:
                img = np.pad(img, ((base_pad, base_pad + pad_y), (base_pad, base_pad + pad_x)), mode=mode)
        img = filter_streak_dual_band(img, s1, s2, level, wavelet, threshold, axes=(-1, -2) if bidirectional else -1)
        if pad_y > 0 or pad_x > 0 or base_pad > 0:
            img = img[base_pad: img.shape[0] - (base_pad + pad_y), base_pad: img.shape[1] - (base_pad + pad_x)]
            assert img.shape == shape
    if bleach_correction_frequency is not None:                                    # core.py:1131-1139
        img = correct_bleaching(np.ascontiguousarray(img), bleach_correction_frequency, bleach_correction_clip_min,
                                bleach_correction_clip_med, bleach_correction_clip_max, bleach_correction_max_method)
    if return_log_domain:
        return np.ascontiguousarray(img)
    if log1p_normalization_needed:
        img = expm1_f32(img) if img.dtype == np.float32 else np.expm1(img).astype(np.float32)
    if np.dtype(d_type).kind in "ui":
        img = np.rint(img)
        info = np.iinfo(d_type)
        np.clip(img, info.min, info.max, out=img)
    if img.dtype != d_type:
        img = img.astype(d_type)
    return img


# ---- process_img pieces ------------------------------------------------------------------------
def is_uniform_2d(arr):
    """core.py:106-121."""
    if arr.size == 0:
        return None
    return bool((arr == arr.flat[0]).all())


def normalize_flat(flat):
    """core.py:2047-2049."""
    f = flat.astype(np.float32)
    return f / f.max()


def convert_to_16bit_fun(img):
    """core.py:397-399 (clip then C-cast: truncation toward zero)."""
    img = np.clip(img, 0, 65535)
    return img.astype(np.uint16)


def convert_to_8bit_fun(img, bit_shift_to_right=8):
    """core.py:402-423."""
    if img is None or img.dtype == np.uint8:
        return img
    if img.dtype != np.uint16:
        img = convert_to_16bit_fun(img)
    if bit_shift_to_right is None:
        bit_shift_to_right = 8
    if not 0 <= bit_shift_to_right < 9:
        raise RuntimeError("right shift should be between 0 and 8")
    lower = 2 ** bit_shift_to_right
    img = np.where((0 < img) & (img < lower), 1, img >> bit_shift_to_right)
    img = np.clip(img, 0, 255)
    return img.astype(np.uint8)


def block_reduce(img, block, method):
    """skimage.measure.block_reduce(img, block_size=block, func=np.{max,min,mean,median}), cval=0 padding of the
    trailing edges (core.py:1286-1300).  mean/median return float64 like numpy."""
    func = {'max': np.max, 'min': np.min, 'mean': np.mean, 'median': np.median}[method]
    by, bx = block
    py, px = (-img.shape[0]) % by, (-img.shape[1]) % bx
    if py or px:
        img = np.pad(img, ((0, py), (0, px)), mode='constant', constant_values=0)
    v = img.reshape(img.shape[0] // by, by, img.shape[1] // bx, bx)
    return func(v, axis=(1, 3))


def calculate_down_sampled_size(tile_size, down_sample):
    """core.py:1162-1170."""
    return tuple(ceil(s / d) if d is not None else s for s, d in zip(tile_size, down_sample))


def gaussian_blur_5x5(img):
    """cv2.GaussianBlur(img, ksize=(5,5), sigmaX=1, sigmaY=1) — core.py:1284 (intended semantics)."""
    import cv2
    return cv2.GaussianBlur(img, ksize=(5, 5), sigmaX=1, sigmaY=1)


# ---- lightsheet_correct ------------------------------------------------------------------------
def _percentile_numba(data, q):
    """numba's np.percentile (numba/np/arraymath.py _collect_percentiles_inner): linear interpolation between
    closest ranks, float64; called through pystripe/lightsheet_correct.py:240-242."""
    n = data.size
    if n == 0:
        return 0
    a = np.sort(data.astype(np.float64), axis=None)
    if n == 1:
        return a[0]
    if q == 100:
        return a[-1]
    rank = 1 + (n - 1) * (q / 100.0)
    f = int(np.floor(rank))
    m = rank - f
    lower = a[f - 1]
    upper = a[min(f, n - 1)]
    return lower * (1 - m) + upper * m


def local_percentile(source, percentile, selem, spacing=None, step=None, interpolate=1, dtype=None):
    """lightsheet_correct.py:245-312 -> apply_local_function :113-237 for a 2-D plane (the reshape to (H,W,1)
    and the unit third axis are dropped: they do not change any number)."""
    q = 100 * percentile
    shape = source.shape
    if spacing is None:
        spacing = selem
    if step is None:
        step = (None, None)
    n_centers = tuple(s // h for s, h in zip(shape, spacing))
    left = tuple((s - (n - 1) * h) // 2 for s, n, h in zip(shape, n_centers, spacing))
    cy = list(range(left[0], shape[0], spacing[0]))
    cx = list(range(left[1], shape[1], spacing[1]))
    rdtype = source.dtype if dtype is None else dtype
    # NB: meshgrid over range(le, s, h) may yield MORE centres than n_centers when (s - le) > n*h is impossible
    # by construction ((n-1)*h + le < s <= n*h + le), so len(range) == n.
    assert len(cy) == n_centers[0] and len(cx) == n_centers[1]
    results = np.zeros(n_centers, dtype=rdtype)
    hl = tuple(h // 2 for h in selem)
    hr = tuple(h - l for h, l in zip(selem, hl))
    for iy, yy in enumerate(cy):
        sy = slice(max(0, yy - hl[0]), min(yy + hr[0], shape[0]), step[0])
        for ix, xx in enumerate(cx):
            sx = slice(max(0, xx - hl[1]), min(xx + hr[1], shape[1]), step[1])
            results[iy, ix] = _percentile_numba(source[sy, sx], q)   # float64 -> rdtype C cast (truncation)
    if interpolate:
        zoom = tuple(float(s) / float(r) for s, r in zip(shape, results.shape))
        results = ndi_zoom(results, zoom=zoom, order=interpolate)
    return results


def correct_lightsheet(img, percentile=0.25, artifact_length=150, background_window_size=200,
                       lightsheet_vs_background=2.0, d_type=None):
    """lightsheet_correct.py:31-106 as called from core.py:1333-1348."""
    if d_type is None:
        d_type = img.dtype
    ls = local_percentile(img, percentile, selem=(1, artifact_length), dtype=d_type)
    bg = local_percentile(img, percentile, selem=(background_window_size,) * 2, spacing=(25, 25), step=(2, 2),
                          interpolate=1, dtype=d_type)
    img = img.copy()
    if isinstance(lightsheet_vs_background, float) and all(
            a.dtype in (np.uint8, np.uint16) for a in (img, ls, bg)):
        img -= np.minimum(img, np.minimum(ls, bg * int(lightsheet_vs_background)))
    else:
        img -= np.minimum(img, np.minimum(ls, bg * lightsheet_vs_background)).astype(img.dtype)
    return img


# ---- skimage.transform.resize (absent here; restated) -------------------------------------------
def skimage_resize(image, output_shape, preserve_range=True, anti_aliasing=True, order=1, mode='reflect', cval=0,
                   clip=True):
    """skimage.transform.resize as pystripe calls it (core.py:1356-1359), restated from skimage/transform/_warps.py
    (0.19 ... 0.25: `resize` -> scipy.ndimage.zoom(grid_mode=True), `_clip_warp_output`).  skimage is not installed and
    not pinned by the reference: wrapper parity is UNPINNED; the interpolation itself is the real scipy.ndimage.zoom."""
    from scipy import ndimage as ndi
    image = np.asarray(image)
    output_shape = tuple(int(v) for v in output_shape)
    assert image.ndim == len(output_shape) == 2 and preserve_range and order == 1
    if image.dtype == np.float16:
        image = image.astype(np.float32)
    if image.dtype.char not in 'df':                       # convert_to_float(image, preserve_range=True)
        image = image.astype(np.float64)
    factors = np.divide(image.shape, output_shape)
    ndi_mode = {'constant': 'constant', 'edge': 'nearest', 'symmetric': 'reflect', 'reflect': 'mirror', 'wrap': 'wrap'}[mode]
    if anti_aliasing:
        sigma = np.maximum(0, (factors - 1) / 2)
        filtered = ndi.gaussian_filter(image, sigma, cval=cval, mode=ndi_mode)
    else:
        filtered = image
    zoom_factors = [1 / f for f in factors]
    out = ndi.zoom(filtered, zoom_factors, order=order, mode=ndi_mode, cval=cval, grid_mode=True)
    if clip:                                               # _clip_warp_output (mode != 'constant': cval plays no part)
        min_val, max_val = np.min(image), np.max(image)
        np.clip(out, min_val, max_val, out=out)
    return out


# ---- isotropic down-sampling of the post-stitch path (parallel_image_processor.py:371-433) -------
def down_sample_xy(img, target_shape, down_sampling_methods):
    """parallel_image_processor.py:371-385 (methods as 'max' / 'mean' / None)"""
    if is_uniform_2d(img):
        return np.zeros(target_shape, dtype=np.float32)
    img = img.astype(np.float32)
    for y_method, x_method in down_sampling_methods:
        if y_method is not None and ceil(img.shape[0] / 2) >= target_shape[0]:
            img = block_reduce(img, (2, 1), y_method)
        if x_method is not None and ceil(img.shape[1] / 2) >= target_shape[1]:
            img = block_reduce(img, (1, 2), x_method)
    img = skimage_resize(img, target_shape, preserve_range=True, anti_aliasing=True)
    return img.astype(np.float32)


def down_sample_z(z_stack, down_sampling_method_z, down_sampled_dtype="float32", post_processed_dtype=None):
    """parallel_image_processor.py:412-433"""
    if (z_stack == z_stack.flat[0]).all():
        return np.zeros(z_stack.shape[1:], dtype=np.float32)
    for z_method in down_sampling_method_z:
        if z_method is not None and z_stack.shape[0] > 1:
            func = {'max': np.max, 'mean': np.mean}[z_method]
            if z_stack.shape[0] % 2:
                z_stack = np.concatenate([z_stack, np.zeros((1,) + z_stack.shape[1:], z_stack.dtype)])
            z_stack = func(z_stack.reshape(z_stack.shape[0] // 2, 2, *z_stack.shape[1:]), axis=1)
    assert z_stack.shape[0] == 1
    img = z_stack[0]
    dt = np.dtype(down_sampled_dtype)
    if dt != np.float32:
        if dt == np.uint16:
            img = convert_to_16bit_fun(img)
        elif dt == np.uint8:
            if post_processed_dtype is not None and np.dtype(post_processed_dtype) == np.uint8:
                img = img.astype(np.uint8)
            else:
                img = convert_to_8bit_fun(img)
        else:
            raise RuntimeError
    return img


# ---- process_img -------------------------------------------------------------------------------
def process_img(img, flat=None, gaussian_filter_2d=False, down_sample=None, down_sample_method='max',
                tile_size=None, new_size=None, sigma=(0, 0), level=0, wavelet='coif15', threshold=None,
                padding_mode="wrap", bidirectional=False, log1p_normalization_needed=True, dark=0,
                lightsheet=False, artifact_length=150, background_window_size=200, percentile=0.25,
                lightsheet_vs_background=2.0, rotate=0, flip_upside_down=False, convert_to_16bit=False,
                convert_to_8bit=False, bit_shift_to_right=8, d_type=None, quirks=False,
                bleach_correction_frequency=None, bleach_correction_clip_min=None, bleach_correction_clip_med=None,
                bleach_correction_clip_max=None, bleach_correction_max_method=False):
    """core.py:1190-1381 (order of operations preserved; bleach / dark-edge options not restated)."""
    if tile_size is None:
        tile_size = img.shape
    tile_size = tuple(tile_size)
    if new_size is not None:
        new_size = tuple(int(v) for v in new_size)
    if d_type is None:
        d_type = img.dtype
    d_type = np.dtype(d_type)

    if is_uniform_2d(img):                                             # :1232-1246
        if new_size is not None:
            tile_size = new_size
        elif down_sample is not None:
            tile_size = calculate_down_sampled_size(tile_size, down_sample)
        if rotate in (90, 270):
            tile_size = (tile_size[1], tile_size[0])
        out_t = np.uint16 if convert_to_16bit else (np.uint8 if convert_to_8bit else d_type)
        return np.zeros(tile_size, dtype=out_t)

    if flat is not None:                                               # :1248-1254
        if tile_size == flat.shape:
            if quirks and img.dtype.kind in "ui":
                raise TypeError("ufunc 'divide' output cannot be cast (uint16 /= float32), core.py:1250")
            img = img.astype(np.float32) / flat.astype(np.float32)
    if gaussian_filter_2d and not quirks:                              # :1280-1284 (result discarded as written)
        img = gaussian_blur_5x5(img)
    if down_sample is not None:                                        # :1286-1300
        method = down_sample_method.lower()
        if method not in ('min', 'max', 'mean', 'median'):
            raise RuntimeError(f"unsupported down-sampling method: {down_sample_method}")
        img = block_reduce(img, down_sample, method)
        tile_size = tuple(calculate_down_sampled_size(tile_size, down_sample))
    if bleach_correction_frequency is not None or tuple(sigma) > (0, 0):   # :1302-1320
        img = filter_streaks(img, sigma=sigma, level=level, wavelet=wavelet, threshold=threshold,
                             padding_mode=padding_mode, bidirectional=bidirectional,
                             log1p_normalization_needed=log1p_normalization_needed,
                             bleach_correction_frequency=bleach_correction_frequency,
                             bleach_correction_clip_min=bleach_correction_clip_min,
                             bleach_correction_clip_med=bleach_correction_clip_med,
                             bleach_correction_clip_max=bleach_correction_clip_max,
                             bleach_correction_max_method=bleach_correction_max_method)
    if dark is not None and dark > 0:                                  # :1324-1330 (numpy branch)
        img = np.where(img > dark, img - dark, 0)
    if lightsheet:                                                     # :1333-1348
        img = correct_lightsheet(img, percentile, artifact_length, background_window_size,
                                 lightsheet_vs_background, d_type=d_type)
    if new_size is not None and tile_size < new_size:                  # :1356-1359 (tuple comparison)
        img = skimage_resize(img, new_size, preserve_range=True, anti_aliasing=True)
    elif new_size is not None and tile_size > new_size:
        img = skimage_resize(img, new_size, preserve_range=True, anti_aliasing=False)
    if convert_to_16bit and img.dtype != np.uint16:                    # :1361-1369
        img = convert_to_16bit_fun(img)
    elif convert_to_8bit and img.dtype != np.uint8:
        img = convert_to_8bit_fun(img, bit_shift_to_right=bit_shift_to_right)
    elif d_type.kind in "ui":
        img = np.clip(img, np.iinfo(d_type).min, np.iinfo(d_type).max).astype(d_type)
    else:
        img = img.astype(d_type)
    if flip_upside_down:                                               # :1371-1379
        img = np.flipud(img)
    if rotate in (90, 180, 270):
        img = np.rot90(img, rotate // 90)
    return np.ascontiguousarray(img)


# ---- stack statistics (process_images.py:320-331, 594-659) ---------------------------------------
def threshold_multiotsu(image, classes=4, nbins=256):
    """skimage.filters.threshold_multiotsu (skimage is absent: PARITY UNPINNED against skimage itself).  Restated from the
    published algorithm — skimage/filters/thresholding.py + _multiotsu.pyx: 256-bin np.histogram of the image, float32
    probabilities, cumulative zeroth / first moments, between-class-variance look-up table, exhaustive recursive search
    (first maximum wins).  Deliberately written as the plain scalar loops of the Cython source; the product's vectorised
    search (pystripe/stack_stats.py) and a float64 brute force are checked against it in tests/test_stack_stats.py."""
    image = np.asarray(image)
    hist, bin_edges = np.histogram(image.reshape(-1), bins=nbins, range=None)
    bin_centers = (bin_edges[:-1] + bin_edges[1:]) / 2.0
    prob = (hist / np.sum(hist)).astype(np.float32)
    if np.count_nonzero(prob) < classes:
        raise ValueError("cannot be thresholded in %d classes" % classes)
    f32 = np.float32
    n = nbins
    thresh_count = classes - 1
    var = np.zeros(n * (n + 1) // 2, np.float32)
    m0 = np.empty(n, np.float32)
    m1 = np.empty(n, np.float32)
    m0[0] = prob[0]
    m1[0] = prob[0]
    for i in range(1, n):
        m0[i] = m0[i - 1] + prob[i]
        m1[i] = m1[i - 1] + f32(i) * prob[i]
        if m0[i] > 0:
            var[i] = (m1[i] * m1[i]) / m0[i]
    idx = n
    for i in range(1, n):
        zi = m0[i:] - m0[i - 1]
        fi = m1[i:] - m1[i - 1]
        with np.errstate(divide="ignore", invalid="ignore"):
            row = np.where(zi > 0, (fi * fi) / np.where(zi > 0, zi, f32(1)), f32(0)).astype(np.float32)
        var[idx:idx + n - i] = row
        idx += n - i

    def lut(i, j):
        return var[(i * (2 * n - i + 1)) // 2 + j - i]

    # the recursion of _set_thresh_indices_lut for three thresholds; the innermost loop (c2) is evaluated as one float32
    # vector expression with the same association order: (var(0, c0) + var(c2 + 1, n - 1)) + var(c0 + 1, c1) + var(c1 + 1, c2)
    assert thresh_count == 3
    best = [f32(0), None]
    tail = np.array([lut(c2 + 1, n - 1) for c2 in range(n - 1)], dtype=np.float32)
    for c0 in range(0, n - 3):
        head = lut(0, c0)
        for c1 in range(c0 + 1, n - 2):
            c2 = np.arange(c1 + 1, n - 1)
            base = (c1 + 1) * (2 * n - (c1 + 1) + 1) // 2 - (c1 + 1)
            sigma = ((head + tail[c2]).astype(np.float32) + lut(c0 + 1, c1)).astype(np.float32)
            sigma = (sigma + var[base + c2]).astype(np.float32)
            k = int(np.argmax(sigma))                      # first maximum, as the strict `>` of the scalar loop keeps
            if sigma[k] > best[0]:
                best[0] = sigma[k]
                best[1] = [c0, c1, int(c2[k])]
    return tuple(np.float32(bin_centers[i]) for i in best[1])


def estimate_bit_shift(img, threshold, percentile=99.9):
    """process_images.py:320-331 on the float32 log image."""
    try:
        sel = img[img > threshold]
        if sel.size == 0:
            raise ValueError
        upper_bound = _percentile_numba(sel, percentile)
    except (ValueError, AssertionError):
        upper_bound = np.max(img)
    upper_bound = int(np.round(np.expm1(upper_bound)))
    for b in range(0, 9):
        if 256 * 2 ** b >= upper_bound:
            return b
    return 8
