"""Stack statistics on the GPU: the bit shift, dark level and bleach-correction clip levels the
reference derives from sample planes before it processes a channel (`process_images.py:594-659`
`estimate_img_related_params`, `:320-331` `estimate_bit_shift`).

Everything that routine computes is a function of the intensity HISTOGRAM of `log1p(plane)`:
    threshold_multiotsu(log1p(img), classes=4)            skimage: 256-bin histogram -> exhaustive search over bin triples
    prctl(img[img > clip_max], 99.99)                      numba's np.percentile over the pixels above the top threshold
and `log1p` is monotone on integer pixels.  So the GPU counts the integer pixels exactly (`b2s_histogram`, 65 536 bins,
any number of planes), and the host maps the occupied bins through numpy's own `log1p` / `np.histogram` / percentile
arithmetic — bit for bit what the reference gets from the pixels, at a cost that no longer depends on the number of pixels.

Histograms add.  With one process per GPU, each rank counts its Z-shard and `allreduce_histogram` sums 65 536 int64
counters over the ranks (NCCL on the device, gloo on the host): the only collective on this path, and what lets the
statistics use EVERY plane of a stack instead of the reference's three samples (`whole_stack_params`).
"""
import ctypes as C
from math import floor
from typing import Callable, Optional, Sequence, Tuple

import numpy as np

from . import _native

N_BINS = 65536


# ------------------------------------------------------------------------------------------------ histogram (GPU)
def histogram(planes, per_plane: bool = False, device: int = None, out=None):
    """exact histogram of uint8 / uint16 planes ((H, W) or (Z, H, W); numpy or CUDA torch tensor): int64[65536], or
    int64[Z, 65536] with per_plane=True.  `out` (same kind as the result) is added to."""
    is_torch = _native._is_torch(planes)
    if is_torch:
        import torch
        t = planes.contiguous()
        if not t.is_cuda:
            raise TypeError("torch tensors must live on a CUDA device (pass numpy arrays for host data)")
        code = {torch.uint8: _native.U8, torch.uint16: _native.U16}.get(t.dtype)
        dev = t.device.index or 0
        shape = tuple(t.shape)
    else:
        t = np.ascontiguousarray(planes)
        if not t.dtype.isnative:
            t = t.astype(t.dtype.newbyteorder("="))
        code = {np.dtype(np.uint8): _native.U8, np.dtype(np.uint16): _native.U16}.get(t.dtype)
        dev = device
        if dev is None:
            from .core import _device_of
            dev = _device_of(t)
        shape = t.shape
    if code is None:
        raise TypeError("histogram takes uint8 or uint16 planes (the dtypes the reference's tiles have)")
    if len(shape) not in (2, 3):
        raise ValueError("expected an (H, W) plane or a (Z, H, W) stack")
    n = 1 if len(shape) == 2 else shape[0]
    elems = shape[-1] * shape[-2]
    ctx = _native.context(dev)
    hshape = (n, N_BINS) if per_plane else (N_BINS,)
    if is_torch:
        import torch
        h = torch.zeros(hshape, dtype=torch.int64, device=t.device) if out is None else out
        stream = torch.cuda.current_stream(t.device).cuda_stream
        ctx.check(_native.lib().b2s_histogram(ctx._h, C.c_void_p(t.data_ptr()), 1, code, elems, n, C.c_void_p(h.data_ptr()), 1,
                                              int(per_plane), C.c_void_p(stream)))
        return h
    h = np.zeros(hshape, dtype=np.int64) if out is None else out
    ctx.check(_native.lib().b2s_histogram(ctx._h, C.c_void_p(t.ctypes.data), 0, code, elems, n, C.c_void_p(h.ctypes.data), 0,
                                          int(per_plane), None))
    return h


def allreduce_histogram(h):
    """sum a histogram over the ranks of the default process group (no-op without one).  CUDA tensors go through NCCL,
    numpy arrays through the group's own backend (gloo on CPU).  Returns the same kind of array it was given."""
    try:
        import torch
        import torch.distributed as dist
    except ImportError:
        return h
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return h
    if _native._is_torch(h):
        dist.all_reduce(h, op=dist.ReduceOp.SUM)
        return h
    backend = dist.get_backend()
    t = torch.from_numpy(np.ascontiguousarray(h, dtype=np.int64))
    if backend == "nccl":
        t = t.cuda()
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.cpu().numpy()
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.numpy()


# ------------------------------------------------------------------------------------------------ host: numpy on the bins
def _log_values(counts: np.ndarray, log1p: bool = True) -> Tuple[np.ndarray, np.ndarray]:
    """(values, weights) of the occupied bins as the reference's float32 image holds them: log1p_jit(img, dtype=float32)."""
    counts = np.asarray(counts, dtype=np.int64).reshape(-1)
    occupied = np.flatnonzero(counts)
    v = occupied.astype(np.float32)
    if log1p:
        v = np.log1p(v, dtype=np.float32)
    return v, counts[occupied]


def _multiotsu_indices(prob: np.ndarray, thresh_count: int = 3) -> np.ndarray:
    """skimage.filters._multiotsu._get_multiotsu_thresh_indices_lut restated in numpy, float32 like the Cython code:
    cumulative zeroth / first moments, the between-class variance look-up table var(i, j) = m1(i..j)^2 / m0(i..j), then the
    exhaustive search for the bin triple that maximises  var(0, c0) + var(c2 + 1, n - 1) + var(c0 + 1, c1) + var(c1 + 1, c2)
    (float32 sums in that order; the first maximum in lexicographic order wins).  thresh_count is 3 (classes=4)."""
    if thresh_count != 3:
        raise NotImplementedError("four classes (three thresholds) is what the reference asks for")
    prob = np.asarray(prob, dtype=np.float32)
    n = prob.size
    f32 = np.float32
    idx = np.arange(n, dtype=np.float32)
    m0 = np.empty(n, f32)
    m1 = np.empty(n, f32)
    m0[0] = prob[0]
    m1[0] = prob[0]
    for i in range(1, n):                                    # sequential float32 accumulation, as the Cython loop
        m0[i] = m0[i - 1] + prob[i]
        m1[i] = m1[i - 1] + idx[i] * prob[i]
    # var[i, j] for i <= j (zero where the class is empty)
    var = np.zeros((n, n), f32)
    with np.errstate(divide="ignore", invalid="ignore"):
        z = m0.copy()
        var[0, :] = np.where(z > 0, (m1 * m1) / np.where(z > 0, z, f32(1)), f32(0))
        var[0, 0] = 0                                        # the LUT's first row starts at i = 1 (var_btwcls[0] stays 0)
        for i in range(1, n):
            z = m0[i:] - m0[i - 1]
            f = m1[i:] - m1[i - 1]
            var[i, i:] = np.where(z > 0, (f * f) / np.where(z > 0, z, f32(1)), f32(0))
    best, best_idx = f32(0), None
    last = var[:, n - 1]                                     # var(c2 + 1, n - 1)
    for c0 in range(0, n - 3):
        first = var[0, c0]
        # c1 in (c0, n - 2), c2 in (c1, n - 1)
        c1 = np.arange(c0 + 1, n - 2)
        c2 = np.arange(c0 + 2, n - 1)
        s = (first + last[c2 + 1])[None, :].astype(f32)                     # (1, c2)
        s = s + var[c0 + 1, c1][:, None]                                   # + var(c0 + 1, c1)
        mid = var[c1[:, None] + 1, c2[None, :]]                            # var(c1 + 1, c2)
        s = (s + mid).astype(f32)
        valid = c2[None, :] > c1[:, None]
        s = np.where(valid, s, f32(-1))
        k = int(np.argmax(s))                                              # first maximum in (c1, c2) lexicographic order
        v = s.reshape(-1)[k]
        if v > best:
            best = v
            best_idx = (c0, int(c1[k // c2.size]), int(c2[k % c2.size]))
    if best_idx is None:
        best_idx = (0, 1, 2)
    return np.array(best_idx, dtype=np.intp)


def threshold_multiotsu_from_histogram(counts, classes: int = 4, nbins: int = 256, log1p: bool = True):
    """`skimage.filters.threshold_multiotsu(log1p(img), classes=4)` (pystripe/core.py:1070, process_images.py:625) from
    the exact integer histogram of `img`: the 256-bin histogram of the float32 image is numpy's own np.histogram over the
    occupied values, weighted by their counts (bin assignment is element-wise, so this is what np.histogram returns for the
    pixels themselves).  Returns three float32 thresholds."""
    v, w = _log_values(counts, log1p)
    if v.size == 0:
        raise ValueError("empty histogram")
    if v[0] == v[-1]:
        raise ValueError("threshold_multiotsu is expected to work with images having more than one value")
    hist, bin_edges = np.histogram(v, bins=nbins, range=None, weights=w)
    bin_centers = (bin_edges[:-1] + bin_edges[1:]) / 2.0
    prob = (hist / np.sum(hist)).astype(np.float32)
    if np.count_nonzero(prob) < classes:
        raise ValueError(f"After discretization into bins, the input image has only {np.count_nonzero(prob)} different "
                         f"values. It cannot be thresholded in {classes} classes.")
    idx = _multiotsu_indices(prob, classes - 1)
    return tuple(np.float32(bin_centers[i]) for i in idx)


def percentile_above_from_histogram(counts, threshold, q: float, log1p: bool = True):
    """`prctl(img[img > threshold], q)` (numba's np.percentile: linear interpolation between the closest ranks, float64)
    over the float32 image log1p(pixels), from the histogram.  Raises ValueError when no pixel is above the threshold, as
    the masked percentile of an empty array does."""
    v, w = _log_values(counts, log1p)
    keep = v > threshold
    v, w = v[keep].astype(np.float64), w[keep]
    n = int(w.sum())
    if n == 0:
        raise ValueError("no pixel above the threshold")
    if n == 1:
        return v[0]
    if q == 100:
        return v[-1]
    cum = np.cumsum(w)
    rank = 1 + (n - 1) * (q / 100.0)
    f = int(np.floor(rank))
    m = rank - f
    lower = v[np.searchsorted(cum, f, side="left")]                 # the f-th smallest pixel (1-based)
    upper = v[np.searchsorted(cum, min(f + 1, n), side="left")]
    return lower * (1 - m) + upper * m


def estimate_bit_shift_from_histogram(counts, threshold, percentile: float = 99.9) -> int:
    """process_images.py:320-331 `estimate_bit_shift(log1p(img), threshold, percentile)`."""
    try:
        upper_bound = percentile_above_from_histogram(counts, threshold, percentile)
    except (ValueError, AssertionError):
        v, _ = _log_values(counts)
        upper_bound = v[-1]
    upper_bound = int(np.round(np.expm1(upper_bound)))
    right_bit_shift = 8
    for b in range(0, 9):
        if 256 * 2 ** b >= upper_bound:
            right_bit_shift = b
            break
    return right_bit_shift


# ------------------------------------------------------------------------------------------------ the reference's routine
def estimate_img_related_params(read_plane: Callable[[int], np.ndarray], n_planes: int,
                                need_16bit_to_8bit_conversion: bool = True, need_bleach_correction: bool = False,
                                tile_size: Optional[Sequence[int]] = None, new_tile_size: Optional[Sequence[int]] = None,
                                down_sampling_factor: Optional[Sequence[int]] = None):
    """process_images.py:594-650: three sample planes (25 %, 50 %, 75 % of the stack; a uniform or unusable plane moves on to
    the next one), per plane the four-class multi-Otsu thresholds of log1p(plane) and the bit shift from the 99.99th
    percentile above the top class; the bit shift is the maximum of the three, the clip levels are those of the last sample.
    `read_plane(z)` returns plane z (uint8 / uint16, numpy or CUDA tensor).  Returns
    (background, bit_shift, sigma, clip_min, clip_med, clip_max, frequency)."""
    sig, frequency = 0, None
    background, bit_shift, clip_min, clip_med, clip_max = 0, 8, None, None, None
    if need_16bit_to_8bit_conversion or need_bleach_correction:
        z = [floor(n_planes * 0.25), floor(n_planes * 0.5), floor(n_planes * 0.75)]
        shifts = []
        for i in range(3):
            while True:
                if z[i] >= n_planes:
                    raise ValueError("no usable sample plane: every plane from the sampling point on is uniform")
                try:
                    h = histogram(read_plane(z[i]))
                    h = h.cpu().numpy() if _native._is_torch(h) else h
                    assert np.count_nonzero(h) > 1                                    # `assert not is_uniform_2d(img)`
                    clip_min, clip_med, clip_max = threshold_multiotsu_from_histogram(h, classes=4)
                    shifts.append(estimate_bit_shift_from_histogram(h, threshold=clip_max, percentile=99.99))
                    break
                except (ValueError, AssertionError):
                    z[i] += 1
        bit_shift = max(shifts)
        if need_bleach_correction:
            background = int(np.round(np.expm1(clip_min)))
            if new_tile_size is not None:
                sig = min(new_tile_size)
            elif down_sampling_factor is not None:
                sig = min(new_tile_size) // min(down_sampling_factor)                 # (as written: raises for new_tile_size None)
            else:
                sig = min(tile_size)
    sigma = (int(sig * 2),) * 2
    return background, bit_shift, sigma, clip_min, clip_med, clip_max, frequency


def whole_stack_params(planes, need_bleach_correction: bool = False):
    """The same statistics over EVERY plane this rank holds, summed over all ranks: one histogram kernel pass over the
    stack and one all-reduce of 65 536 counters, instead of three sample planes.  `planes`: (Z, H, W) uint8 / uint16 array
    or CUDA tensor (this rank's Z-shard), or an iterable of such blocks.
    Returns dict(bit_shift, clip_min, clip_med, clip_max, background, pixels)."""
    blocks = [planes] if hasattr(planes, "shape") else planes
    h = None
    for b in blocks:
        h = histogram(b, out=h)
    h = allreduce_histogram(h)
    h = h.cpu().numpy() if _native._is_torch(h) else h
    clip_min, clip_med, clip_max = threshold_multiotsu_from_histogram(h, classes=4)
    return dict(bit_shift=estimate_bit_shift_from_histogram(h, clip_max, 99.99), clip_min=clip_min, clip_med=clip_med,
                clip_max=clip_max, background=int(np.round(np.expm1(clip_min))) if need_bleach_correction else 0,
                pixels=int(h.sum()))
