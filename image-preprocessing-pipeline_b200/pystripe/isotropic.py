"""Isotropic down-sampling of the post-stitch path on the GPU — the arithmetic
`parallel_image_processor.py:156-187, 371-435` runs per processed slice: alternating max / mean 2x1 and 1x2
`block_reduce` passes, an anti-aliased `skimage.transform.resize` to the target shape, then the approximate reduction
along z and the conversion to the requested dtype.  Function names follow the reference's method / variable names.

Arrays: numpy (host round trip through torch, which is plumbing only) or CUDA torch tensors (zero-copy).  No CPU fallback."""
import ctypes as C
from math import ceil
from typing import Sequence, Tuple

import numpy as np

from . import _native

_METHOD = {"max": 0, "mean": 2, None: -1}


def _name(m):
    """np_max / np_mean / None of the reference's method tuples -> 'max' / 'mean' / None."""
    if m is None or isinstance(m, str):
        return m
    return {"max": "max", "amax": "max", "mean": "mean"}[m.__name__]


def calculate_down_sampling_target(shape: Tuple[int, int], new_shape: Tuple[int, int], source_voxel, target_voxel,
                                   is_rotated: bool = False, alternating_downsampling_method: bool = True):
    """parallel_image_processor.py:156-187: returns (target_shape, down_sampling_methods as ('max' | 'mean' | None) pairs).
    shape: the source plane; new_shape: the plane after the processing function (and rotation)."""
    new_shape = np.array(new_shape)
    new_voxel_size = list(source_voxel)
    if is_rotated:
        new_voxel_size[1] *= shape[0] / new_shape[1]
        new_voxel_size[2] *= shape[1] / new_shape[0]
        new_voxel_size[1], new_voxel_size[2] = new_voxel_size[2], new_voxel_size[1]
    else:
        new_voxel_size[1] *= shape[0] / new_shape[0]
        new_voxel_size[2] *= shape[1] / new_shape[1]
    reduction_times = target_voxel / np.array(new_voxel_size[1:3])
    target_shape = tuple(int(v) for v in (new_shape / reduction_times).round().astype(int))
    reduction_factors = np.floor(np.sqrt(reduction_times)).astype(int)
    method_y = ["max" if i % 2 == 0 else "mean" for i in range(reduction_factors[0])]
    method_x = ["mean" if i % 2 == 0 else "max" for i in range(reduction_factors[1])]
    if reduction_factors[0] > reduction_factors[1]:
        method_x += [None] * (reduction_factors[0] - reduction_factors[1])
    elif reduction_factors[0] < reduction_factors[1]:
        method_y += [None] * (reduction_factors[1] - reduction_factors[0])
    methods = list(zip(method_y, method_x))
    if not alternating_downsampling_method:
        methods = [("mean", "mean") for _ in methods]
    return target_shape, methods


def reduced_shape(shape, target_shape, down_sampling_methods):
    """the shape the 2x1 / 1x2 reductions stop at (parallel_image_processor.py:377-381)."""
    r, c = int(shape[0]), int(shape[1])
    for y_method, x_method in down_sampling_methods:
        if y_method is not None and ceil(r / 2) >= target_shape[0]:
            r = ceil(r / 2)
        if x_method is not None and ceil(c / 2) >= target_shape[1]:
            c = ceil(c / 2)
    return r, c


def anti_aliasing_kernels(shape, target_shape):
    """skimage.transform.resize(anti_aliasing=True): sigma = max(0, (in / out - 1) / 2) per axis, then
    scipy.ndimage._filters._gaussian_kernel1d(sigma, 0, int(4 sigma + 0.5)) — numpy arithmetic, evaluated here as scipy
    does.  Returns [(radius, weights) | None] per axis."""
    factors = np.divide(shape, target_shape)
    sigmas = np.maximum(0, (factors - 1) / 2)
    out = []
    for sd in (float(sigmas[0]), float(sigmas[1])):
        radius = int(4.0 * sd + 0.5) if sd > 1e-15 else 0
        if radius <= 0:
            out.append(None)
            continue
        x = np.arange(-radius, radius + 1)
        phi = np.exp(-0.5 / (sd * sd) * x ** 2)
        out.append((radius, np.ascontiguousarray((phi / phi.sum())[::-1])))
    return out


def _device_array(x):
    import torch
    if _native._is_torch(x):
        if not x.is_cuda:
            raise TypeError("torch tensors must live on the GPU")
        return x.contiguous(), True
    return torch.from_numpy(np.ascontiguousarray(x)).cuda(), False


def down_sample_xy(img, target_shape: Tuple[int, int], down_sampling_methods: Sequence):
    """parallel_image_processor.py:371-385 for one plane (H, W) or a batch (Z, H, W): float32 planes of `target_shape`."""
    import torch
    t, was_torch = _device_array(img)
    squeeze = t.ndim == 2
    if squeeze:
        t = t[None]
    if t.dtype not in (torch.uint8, torch.uint16, torch.float32):
        raise TypeError(f"unsupported dtype {t.dtype}")
    code = {torch.uint8: _native.U8, torch.uint16: _native.U16, torch.float32: _native.F32}[t.dtype]
    methods = [(_name(a), _name(b)) for a, b in down_sampling_methods]
    target_shape = (int(target_shape[0]), int(target_shape[1]))
    n, rows, cols = t.shape
    pre = reduced_shape((rows, cols), target_shape, methods)
    for n_in, n_out in zip(pre, target_shape):       # the shape scipy.ndimage.zoom derives from skimage's factors
        if int(round(n_in * (1 / np.divide(n_in, n_out)))) != n_out:
            raise NotImplementedError(f"target {target_shape}: skimage's zoom factor rounds to another output shape")
    aa = anti_aliasing_kernels(pre, target_shape)
    steps = (C.c_int32 * max(1, 2 * len(methods)))(*[_METHOD[m] for pair in methods for m in pair])
    out = torch.empty((n,) + target_shape, dtype=torch.float32, device=t.device)
    ctx = _native.context(t.device.index or 0)
    wy = aa[0][1] if aa[0] else None
    wx = aa[1][1] if aa[1] else None
    ctx.check(_native.lib().b2s_isotropic_xy(
        ctx._h, C.c_void_p(t.data_ptr()), code, rows, cols, len(methods), steps, target_shape[0], target_shape[1],
        pre[0], pre[1], C.c_void_p(wy.ctypes.data if wy is not None else None), aa[0][0] if aa[0] else 0,
        C.c_void_p(wx.ctypes.data if wx is not None else None), aa[1][0] if aa[1] else 0,
        C.c_void_p(out.data_ptr()), n, C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)))
    if squeeze:
        out = out[0]
    return out if was_torch else out.cpu().numpy()


def is_uniform(x) -> bool:
    """is_uniform_2d / is_uniform_3d (core.py:106-121) of a device array."""
    import torch
    t, _ = _device_array(x)
    code = {torch.uint8: _native.U8, torch.uint16: _native.U16, torch.float32: _native.F32}[t.dtype]
    flag = C.c_int32(0)
    ctx = _native.context(t.device.index or 0)
    ctx.check(_native.lib().b2s_is_uniform(ctx._h, C.c_void_p(t.data_ptr()), code, t.numel(), C.byref(flag),
                                           C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)))
    return bool(flag.value)


def down_sample_z(z_stack, down_sampling_method_z: Sequence, down_sampled_dtype="float32", post_processed_dtype=None):
    """parallel_image_processor.py:412-433: the (Z, h, w) float32 stack of down-sampled planes -> one plane, reduced pair
    by pair along z with the alternating methods, then converted to `down_sampled_dtype`."""
    import torch
    t, was_torch = _device_array(z_stack)
    if t.dtype != torch.float32 or t.ndim != 3:
        raise TypeError("z_stack must be a (Z, h, w) float32 array")
    ctx = _native.context(t.device.index or 0)
    stream = C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)
    shape = tuple(t.shape[1:])
    if is_uniform(t):                                                     # :413-415
        img = torch.zeros(shape, dtype=torch.float32, device=t.device)
    else:
        for z_method in down_sampling_method_z:
            z_method = _name(z_method)
            if z_method is not None and t.shape[0] > 1:
                out = torch.empty((ceil(t.shape[0] / 2),) + shape, dtype=torch.float32, device=t.device)
                ctx.check(_native.lib().b2s_isotropic_z(ctx._h, C.c_void_p(t.data_ptr()), t.shape[0], shape[0] * shape[1],
                                                        _METHOD[z_method], C.c_void_p(out.data_ptr()), stream))
                t = out
        assert t.shape[0] == 1                                             # :420
        img = t[0]
        dt = np.dtype(down_sampled_dtype)
        if dt != np.float32:
            if dt == np.uint16:
                mode, shift, tdt = 1, 0, torch.uint16                     # convert_to_16bit_fun
            elif dt == np.uint8:
                if post_processed_dtype is not None and np.dtype(post_processed_dtype) == np.uint8:
                    mode, shift, tdt = 4, 0, torch.uint8                  # img.astype(uint8)
                else:
                    mode, shift, tdt = 2, 8, torch.uint8                  # convert_to_8bit_fun(img): default shift 8
            else:
                raise RuntimeError("requested downsampled format is not supported")
            out = torch.empty(shape, dtype=tdt, device=t.device)
            ctx.check(_native.lib().b2s_isotropic_convert(ctx._h, C.c_void_p(img.data_ptr()), img.numel(), mode, shift,
                                                          C.c_void_p(out.data_ptr()), stream))
            img = out
    return img if was_torch else img.cpu().numpy()
