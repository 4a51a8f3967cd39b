"""ctypes binding of libb200stripe.so (include/b200stripe.h).  No CPU fallback: if the library is missing or no
B200 is visible, every entry point raises.
"""
import ctypes as C
import os
import threading
import weakref
from pathlib import Path

import numpy as np

_PKG = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("B200STRIPE_LIB", _PKG.parent / "lib" / "libb200stripe.so"))

U8, U16, F32 = 0, 1, 2
PAD_MODES = {"reflect": 0, "wrap": 1, "symmetric": 2, "edge": 3, "constant": 4, "linear_ramp": 5, "maximum": 6, "mean": 7,
             "median": 8, "minimum": 9, "empty": 10}
DS_METHODS = {"max": 0, "min": 1, "mean": 2, "median": 3}
STAGE = {"all": 0, "prologue": 1, "forward": 2, "notch": 3, "inverse": 4}
N_KERNEL_CLASSES = 8
TIMING_LEVELS = 33
KERNEL_CLASSES = ("pre", "prologue", "dwt_fwd", "notch", "dwt_inv", "epilogue", "lightsheet", "other")
OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA, ERR_NOMEM = 0, -1, -2, -3, -4


class Params(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("height", C.c_int32), ("width", C.c_int32),
        ("in_dtype", C.c_int32), ("out_dtype", C.c_int32),
        ("sigma1", C.c_double), ("sigma2", C.c_double),
        ("threshold_nonpositive", C.c_int32), ("level", C.c_int32), ("n_taps", C.c_int32),
        ("dec_lo", C.POINTER(C.c_double)),
        ("pad_mode", C.c_int32), ("bidirectional", C.c_int32), ("log1p", C.c_int32),
        ("process_img", C.c_int32), ("has_flat", C.c_int32), ("gaussian", C.c_int32),
        ("down_sample_y", C.c_int32), ("down_sample_x", C.c_int32), ("down_sample_method", C.c_int32),
        ("dark", C.c_double),
        ("lightsheet", C.c_int32), ("artifact_length", C.c_int32), ("background_window_size", C.c_int32),
        ("percentile", C.c_double), ("lightsheet_vs_background", C.c_double),
        ("convert_to_16bit", C.c_int32), ("convert_to_8bit", C.c_int32), ("bit_shift_to_right", C.c_int32),
        ("rotate", C.c_int32), ("flip_upside_down", C.c_int32), ("reference_quirks", C.c_int32),
        ("new_height", C.c_int32), ("new_width", C.c_int32),
        ("bleach", C.c_int32), ("bleach_per_plane", C.c_int32),
        ("bleach_b0", C.c_double), ("bleach_b1", C.c_double), ("bleach_a1", C.c_double), ("bleach_zi", C.c_double),
        ("bleach_clip_min", C.c_double), ("bleach_clip_med", C.c_double), ("bleach_clip_max", C.c_double),
        ("pad_constant", C.c_double),
        ("aa_radius_y", C.c_int32), ("aa_radius_x", C.c_int32),
        ("mask", C.c_int32), ("mask_close", C.c_int32), ("mask_open", C.c_int32), ("mask_per_plane", C.c_int32),
        ("mask_threshold", C.c_double),
        ("max_batch", C.c_int32), ("debug_stop_after", C.c_int32), ("exact", C.c_int32),
    ]


class PlanInfo(C.Structure):
    _fields_ = [
        ("out_height", C.c_int32), ("out_width", C.c_int32), ("out_dtype", C.c_int32), ("n_passes", C.c_int32),
        ("work_height", C.c_int32), ("work_width", C.c_int32),
        ("base_pad", C.c_int32), ("pad_y", C.c_int32), ("pad_x", C.c_int32),
        ("padded_height", C.c_int32), ("padded_width", C.c_int32), ("levels", C.c_int32),
        ("level_rows", C.c_int32 * 32), ("level_cols", C.c_int32 * 32),
        ("workspace_bytes", C.c_int64), ("algorithmic_bytes_per_plane", C.c_int64), ("flops_per_plane", C.c_int64),
    ]


EXPORTS = (
    "b2s_version", "b2s_params_default", "b2s_create", "b2s_destroy", "b2s_last_error", "b2s_device_sm_count",
    "b2s_plan_create", "b2s_plan_destroy", "b2s_plan_query", "b2s_plan_geometry", "b2s_resize_table", "b2s_plan_set_flat", "b2s_plan_set_notch", "b2s_plan_wants_notch_matrix", "b2s_plan_set_notch_matrix", "b2s_plan_set_bleach_levels", "b2s_plan_set_mask_thresholds", "b2s_plan_set_aa_weights", "b2s_run",
    "b2s_host_alloc", "b2s_host_free", "b2s_launch_count", "b2s_timing_enable", "b2s_timing_read",
    "b2s_debug_read", "b2s_debug_math", "b2s_debug_expm1_table_check",
    "b2s_resize_aa", "b2s_isotropic_xy", "b2s_isotropic_z", "b2s_isotropic_convert", "b2s_is_uniform", "b2s_histogram", "b2s_img_mask", "b2s_deflate_bound", "b2s_deflate_strips",
)

_lib = None
_lock = threading.Lock()


class B200StripeError(RuntimeError):
    pass


def lib():
    """load the shared library (fails loudly when it has not been built: `python __graft_entry__.py build`)."""
    global _lib
    with _lock:
        if _lib is None:
            if not LIB_PATH.exists():
                raise B200StripeError(
                    f"{LIB_PATH} not found — build it with `make -C image-preprocessing-pipeline_b200/csrc` "
                    "(there is no CPU fallback)")
            L = C.CDLL(str(LIB_PATH))
            vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
            L.b2s_version.restype = i32
            L.b2s_params_default.argtypes = [C.POINTER(Params)]
            L.b2s_params_default.restype = None
            L.b2s_create.argtypes = [i32, C.POINTER(vp)]
            L.b2s_destroy.argtypes = [vp]
            L.b2s_destroy.restype = None
            L.b2s_last_error.argtypes = [vp]
            L.b2s_last_error.restype = C.c_char_p
            L.b2s_device_sm_count.argtypes = [vp]
            L.b2s_plan_create.argtypes = [vp, C.POINTER(Params), C.POINTER(vp)]
            L.b2s_plan_destroy.argtypes = [vp]
            L.b2s_plan_destroy.restype = None
            L.b2s_plan_query.argtypes = [vp, C.POINTER(PlanInfo)]
            L.b2s_plan_geometry.argtypes = [C.POINTER(Params), C.POINTER(PlanInfo), C.c_char_p, C.c_size_t]
            L.b2s_plan_set_flat.argtypes = [vp, vp, i32]
            L.b2s_plan_set_notch.argtypes = [vp, i32, i32, i32, vp, i32]
            L.b2s_plan_set_aa_weights.argtypes = [vp, i32, vp, i32]
            L.b2s_plan_set_bleach_levels.argtypes = [vp, vp, vp, i64]
            L.b2s_plan_set_mask_thresholds.argtypes = [vp, vp, i64]
            L.b2s_plan_wants_notch_matrix.argtypes = [vp]
            L.b2s_plan_set_notch_matrix.argtypes = [vp, i32, i32, i32, vp, i32]
            L.b2s_isotropic_xy.argtypes = [vp, vp, i32, i32, i32, i32, vp, i32, i32, i32, i32, vp, i32, vp, i32, vp, i32, vp]
            L.b2s_resize_aa.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp, i32, vp, i32, vp, i32, vp]
            L.b2s_isotropic_z.argtypes = [vp, vp, i32, i64, i32, vp, vp]
            L.b2s_isotropic_convert.argtypes = [vp, vp, i64, i32, i32, vp, vp]
            L.b2s_deflate_bound.argtypes = [i32, i32, i32, i32, i32]
            L.b2s_deflate_bound.restype = i64
            L.b2s_deflate_strips.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp, i64, vp, vp, C.POINTER(i64), vp]
            L.b2s_img_mask.argtypes = [vp, vp, i32, i32, i32, i32, C.c_double, i32, i32, vp, vp]
            L.b2s_histogram.argtypes = [vp, vp, i32, i32, i64, i32, vp, i32, i32, vp]
            L.b2s_is_uniform.argtypes = [vp, vp, i32, i64, vp, vp]
            L.b2s_run.argtypes = [vp, vp, vp, i64, i32, i32, vp]
            L.b2s_host_alloc.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
            L.b2s_host_free.argtypes = [vp, vp]
            L.b2s_launch_count.argtypes = [vp]
            L.b2s_launch_count.restype = i64
            L.b2s_timing_enable.argtypes = [vp, i32]
            L.b2s_timing_read.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(i64), i32]
            L.b2s_debug_read.argtypes = [vp, i32, i32, i32, vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
            L.b2s_debug_math.argtypes = [vp, i32, vp, vp, i64]
            L.b2s_debug_expm1_table_check.argtypes = [vp, i32, C.c_double, i32, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64)]
            _lib = L
    return _lib


def _raise(code, msg):
    msg = msg.decode() if isinstance(msg, bytes) else str(msg)
    if code == ERR_INVALID:
        # the reference raises ValueError (np_notch, pywt) or RuntimeError (shift / modes) for these
        if "right shift" in msg or "padding mode" in msg or "down-sampling" in msg:
            raise RuntimeError(msg)
        raise ValueError(msg)
    if code == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    if code == ERR_NOMEM:
        raise MemoryError(msg)
    raise B200StripeError(msg)


def np_dtype_code(dt) -> int:
    dt = np.dtype(dt)
    if dt == np.uint8:
        return U8
    if dt == np.uint16:
        return U16
    if dt == np.float32:
        return F32
    raise TypeError(f"unsupported dtype {dt} (the GPU path takes uint8, uint16 or float32 planes)")


CODE_TO_NP = {U8: np.uint8, U16: np.uint16, F32: np.float32}


def default_params() -> Params:
    p = Params()
    lib().b2s_params_default(C.byref(p))
    return p


def plan_geometry(p: Params) -> PlanInfo:
    """host-only geometry (works without a GPU)."""
    info = PlanInfo()
    err = C.create_string_buffer(512)
    rc = lib().b2s_plan_geometry(C.byref(p), C.byref(info), err, 512)
    if rc:
        _raise(rc, err.value)
    return info


class Context:
    """one per GPU (b2s_create)."""

    def __init__(self, device: int = 0):
        self.device = int(device)
        self._h = C.c_void_p()
        self._pinned = {}
        self._pool_lock = threading.Lock()
        self._pool_free = {}
        self._pool_bytes = 0
        rc = lib().b2s_create(self.device, C.byref(self._h))
        if rc:
            msg = lib().b2s_last_error(self._h) if self._h else b"b2s_create failed"
            if self._h:
                lib().b2s_destroy(self._h)
                self._h = None
            _raise(rc, msg)

    def check(self, rc):
        if rc:
            _raise(rc, lib().b2s_last_error(self._h))

    @property
    def sm_count(self):
        return lib().b2s_device_sm_count(self._h)

    @property
    def launch_count(self):
        return lib().b2s_launch_count(self._h)

    def timing_enable(self, on=True):
        self.check(lib().b2s_timing_enable(self._h, int(bool(on))))

    def timing_read(self, reset=True, per_level=False):
        """{class: (ms, launches)}; per_level=True -> {(class, level): (ms, launches)} (level 0 = not level-specific)."""
        ms = (C.c_double * (N_KERNEL_CLASSES * TIMING_LEVELS))()
        n = (C.c_int64 * (N_KERNEL_CLASSES * TIMING_LEVELS))()
        self.check(lib().b2s_timing_read(self._h, ms, n, int(reset)))
        if per_level:
            return {(k, l): (ms[i * TIMING_LEVELS + l], n[i * TIMING_LEVELS + l])
                    for i, k in enumerate(KERNEL_CLASSES) for l in range(TIMING_LEVELS) if n[i * TIMING_LEVELS + l]}
        return {k: (sum(ms[i * TIMING_LEVELS:(i + 1) * TIMING_LEVELS]), sum(n[i * TIMING_LEVELS:(i + 1) * TIMING_LEVELS]))
                for i, k in enumerate(KERNEL_CLASSES)}

    def pinned_empty(self, shape, dtype):
        """numpy array over page-locked host memory (b2s_host_alloc).  Owned by the context: released by
        pinned_free(arr) or when the context closes."""
        dtype = np.dtype(dtype)
        count = int(np.prod(shape))
        nbytes = max(count * dtype.itemsize, 16)
        ptr = C.c_void_p()
        self.check(lib().b2s_host_alloc(self._h, nbytes, C.byref(ptr)))
        buf = (C.c_char * nbytes).from_address(ptr.value)
        arr = np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)
        self._pinned[arr.ctypes.data] = ptr
        return arr

    def pinned_free(self, arr):
        ptr = self._pinned.pop(arr.ctypes.data, None)
        if ptr is not None and self._h:
            lib().b2s_host_free(self._h, ptr)

    # -- recycled page-locked result buffers: run_host() writes results straight into them (the device-to-host copy lands
    #    in the array the caller receives: no staging copy, no first-touch page faults on a fresh pageable array).  A block
    #    returns to the free list when the last view of the array is garbage-collected.
    POOL_LIMIT = int(float(os.environ.get("B200STRIPE_PINNED_POOL_GB", "16")) * 2 ** 30)
    POOL_MIN = 1 << 20

    def pooled_empty(self, shape, dtype):
        """array for a result: pinned and recycled when large enough and the pool has room, else plain numpy.empty."""
        dtype = np.dtype(dtype)
        count = int(np.prod(shape))
        nbytes = count * dtype.itemsize
        if nbytes < self.POOL_MIN or not self._h:
            return np.empty(shape, dtype)
        size = -(-nbytes // (1 << 21)) << 21                      # 2 MiB classes
        with self._pool_lock:
            free = self._pool_free.setdefault(size, [])
            addr = free.pop() if free else None
            if addr is None:
                if self._pool_bytes + size > self.POOL_LIMIT:
                    # give idle blocks of other sizes back before giving up on pinned results
                    for sz, lst in list(self._pool_free.items()):
                        while lst and self._pool_bytes + size > self.POOL_LIMIT:
                            lib().b2s_host_free(self._h, C.c_void_p(lst.pop()))
                            self._pool_bytes -= sz
                if self._pool_bytes + size > self.POOL_LIMIT:
                    return np.empty(shape, dtype)
                ptr = C.c_void_p()
                if lib().b2s_host_alloc(self._h, size, C.byref(ptr)):
                    return np.empty(shape, dtype)
                addr = ptr.value
                self._pool_bytes += size
        buf = (C.c_char * size).from_address(addr)
        base = np.frombuffer(buf, dtype=np.uint8, count=size)
        weakref.finalize(base, self._pool_release, size, addr)
        return base[:nbytes].view(dtype).reshape(shape)

    def _pool_release(self, size, addr):
        with self._pool_lock:
            self._pool_free.setdefault(size, []).append(addr)

    def debug_math(self, which: int, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.empty_like(x)
        self.check(lib().b2s_debug_math(self._h, which, x.ctypes.data, out.ctypes.data, x.size))
        return out

    def close(self):
        if self._h:
            for ptr in list(self._pinned.values()):
                lib().b2s_host_free(self._h, ptr)
            self._pinned.clear()
            with self._pool_lock:                       # blocks still held by live result arrays are left alone
                for lst in self._pool_free.values():
                    for addr in lst:
                        lib().b2s_host_free(self._h, C.c_void_p(addr))
                self._pool_free.clear()
            lib().b2s_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_contexts = {}
_contexts_lock = threading.Lock()


def context(device: int = 0) -> Context:
    with _contexts_lock:
        if device not in _contexts:
            _contexts[device] = Context(device)
        return _contexts[device]


class Plan:
    """b2s_plan wrapper: owns tables + workspace for one (shape, dtype, parameter set)."""

    def __init__(self, ctx: Context, params: Params, dec_lo=None, flat=None):
        self.ctx = ctx
        self._keep = None
        self.lock = threading.RLock()      # one b2s_run per plan at a time (its workspace is the plan's)
        self._users = 0
        self._last_stream = None
        if dec_lo is not None:
            arr = (C.c_double * len(dec_lo))(*[float(v) for v in dec_lo])
            params.n_taps = len(dec_lo)
            params.dec_lo = C.cast(arr, C.POINTER(C.c_double))
            self._keep = arr
        self.params = params
        self._h = C.c_void_p()
        ctx.check(lib().b2s_plan_create(ctx._h, C.byref(params), C.byref(self._h)))
        self.info = PlanInfo()
        ctx.check(lib().b2s_plan_query(self._h, C.byref(self.info)))
        if flat is not None:
            self.set_flat(flat)

    # -- properties
    @property
    def out_shape(self):
        return (self.info.out_height, self.info.out_width)

    @property
    def out_dtype(self):
        return CODE_TO_NP[self.info.out_dtype]

    @property
    def in_shape(self):
        return (self.params.height, self.params.width)

    @property
    def in_dtype(self):
        return CODE_TO_NP[self.params.in_dtype]

    def set_flat(self, flat):
        if _is_torch(flat):
            f = flat.contiguous().float()
            self.ctx.check(lib().b2s_plan_set_flat(self._h, f.data_ptr(), int(f.is_cuda)))
        else:
            f = np.ascontiguousarray(flat, dtype=np.float32)
            self.ctx.check(lib().b2s_plan_set_flat(self._h, f.ctypes.data, 0))

    def set_notch(self, pass_idx: int, level: int, axis: int, g: np.ndarray):
        """upload a host-evaluated np_notch table (float32) for (pass, 1-based level, axis 0 = cH / 1 = cV)."""
        g = np.ascontiguousarray(g, dtype=np.float32)
        self.ctx.check(lib().b2s_plan_set_notch(self._h, pass_idx, level, axis, C.c_void_p(g.ctypes.data), int(g.size)))

    @property
    def wants_notch_matrix(self) -> bool:
        """the plan runs the destripe in float64 (integer pixels without log1p) and takes dense notch matrices."""
        return bool(lib().b2s_plan_wants_notch_matrix(self._h))

    def set_notch_matrix(self, pass_idx: int, level: int, axis: int, R: np.ndarray):
        R = np.ascontiguousarray(R, dtype=np.float64)
        self.ctx.check(lib().b2s_plan_set_notch_matrix(self._h, pass_idx, level, axis, C.c_void_p(R.ctypes.data), int(R.shape[0])))

    def set_bleach_levels(self, clip: np.ndarray, pad_value: np.ndarray = None):
        """per-plane (clip_min, clip_med, clip_max) [n, 3] float64 and optional constant-padding values [n] float32 for the
        next run of a plan created with bleach_per_plane."""
        clip = np.ascontiguousarray(clip, dtype=np.float64).reshape(-1, 3)
        pv = None if pad_value is None else np.ascontiguousarray(pad_value, dtype=np.float32)
        self.ctx.check(lib().b2s_plan_set_bleach_levels(self._h, C.c_void_p(clip.ctypes.data),
                                                        C.c_void_p(pv.ctypes.data) if pv is not None else None, clip.shape[0]))

    def set_mask_thresholds(self, thr: np.ndarray):
        """per-plane get_img_mask thresholds [n] float64 for the next run of a plan created with mask_per_plane."""
        thr = np.ascontiguousarray(thr, dtype=np.float64).reshape(-1)
        self.ctx.check(lib().b2s_plan_set_mask_thresholds(self._h, C.c_void_p(thr.ctypes.data), thr.shape[0]))

    def set_aa_weights(self, axis: int, w: np.ndarray):
        """upload the anti-aliasing Gaussian of skimage.transform.resize along one axis (2 * radius + 1 float64 weights)."""
        w = np.ascontiguousarray(w, dtype=np.float64)
        self.ctx.check(lib().b2s_plan_set_aa_weights(self._h, axis, C.c_void_p(w.ctypes.data), int(w.size)))

    def run_host(self, src: np.ndarray, dst: np.ndarray = None) -> np.ndarray:
        """src: (n, H, W) or (H, W) numpy array in host memory (pinned or pageable)."""
        single = src.ndim == 2
        s = src[None] if single else src
        if s.shape[1:] != self.in_shape or s.dtype != self.in_dtype:
            raise ValueError(f"plan expects planes {self.in_shape} {np.dtype(self.in_dtype)}, got {s.shape[1:]} {s.dtype}")
        s = np.ascontiguousarray(s)
        n = s.shape[0]
        if dst is None:
            dst = self.ctx.pooled_empty((n,) + self.out_shape, self.out_dtype)
        d = dst[None] if dst.ndim == 2 else dst
        if not d.flags.c_contiguous or d.shape != (n,) + self.out_shape or d.dtype != self.out_dtype:
            raise ValueError("dst must be a C-contiguous array of the plan's output shape and dtype")
        with self.lock:
            if not self._h:
                raise B200StripeError("plan was closed")
            self.ctx.check(lib().b2s_run(self._h, s.ctypes.data, d.ctypes.data, n, 0, 0, None))
        return d[0] if single else d

    def run_device(self, src_ptr: int, dst_ptr: int, n: int, stream: int = 0):
        """raw device pointers (zero-copy torch tensors); asynchronous on `stream`."""
        with self.lock:
            if not self._h:
                raise B200StripeError("plan was closed")
            if self._last_stream is not None and self._last_stream != stream:
                # the workspace is still in flight on another stream: order the two users
                import torch
                torch.cuda.synchronize(self.ctx.device)
            self._last_stream = stream
            self.ctx.check(lib().b2s_run(self._h, C.c_void_p(src_ptr), C.c_void_p(dst_ptr), n, 1, 1,
                                         C.c_void_p(stream) if stream else None))

    def run_torch(self, src, dst=None):
        """src: CUDA torch tensor (n, H, W) / (H, W); runs on torch's current stream, no copies."""
        import torch
        single = src.dim() == 2
        s = (src[None] if single else src).contiguous()
        if tuple(s.shape[1:]) != self.in_shape:
            raise ValueError(f"plan expects planes {self.in_shape}, got {tuple(s.shape[1:])}")
        if dst is None:
            dst = torch.empty((s.shape[0],) + self.out_shape, dtype=_np_to_torch(self.out_dtype), device=s.device)
        stream = torch.cuda.current_stream(s.device).cuda_stream
        self.run_device(s.data_ptr(), dst.data_ptr(), s.shape[0], stream)
        return dst[0] if single else dst

    def debug_read(self, what: int, level: int = 0, plane: int = 0) -> np.ndarray:
        if what == 0:
            shape = (self.info.padded_height, self.info.padded_width)
        else:
            shape = (self.info.level_rows[level - 1], self.info.level_cols[level - 1])
        out = np.empty(shape, dtype=np.float32)
        r, c = C.c_int32(), C.c_int32()
        self.ctx.check(lib().b2s_debug_read(self._h, what, level, plane, out.ctypes.data, C.byref(r), C.byref(c)))
        assert (r.value, c.value) == shape
        return out

    def close(self):
        with self.lock:
            if self._h:
                lib().b2s_plan_destroy(self._h)
                self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def _np_to_torch(dt):
    import torch
    return {np.uint8: torch.uint8, np.uint16: torch.uint16, np.float32: torch.float32}[dt]
