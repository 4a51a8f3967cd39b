"""ORACLE — TEST INFRASTRUCTURE ONLY.

Loads the reference's own source files, UNMODIFIED, from /root/reference (this container only — the
GPU box has no /root/reference) so that the restated oracle (oracle/pystripe_oracle.py) can be pinned
against them and golden vectors can be generated (tests/golden/make_golden.py).

The reference cannot be imported as-is: pywt, ptwt, numexpr, skimage, tifffile,
imageio, dcimg are not installed.  We inject sys.modules stubs for the I/O / dead-GPU packages, the
restated `pywt` (oracle/pywt_shim.py), trivial restatements of skimage.measure.block_reduce, and force
the numpy fall-backs the reference itself contains (`core.USE_NUMEXPR = False`).

log1p / expm1 pin: the shipped reference evaluates them with numexpr -> libm.  With USE_NUMEXPR=False the
same functions call numpy, which is libm too UNLESS numpy dispatches AVX512 SVML.  `load()` therefore
refuses to run unless NPY_DISABLE_CPU_FEATURES disabled AVX512F before numpy was imported (use
`python -m oracle.ref_runner --selftest` or the make_golden script, both re-exec with the variable set).
"""
import importlib.util
import os
import sys
import types
from pathlib import Path

REFERENCE_ROOT = Path("/root/reference")
NPY_PIN = "AVX512F AVX512CD AVX512_SKX AVX512_CLX AVX512_CNL AVX512_ICL AVX512_SPR"


def available() -> bool:
    return (REFERENCE_ROOT / "pystripe" / "core.py").exists()


def numpy_is_pinned_to_libm() -> bool:
    import numpy as np
    import ctypes
    libm = ctypes.CDLL("libm.so.6")
    libm.log1pf.restype = ctypes.c_float
    libm.log1pf.argtypes = [ctypes.c_float]
    x = np.arange(1, 4097, dtype=np.float32) * np.float32(7.3)
    a = np.log1p(x)
    b = np.array([libm.log1pf(float(v)) for v in x], dtype=np.float32)
    return bool((a == b).all())


def ensure_pinned_env():
    """re-exec the current interpreter with numpy's AVX512 dispatch disabled (must precede `import numpy`)."""
    if os.environ.get("NPY_DISABLE_CPU_FEATURES") is None:
        env = dict(os.environ, NPY_DISABLE_CPU_FEATURES=NPY_PIN)
        os.execve(sys.executable, list(sys.orig_argv), env)


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    return m


def _not_available(what):
    def f(*a, **k):
        raise RuntimeError(f"{what} is not available in the oracle harness")
    return f


def _resize(image, output_shape, **kw):
    """skimage.transform.resize restated (oracle.pystripe_oracle.skimage_resize): skimage is absent here."""
    from oracle.pystripe_oracle import skimage_resize
    return skimage_resize(image, output_shape, **kw)


def _block_reduce(image, block_size=2, func=None, cval=0, func_kwargs=None):
    """skimage.measure.block_reduce restated: pad trailing edges with cval up to a block multiple,
    view as blocks, reduce over the block axes (skimage/measure/block.py)."""
    import numpy as np
    if np.isscalar(block_size):
        block_size = (block_size,) * image.ndim
    pad = [(0, (-s) % b) for s, b in zip(image.shape, block_size)]
    if any(p[1] for p in pad):
        image = np.pad(image, pad, mode="constant", constant_values=cval)
    shp = []
    for s, b in zip(image.shape, block_size):
        shp += [s // b, b]
    v = image.reshape(shp)
    red_axes = tuple(range(1, 2 * image.ndim, 2))
    return func(v, axis=red_axes)


_CACHE = {}


def load():
    """returns (core_module, lightsheet_module) of the reference, executed verbatim."""
    if "core" in _CACHE:
        return _CACHE["core"], _CACHE["ls"]
    if not available():
        raise RuntimeError("/root/reference is not present on this machine")
    if not numpy_is_pinned_to_libm():
        raise RuntimeError("numpy log1p is not libm's here: set NPY_DISABLE_CPU_FEATURES=" + repr(NPY_PIN))
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    from oracle import pywt_shim

    saved = {k: v for k, v in sys.modules.items() if k == "pystripe" or k.startswith("pystripe.") or
             k == "supplements" or k.startswith("supplements.")}
    for k in saved:
        del sys.modules[k]
    injected = {
        "dcimg": _stub("dcimg", DCIMGFile=_not_available("dcimg")),
        "imageio": _stub("imageio"),
        "imageio.v3": _stub("imageio.v3", imread=_not_available("imageio")),
        "numexpr": _stub("numexpr", evaluate=_not_available("numexpr")),
        "ptwt": _stub("ptwt", wavedec2=_not_available("ptwt"), waverec2=_not_available("ptwt")),
        "pywt": pywt_shim,
        "skimage": _stub("skimage"),
        "skimage.filters": _stub("skimage.filters", threshold_otsu=_not_available("skimage"),
                                 threshold_multiotsu=_not_available("skimage")),
        "skimage.measure": _stub("skimage.measure", block_reduce=_block_reduce),
        "skimage.transform": _stub("skimage.transform", resize=_resize),
        "tifffile": _stub("tifffile", imwrite=_not_available("tifffile")),
        "tifffile.tifffile": _stub("tifffile.tifffile", TiffFileError=type("TiffFileError", (Exception,), {})),
    }
    prev = {k: sys.modules.get(k) for k in injected}
    sys.modules.update(injected)
    # namespace packages pointing INTO /root/reference (nothing is copied)
    pkg = types.ModuleType("pystripe")
    pkg.__path__ = [str(REFERENCE_ROOT / "pystripe")]
    sup = types.ModuleType("supplements")
    sup.__path__ = [str(REFERENCE_ROOT / "supplements")]
    sys.modules["pystripe"] = pkg
    sys.modules["supplements"] = sup
    try:
        def _load(modname, path):
            spec = importlib.util.spec_from_file_location(modname, path)
            mod = importlib.util.module_from_spec(spec)
            sys.modules[modname] = mod
            spec.loader.exec_module(mod)
            return mod
        ls = _load("pystripe.lightsheet_correct", REFERENCE_ROOT / "pystripe" / "lightsheet_correct.py")
        core = _load("pystripe.core", REFERENCE_ROOT / "pystripe" / "core.py")
        core.USE_NUMEXPR = False      # numexpr absent -> the numpy branches of the same functions
    finally:
        for k in list(sys.modules):
            if k == "pystripe" or k.startswith("pystripe.") or k == "supplements" or k.startswith("supplements."):
                del sys.modules[k]
        sys.modules.update(saved)
        for k, v in prev.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    _CACHE["core"], _CACHE["ls"] = core, ls
    return core, ls


if __name__ == "__main__":
    ensure_pinned_env()
    import numpy as np
    core, ls = load()
    rng = np.random.default_rng(0)
    img = rng.integers(90, 700, size=(128, 160)).astype(np.uint16)
    out = core.filter_streaks(img.copy(), sigma=(64, 64), wavelet="db4")
    print("reference filter_streaks ran verbatim:", out.dtype, out.shape, int(out.min()), int(out.max()))
    print("calculate_pad_size((2048,2048),256) =", core.calculate_pad_size((2048, 2048), 256))
