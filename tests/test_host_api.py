"""CPU: the C-ABI library loads and exports what include/b200stripe.h declares; host-only geometry; the drop-in
module surface; failure (not fallback) without a GPU."""
import ctypes as C
import inspect
import re
from pathlib import Path

import numpy as np
import pytest

from oracle import pystripe_oracle as orc
from oracle import pywt_shim as pw

ROOT = Path(__file__).resolve().parents[1]


def _lib():
    from pystripe import _native
    return _native


def test_library_exports_every_declared_symbol():
    nat = _lib()
    header = (ROOT / "include" / "b200stripe.h").read_text()
    declared = set(re.findall(r"\b(b2s_[a-z0-9_]+)\s*\(", header))
    assert declared == set(nat.EXPORTS), declared ^ set(nat.EXPORTS)
    L = nat.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert L.b2s_version() == 100


def test_params_struct_size_matches_c():
    nat = _lib()
    p = nat.default_params()
    assert p.struct_size == C.sizeof(nat.Params)
    assert p.pad_mode == nat.PAD_MODES["wrap"] and p.log1p == 1 and p.exact == 1


def _geom(shape, sigma, wavelet, level=0, mode="wrap"):
    nat = _lib()
    from pystripe import core
    p = nat.default_params()
    p.height, p.width = shape
    p.sigma1, p.sigma2 = sigma
    p.level = level
    p.pad_mode = nat.PAD_MODES[mode]
    taps = core._dec_lo(wavelet)
    arr = (C.c_double * len(taps))(*taps)
    p.n_taps = len(taps)
    p.dec_lo = C.cast(arr, C.POINTER(C.c_double))
    return nat.plan_geometry(p)


@pytest.mark.parametrize("shape,sigma,wavelet", [
    ((2048, 2048), (256, 256), "db10"), ((2048, 2048), (250, 250), "db9"), ((2000, 2000), (100, 100), "db9"),
    ((1600, 2000), (100, 100), "db9"), ((2047, 2049), (128, 256), "db4"), ((30, 30), (2, 2), "db2"),
    ((64, 64), (1, 1), "db9"), ((96, 128), (24, 24), "db10"), ((1024, 1024), (256, 256), "db10"),
])
def test_geometry_matches_oracle(shape, sigma, wavelet):
    info = _geom(shape, sigma, wavelet)
    base, py, px = orc.padded_geometry(shape, sigma, "wrap")
    assert (info.base_pad, info.pad_y, info.pad_x) == (base, py, px)
    PH, PW = shape[0] + 2 * base + py, shape[1] + 2 * base + px
    assert (info.padded_height, info.padded_width) == (PH, PW)
    coeffs = pw.wavedec2(np.zeros((PH, PW), np.float32), wavelet)
    assert info.levels == len(coeffs) - 1
    shapes = [c[0].shape for c in coeffs[1:]][::-1]
    assert [(info.level_rows[i], info.level_cols[i]) for i in range(info.levels)] == shapes
    assert info.n_passes == (1 if sigma[0] == sigma[1] else 2)


def test_headline_geometry_and_byte_model():
    info = _geom((2048, 2048), (256, 256), "db10")
    assert info.base_pad == 300 and info.padded_height == 2648 and info.levels == 7
    assert [info.level_rows[i] for i in range(7)] == [1333, 676, 347, 183, 101, 60, 39]
    assert abs(info.algorithmic_bytes_per_plane / 1e6 - 233.0) < 0.5      # SURVEY.md §8(d) stage model
    assert abs(info.flops_per_plane / 2 / 1e6 - 761) < 10                 # MACs


def test_invalid_parameters_map_to_reference_exceptions():
    with pytest.raises(ValueError):          # np_notch: sigma must be positive (core.py:657)
        _geom((64, 64), (0, 8), "db2")
    from pystripe import core
    with pytest.raises(ValueError):          # unknown wavelet, like pywt
        core._dec_lo("nope7")
    with pytest.raises(RuntimeError):        # core.py:1097-1099
        core._get_plan(0, (64, 64), 1, process=0, sigma=(8, 8), level=0, wavelet="db2", threshold=None,
                       padding_mode="bogus", bidirectional=False, log1p=True)


def test_drop_in_signatures():
    """keyword names and defaults of the reference."""
    import pystripe
    from pystripe import core
    fs = inspect.signature(core.filter_streaks).parameters
    assert list(fs)[:8] == ["img", "sigma", "level", "wavelet", "crossover", "threshold", "padding_mode", "bidirectional"]
    assert fs["sigma"].default == (250, 250) and fs["wavelet"].default == "db9" and fs["padding_mode"].default == "wrap"
    assert fs["log1p_normalization_needed"].default is True
    pi = inspect.signature(core.process_img).parameters
    for k, d in dict(flat=None, gaussian_filter_2d=False, down_sample=None, down_sample_method="max", sigma=(0, 0),
                     wavelet="coif15", padding_mode="wrap", dark=0, lightsheet=False, artifact_length=150,
                     background_window_size=200, percentile=0.25, lightsheet_vs_background=2.0, rotate=0,
                     flip_upside_down=False, convert_to_16bit=False, convert_to_8bit=False, bit_shift_to_right=8,
                     d_type=None).items():
        assert pi[k].default == d, k
    bf = inspect.signature(core.batch_filter).parameters
    assert list(bf)[:3] == ["input_path", "output_path", "files_list"]
    assert bf["padding_mode"].default == "reflect" and bf["wavelet"].default == "db9"
    assert bf["compression"].default == ("ADOBE_DEFLATE", 1) and bf["threads_per_gpu"].default == 8
    rfs = inspect.signature(core.read_filter_save).parameters
    assert rfs["convert_to_8bit"].default is True and rfs["padding_mode"].default == "reflect"
    for name in ("batch_filter imread_tif_raw_png imsave_tif MultiProcessQueueRunner progress_manager process_img "
                 "convert_to_8bit_fun log1p_jit prctl np_max np_mean is_uniform_2d calculate_pad_size "
                 "cuda_get_device_properties cuda_device_count CUDA_IS_AVAILABLE_FOR_PT USE_PYTORCH USE_JAX "
                 "is_uniform_3d convert_to_16bit_fun cuda_is_available_for_pt glob_re").split():
        assert hasattr(core, name), name
    for name in ("filter_streaks batch_filter np_gaussian_filter hist_match max_level foreground_fraction "
                 "imread_tif_raw_png imread_dcimg imsave_tif normalize_flat").split():
        assert hasattr(pystripe, name), name


def test_host_helpers_match_oracle():
    from pystripe import core
    for shape, s in [((2048, 2048), 256), ((2000, 2000), 100), ((1600, 2000), 512)]:
        assert core.calculate_pad_size(shape, s) == orc.calculate_pad_size(shape, s)
    rng = np.random.default_rng(0)
    a = rng.integers(0, 65536, (40, 50)).astype(np.uint16)
    for b in (0, 3, 8):
        assert np.array_equal(core.convert_to_8bit_fun(a.copy(), b), orc.convert_to_8bit_fun(a.copy(), b))
    with pytest.raises(RuntimeError):
        core.convert_to_8bit_fun(a.copy(), 9)
    assert core.max_level(2648, "db10") == 7 and core.max_level(3252, "db20") == pw.dwt_max_level(3252, 40)
    assert core.is_uniform_2d(np.full((4, 4), 3)) and not core.is_uniform_2d(a)
    f = rng.uniform(1, 9, (8, 8))
    assert np.array_equal(core.normalize_flat(f), orc.normalize_flat(f))
    assert core.calculate_down_sampled_size((2048, 2047), (2, 2)) == [1024, 1024]


def test_batch_filter_argument_errors(tmp_path):
    from pystripe import core
    with pytest.raises(TypeError):
        core.batch_filter(tmp_path, tmp_path / "o", convert_to_16bit=True, convert_to_8bit=True)
    with pytest.raises(TypeError):
        core.batch_filter(tmp_path, tmp_path / "o", flat=3.0)
    with pytest.raises(AssertionError):
        core.batch_filter(tmp_path / "missing", tmp_path / "o")
    assert core.batch_filter(tmp_path, tmp_path / "o2", sigma=(8, 8)) == 0   # nothing to do


def test_raw_roundtrip(tmp_path):
    from pystripe import raw
    img = np.random.default_rng(1).integers(0, 65536, (33, 47)).astype(np.uint16)
    raw.raw_imsave(tmp_path / "t.raw", img)
    assert np.array_equal(raw.raw_imread(tmp_path / "t.raw"), img)
    assert np.array_equal(raw.raw_imread(tmp_path / "t.raw", dtype="<u2", shape=(33, 47)), img)


def test_no_silent_cpu_fallback():
    """without a GPU the product path raises; it never computes on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pystripe import core, _native
    with pytest.raises(_native.B200StripeError):
        core.filter_streaks(np.zeros((64, 64), np.uint16) + 5, sigma=(8, 8), wavelet="db2")


def test_product_does_not_import_oracle():
    pkg = ROOT / "image-preprocessing-pipeline_b200"
    for f in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.h")):
        txt = f.read_text()
        assert "import oracle" not in txt and "from oracle" not in txt and "oracle/" not in txt.replace("the oracle", ""), f


@pytest.mark.parametrize("ins,outs", [((100, 130), (87, 111)), ((64, 64), (150, 97)), ((33, 47), (33, 90)),
                                      ((1024, 1024), (864, 864)), ((2, 3), (7, 9)), ((1, 5), (4, 11)),
                                      ((300, 200), (301, 199)), ((17, 19), (171, 23))])
def test_resize_tables_reproduce_scipy_zoom_bit_for_bit(ins, outs):
    """new_size (core.py:1356-1359 -> skimage.transform.resize -> scipy.ndimage.zoom(order=1, mode='mirror',
    grid_mode=True)): the index / weight tables the library builds on the host (b2s_resize_table, GPU-free), evaluated
    with the kernel's expression in numpy float64, equal the real scipy.ndimage.zoom in every bit."""
    import ctypes as C
    from scipy import ndimage as ndi
    from pystripe import _native
    L = _native.lib()

    def table(n_in, n_out):
        i0, i1 = np.zeros(n_out, np.int32), np.zeros(n_out, np.int32)
        w0, w1 = np.zeros(n_out, np.float64), np.zeros(n_out, np.float64)
        L.b2s_resize_table.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        assert L.b2s_resize_table(n_in, n_out, i0.ctypes.data, i1.ctypes.data, w0.ctypes.data, w1.ctypes.data) == 0
        return i0, i1, w0, w1

    img = np.random.default_rng(5).integers(0, 65536, ins).astype(np.uint16)
    ref = ndi.zoom(img.astype(np.float64), [1 / f for f in np.divide(ins, outs)], order=1, mode="mirror", cval=0, grid_mode=True)
    assert ref.shape == outs
    iy0, iy1, wy0, wy1 = table(ins[0], outs[0])
    ix0, ix1, wx0, wx1 = table(ins[1], outs[1])
    v = img.astype(np.float64)
    t = np.zeros(outs)
    t = t + (v[iy0][:, ix0] * wy0[:, None]) * wx0[None, :]
    t = t + (v[iy0][:, ix1] * wy0[:, None]) * wx1[None, :]
    t = t + (v[iy1][:, ix0] * wy1[:, None]) * wx0[None, :]
    t = t + (v[iy1][:, ix1] * wy1[:, None]) * wx1[None, :]
    assert np.array_equal(t, ref)


def test_new_size_geometry_and_rejections():
    from pystripe import _native, core
    p = _native.default_params()
    p.height, p.width, p.in_dtype, p.out_dtype = 96, 128, _native.U16, _native.U16
    p.process_img, p.rotate = 1, 90
    p.new_height, p.new_width = 81, 108
    info = _native.plan_geometry(p)
    assert (info.out_height, info.out_width) == (108, 81)
    p.new_height, p.new_width = 120, 100          # up along y, down along x: skimage's anti-aliasing Gaussian
    with pytest.raises(NotImplementedError):      # ... which the caller has to supply
        _native.plan_geometry(p)
    p.aa_radius_x = 1
    assert (_native.plan_geometry(p).out_height, _native.plan_geometry(p).out_width) == (100, 120)
    assert core._resize_target((96, 128), (96, 128), None, (96, 128)) == (None, (None, None))
    assert core._resize_target((96, 128), (96, 128), (2, 2), (48, 64)) == (None, (None, None))
    assert core._resize_target((96, 128), (96, 128), (2, 2), (40, 50)) == ((40, 50), (None, None))   # down: anti_aliasing=False
    assert core._resize_target((96, 128), (96, 128), None, (130, 171)) == ((130, 171), (None, None))  # up: sigma 0
    # mixed: the weights scipy.ndimage.gaussian_filter would use along the shrinking axis, bit for bit
    from scipy.ndimage import _filters
    size, aa = core._resize_target((96, 128), (96, 128), None, (120, 50))
    assert size == (120, 50) and aa[0] is None
    sd = (128 / 50 - 1) / 2
    radius = int(4.0 * sd + 0.5)
    assert aa[1][0] == radius and np.array_equal(aa[1][1], _filters._gaussian_kernel1d(sd, 0, radius)[::-1])


def test_bleach_plan_args_host_logic():
    """host side of correct_bleaching (core.py:521-531): argument checks, the clip levels in the precision numpy.clip
    compares them in, the Butterworth section from scipy.signal, the constant padding value (core.py:1101-1105)."""
    from scipy.signal import butter, sosfilt_zi
    from pystripe import core
    assert core._bleach_plan_args(None, None, None, None, False, False) == (None, 0.0)
    none, pad = core._bleach_plan_args(None, 4.9, 6.0, 8.0, False, False)       # the production call: frequency None
    assert none is None and pad == float(np.float32(np.log1p(4.9)))
    b, pad = core._bleach_plan_args(1 / 64.0, 0.5, 5.0, np.float64(6.5), False, False)
    sos = butter(1, 1 / 64.0, output='sos')
    assert b[:4] == (sos[0, 0], sos[0, 1], sos[0, 4], sosfilt_zi(sos)[0, 0])
    assert b[4] == float(np.log1p(1))                    # raised to log1p(1), a numpy float64: compared in float64
    assert b[5] == float(np.float32(5.0)) and b[6] == 6.5
    b2, _ = core._bleach_plan_args(0.01, 4.8, 5.6, 7.4, False, False)           # weak Python floats: float32 bounds
    assert b2[4] == float(np.float32(4.8)) and b2[6] == float(np.float32(7.4))
    per_plane, _ = core._bleach_plan_args(0.01, None, 5.0, 6.0, False, False)    # a level left to multi-Otsu (core.py:1066-1077):
    assert per_plane[8] == 1 and per_plane[4:7] == (0.0, 0.0, 0.0)               # per-plane device data, not plan constants
    assert core._bleach_plan_args(0.01, 4.0, 5.0, 6.0, True, False)[0][7] == 2    # max method (batch_filter's default)
    assert b[7] == 1 and b[8] == 0
    # per-plane levels: multi-Otsu values for the missing ones, the same checks and float32 / float64 rules as explicit levels
    lv = core._clip_levels(np.float32(0.5), np.float32(5.0), 6.5)
    assert lv == (float(np.log1p(1)), 5.0, float(np.float32(6.5)))
    assert core._bleach_plan_args(None, None, None, None, False, True) == (None, 0.0)   # masking alone: no bleach data
    # get_img_mask's `img > threshold` as numpy evaluates it: weak Python scalars round to float32 against the log image,
    # a float64 numpy scalar promotes the comparison, integer images compare exactly
    assert core._mask_threshold(6.8, np.float32) == float(np.float32(6.8)) != 6.8
    assert core._mask_threshold(np.float64(6.8), np.float32) == 6.8
    assert core._mask_threshold(np.float32(6.8), np.float32) == float(np.float32(6.8))
    assert core._mask_threshold(900, np.uint16) == 900.0 and core._mask_threshold(900.5, np.uint16) == 900.5
    with pytest.raises(AssertionError):
        core._bleach_plan_args(0.01, 5.0, 4.0, 6.0, False, False)
    with pytest.raises(AssertionError):
        core._bleach_plan_args(1, 4.0, 5.0, 6.0, False, False)                   # int frequency, as the reference asserts


def test_integration_stub_mirrors_the_parameter_struct():
    """the ctypes stub shown to reference maintainers lists struct b2s_params field for field."""
    import re
    doc = (ROOT / "INTEGRATION.md").read_text()
    block = doc[doc.index("class B2SParams"):doc.index("_PAD = ")]
    shown = re.findall(r'\("([a-z_0-9]+)", C\.', block)
    assert shown == [f[0] for f in _lib().Params._fields_]
