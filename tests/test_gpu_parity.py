"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): integer outputs within +-1 LSB on every pixel and >= 99.99 % bit-exact (MIN_EXACT);
float intermediates within 1e-4 relative.  Every stage reproduces the reference's float32 rounding sequence (log1p,
padding, analysis/synthesis filter banks, the scipy.fftpack real FFT passes, expm1, integer epilogues) and is asserted
BIT-EXACT stage by stage; the only length class that is within tolerance instead is an even FFT length > 1000
(DESIGN.md section 5).
"""
import json
from pathlib import Path

import numpy as np
import pytest

from oracle import pystripe_oracle as orc
from oracle import pywt_shim as pw
from tests.golden import cases
from tools import synth

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]
MIN_EXACT = 0.9999         # north-star bar: >= 99.99 % of integer pixels bit-exact per case (measured: 100 %, profiles/r01_parity_report_v3_exact_fft.json)
REPORT = {}


def _cmp_int(name, got, ref, min_exact=MIN_EXACT):
    assert got.dtype == ref.dtype and got.shape == ref.shape, (name, got.dtype, ref.dtype, got.shape, ref.shape)
    d = np.abs(got.astype(np.int64) - ref.astype(np.int64))
    exact = float((d == 0).mean())
    REPORT[name] = {"max_abs_diff": int(d.max()), "exact_fraction": exact, "pixels": int(d.size)}
    assert d.max() <= 1, f"{name}: max |diff| = {d.max()}"
    assert exact >= min_exact, f"{name}: only {exact:.5%} bit-exact"


def test_device_math_is_bit_exact(gpu_ctx):
    rng = np.random.default_rng(0)
    x = np.concatenate([np.arange(65536, dtype=np.float32),
                        rng.uniform(0, 70000, 1 << 20).astype(np.float32),
                        (np.arange(65536, dtype=np.float32) / rng.uniform(0.3, 1, 65536).astype(np.float32))])
    assert np.array_equal(gpu_ctx.debug_math(0, x), orc.log1p_f32(x))
    y = np.concatenate([rng.uniform(-2, 12, 1 << 20).astype(np.float32), orc.log1p_f32(x)])
    assert np.array_equal(gpu_ctx.debug_math(1, y), orc.expm1_f32(y))


def _plan(shape, code, **kw):
    from pystripe import core
    base = dict(process=0, sigma=(24, 24), level=0, wavelet="db10", threshold=None, padding_mode="wrap",
                bidirectional=False, log1p=True)
    base.update(kw)
    return core._get_plan(0, shape, code, **base)


@pytest.mark.parametrize("mode", ["reflect", "wrap", "symmetric", "edge", "constant"])
def test_prologue_log1p_and_padding_bit_exact(mode):
    img = synth.plane(5, (70, 91))
    plan = _plan(img.shape, 1, sigma=(40, 40), padding_mode=mode, stop_after=1)
    plan.run_host(img)
    got = plan.debug_read(0)
    base, py, px = orc.padded_geometry(img.shape, (40, 40), mode)
    ref = np.pad(orc.log1p_f32(img.astype(np.float32)), ((base, base + py), (base, base + px)), mode=mode)
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("mode", ["linear_ramp", "maximum", "mean", "median", "minimum"])
@pytest.mark.parametrize("shape", [(70, 91), (131, 300)])
def test_computed_padding_modes_bit_exact(mode, shape):
    """numpy.pad modes that compute their pad area (core.py:1088-1110 accepts all eleven): column statistics over the
    original rows first, then row statistics over every row of the padded array; float32 rounding as numpy's reductions
    (sequential down a column, pairwise along a row), linear ramps in float64 with numpy's zero-step switch."""
    for variant in range(2):
        img = synth.plane(5 + variant, shape)
        if variant:
            img[0, 3] = 0          # a zero edge value: numpy.linspace switches the whole side to (i / n) * edge
            img[-1, -1] = 0
            img[7, 0] = 0
        plan = _plan(img.shape, 1, sigma=(40, 40), padding_mode=mode, stop_after=1)
        plan.run_host(np.stack([img, img[::-1].copy()]))
        base, py, px = orc.padded_geometry(img.shape, (40, 40), mode)
        for z, src in enumerate((img, img[::-1])):
            got = plan.debug_read(0, plane=z)
            ref = np.pad(orc.log1p_f32(src.astype(np.float32)), ((base, base + py), (base, base + px)), mode=mode)
            assert np.array_equal(got, ref), (mode, shape, variant, z, float(np.abs(got - ref).max()))


def test_filter_streaks_with_computed_padding_modes():
    from pystripe import core
    img = synth.plane(9, (150, 180))
    for mode in ("mean", "maximum", "linear_ramp", "median", "minimum"):
        got = core.filter_streaks(img, sigma=(20, 20), wavelet="db5", padding_mode=mode)
        ref = orc.filter_streaks(img, sigma=(20, 20), wavelet="db5", padding_mode=mode)
        _cmp_int(f"pad_mode/{mode}", got, ref)
    out = core.filter_streaks(img, sigma=(20, 20), wavelet="db5", padding_mode="empty")     # numpy: uninitialised pad area
    assert out.shape == img.shape and out.dtype == img.dtype
    with pytest.raises(RuntimeError):
        core.filter_streaks(img, sigma=(20, 20), padding_mode="no_such_mode")


@pytest.mark.parametrize("shape,wavelet,sigma", [((96, 128), "db10", (24, 24)), ((70, 91), "db5", (10, 10)),
                                                  ((160, 200), "db2", (64, 64)), ((128, 96), "db9", (16, 16)),
                                                  ((200, 300), "db16", (30, 30)),
                                                  # long filters (F >= 42): one kernel per axis through the scratch buffer
                                                  ((300, 410), "coif15", (20, 20)), ((257, 391), "coif8", (12, 12)),
                                                  ((333, 250), "db38", (16, 16)), ((520, 700), "sym20", (40, 40)),
                                                  ((300, 300), "coif7", (8, 8))])
def test_forward_dwt_bit_exact(shape, wavelet, sigma):
    img = synth.plane(6, shape)
    plan = _plan(shape, 1, sigma=sigma, wavelet=wavelet, stop_after=2)
    plan.run_host(img)
    base, py, px = orc.padded_geometry(shape, sigma, "wrap")
    padded = np.pad(orc.log1p_f32(img.astype(np.float32)), ((base, base + py), (base, base + px)), mode="wrap")
    coeffs = pw.wavedec2(padded, wavelet)
    L = len(coeffs) - 1
    assert plan.info.levels == L
    for lvl in range(1, L + 1):
        ch, cv, cd = coeffs[L - lvl + 1]
        for what, ref in ((2, ch), (3, cv), (4, cd)):
            got = plan.debug_read(what, lvl)
            assert np.array_equal(got, ref), f"level {lvl} band {what}: {np.abs(got - ref).max()}"
    assert np.array_equal(plan.debug_read(1, L), coeffs[0])


def test_inverse_dwt_bit_exact_without_notch():
    """sigma so large that... no: run forward+inverse only by stopping before the notch is not possible, so compare the
    reconstruction after a run whose notch is the identity except DC (tiny sigma): oracle does the same steps."""
    img = synth.plane(7, (96, 128))
    plan = _plan(img.shape, 1, sigma=(24, 24), wavelet="db4", stop_after=4)
    plan.run_host(img)
    got = plan.debug_read(0)
    base, py, px = orc.padded_geometry(img.shape, (24, 24), "wrap")
    padded = np.pad(orc.log1p_f32(img.astype(np.float32)), ((base, base + py), (base, base + px)), mode="wrap")
    ref = orc.filter_subband(padded, 24, 0, "db4")
    rel = np.abs(got - ref).max() / np.abs(ref).max()
    REPORT["log_domain_rel_err_db4"] = float(rel)
    assert rel < 1e-4


def test_notch_stage_close_to_pocketfft():
    img = synth.plane(8, (96, 128))
    plan = _plan(img.shape, 1, sigma=(24, 24), wavelet="db10", stop_after=3, bidirectional=True)
    plan.run_host(img)
    base, py, px = orc.padded_geometry(img.shape, (24, 24), "wrap")
    padded = np.pad(orc.log1p_f32(img.astype(np.float32)), ((base, base + py), (base, base + px)), mode="wrap")
    coeffs = pw.wavedec2(padded, "db10")
    L = len(coeffs) - 1
    for lvl in range(1, L + 1):
        ch, cv, _ = coeffs[L - lvl + 1]
        rh = orc.np_filter_coefficient(ch.copy(), 24 / padded.shape[0], axis=-1)
        rv = orc.np_filter_coefficient(cv.copy(), 24 / padded.shape[1], axis=-2)
        gh, gv = plan.debug_read(2, lvl), plan.debug_read(3, lvl)
        for g, r in ((gh, rh), (gv, rv)):
            assert np.abs(g - r).max() <= 1e-5 * max(1.0, float(np.abs(r).max()))


def _mirrored(n):
    """every length the tests reach is reproduced rounding for rounding (rfft_exact.cu): real passes, Bluestein passes,
    and for even lengths > 1000 whose half length has a prime factor >= 7 the half-length complex transform."""
    return n >= 2


def _simd_rows_only(n):
    """scipy (ducc0) sends the sequences outside its 4-wide SIMD batches (the last nseq % 4) through another route when
    n is even, > 1000 and 8 divides n (5-smooth half length) or the half length; those <= 3 sequences are not mirrored."""
    if n <= 1000 or n % 2:
        return False
    h = n // 2
    for p in (2, 3, 5):
        while h % p == 0:
            h //= p
    return n % 8 == 0 if h == 1 else (n // 2) % 8 == 0


@pytest.mark.parametrize("shape,wavelet,sigma", [
    ((96, 128), "db10", (24, 24)),      # small radices 2,3,4,5 and generic 7..31
    ((300, 274), "db2", (8, 8)),        # 149 / 157 primes >= 135: Bluestein passes
    ((538, 560), "db3", (6, 6)),        # 2*137 = 274, 3*... composites with a Bluestein factor
    ((700, 650), "db4", (20, 20)),
    ((1290, 40), "db2", (4, 4)),        # odd length > 1000 along axis -2 (bidirectional)
    ((2628, 44), "db9", (4, 4)),        # 1326 = 2*3*13*17 along axis -2: half-length complex transform, generic radices
    ((44, 2276), "db9", (4, 4)),        # 1150 = 2*5*5*23 along axis -1
    ((2640, 44), "db9", (4, 4)),        # 1332 = 4*9*37
    ((44, 2580), "db9", (4, 4)),        # 1302 = 2*3*7*31
    ((44, 2492), "db9", (4, 4)),        # 1258 = 2*17*37
    ((44, 3316), "db9", (4, 4)),        # 1670 = 2*5*167: complex Bluestein pass (BASELINE config 5, level 1)
    ((1980, 44), "db9", (4, 4)),        # 1002 = 2*3*167 along axis -2
    ((44, 2136), "db9", (4, 4)),        # 1080 = 8*135 (5-smooth half length: plain real passes)
    ((2024, 44), "db9", (4, 4)),        # 1024 along axis -2
    ((44, 1992), "db9", (4, 4)),        # 1008 = 2*504, 8 | 504: half-length complex transform with radix 8
])
def test_notch_stage_bit_exact_vs_scipy_fftpack(shape, wavelet, sigma):
    img = synth.plane(9, shape)
    plan = _plan(shape, 1, sigma=sigma, wavelet=wavelet, stop_after=3, bidirectional=True)
    plan.run_host(img)
    base, py, px = orc.padded_geometry(shape, sigma, "wrap")
    padded = np.pad(orc.log1p_f32(img.astype(np.float32)), ((base, base + py), (base, base + px)), mode="wrap")
    coeffs = pw.wavedec2(padded, wavelet)
    L = len(coeffs) - 1
    checked = 0
    lengths = []
    for lvl in range(1, L + 1):
        ch, cv, _ = coeffs[L - lvl + 1]
        rh = orc.np_filter_coefficient(ch.copy(), sigma[0] / padded.shape[0], axis=-1)
        rv = orc.np_filter_coefficient(cv.copy(), sigma[0] / padded.shape[1], axis=-2)
        gh, gv = plan.debug_read(2, lvl), plan.debug_read(3, lvl)
        for g, r, n, axis in ((gh, rh, ch.shape[1], 0), (gv, rv, cv.shape[0], 1)):
            lengths.append(n)
            if _mirrored(n):
                if _simd_rows_only(n):           # compare the sequences of scipy's SIMD batches
                    nseq = g.shape[axis] // 4 * 4
                    g, r = (g[:nseq], r[:nseq]) if axis == 0 else (g[:, :nseq], r[:, :nseq])
                assert np.array_equal(g, r), f"level {lvl} n={n}: {np.abs(g - r).max()} ({(g != r).mean():.3%} differ)"
                checked += 1
            else:
                assert np.abs(g - r).max() <= 1e-5 * max(1.0, float(np.abs(r).max()))
    REPORT[f"notch_exact_lengths/{shape}/{wavelet}"] = lengths
    assert checked > 0


def _gpu_case(kind, img, kw):
    from pystripe import core
    if kind == "filter_streaks":
        return core.filter_streaks(img.copy(), **kw)
    kw = dict(kw)
    flat = kw.pop("_flat", None)
    if flat is not None:
        return core.process_img(img.copy(), flat=flat, d_type="uint16", **kw)
    return core.process_img(img.copy(), **kw)


def test_golden_vectors_from_the_reference():
    gold = np.load(ROOT / "tests" / "golden" / "pystripe_golden.npz")
    for name, kind, img, kw in cases.all_cases():
        got = _gpu_case(kind, img, kw)       # incl. fs_nolog: integer pixels without log1p run in float64 (csrc/f64path.cu)
        if gold[name].dtype.kind == "f":
            ref = gold[name]
            assert got.dtype == ref.dtype and np.abs(got - ref).max() <= 1e-4 * np.abs(ref).max(), name
            continue
        # full-range random pixels (up to 65535): one float32 ulp in the log domain is ~100x larger in counts than for
        # camera-like data, so the row-FFT rounding difference shows on more pixels (still within 1 LSB)
        _cmp_int("golden/" + name, np.asarray(got), gold[name], min_exact=MIN_EXACT)


@pytest.mark.parametrize("kw", [
    dict(sigma=(256, 256), wavelet="db10"),                                   # BASELINE config 1 (scaled plane)
    dict(sigma=(128, 512), wavelet="db9", padding_mode="reflect", bidirectional=True),
    dict(sigma=(100, 100), wavelet="db9", padding_mode="reflect", bidirectional=True),  # Step 3 as shipped
    dict(sigma=(64, 64), wavelet="db20", level=3),
    dict(sigma=(128, 512), wavelet="coif15", padding_mode="reflect"),         # BASELINE config 5 (scaled plane), two passes
    dict(sigma=(64, 64), wavelet="coif15", padding_mode="reflect", bidirectional=True),   # post-stitch defaults
    dict(sigma=(32, 32), wavelet="db30", level=2),
])
def test_filter_streaks_512(kw):
    from pystripe import core
    img = synth.plane(11, (512, 512))
    got = core.filter_streaks(img, **kw)
    ref = orc.filter_streaks(img, **kw)
    _cmp_int(f"fs512/{kw}", got, ref)


def test_filter_streaks_float_input_stays_float():
    from pystripe import core
    img = synth.plane(12, (128, 160)).astype(np.float32)
    got = core.filter_streaks(img.copy(), sigma=(32, 32), wavelet="db6")
    ref = orc.filter_streaks(img.copy(), sigma=(32, 32), wavelet="db6")
    assert got.dtype == np.float32
    assert np.abs(got - ref).max() <= 1e-4 * np.abs(ref).max()


def test_odd_ragged_and_tiny_shapes():
    from pystripe import core
    for shape, wavelet, sigma in [((255, 257), "db10", (64, 64)), ((30, 30), "db2", (2, 2)), ((64, 64), "db9", (1, 1)),
                                  ((33, 301), "db3", (20, 20))]:
        img = synth.plane(13, shape)
        _cmp_int(f"ragged/{shape}", core.filter_streaks(img, sigma=sigma, wavelet=wavelet),
                 orc.filter_streaks(img, sigma=sigma, wavelet=wavelet))


def test_sigma_zero_is_identity_and_errors():
    from pystripe import core
    img = synth.plane(14, (64, 64))
    assert core.filter_streaks(img, sigma=(0, 0)) is img
    with pytest.raises(ValueError):
        core.filter_streaks(img, sigma=(0, 8), wavelet="db2")


def test_batch_equals_single_and_torch_zero_copy():
    import torch
    from pystripe import core
    stack = synth.stack(5, (128, 160))
    kw = dict(sigma=(32, 32), wavelet="db10", padding_mode="reflect")
    whole = core.filter_streaks(stack, **kw)
    for z in range(5):
        assert np.array_equal(whole[z], core.filter_streaks(stack[z], **kw))
    t = torch.from_numpy(stack).cuda()
    out = core.filter_streaks(t, **kw)
    assert out.is_cuda and out.dtype == torch.uint16
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), whole)


@pytest.mark.parametrize("kw", [
    dict(sigma=(32, 32), wavelet="db10", dark=100, padding_mode="reflect"),
    dict(sigma=(32, 32), wavelet="db10", dark=100, convert_to_8bit=True, bit_shift_to_right=4, down_sample=(2, 2)),
    dict(sigma=(0, 0), dark=120.5, convert_to_8bit=True, bit_shift_to_right=0),
    dict(sigma=(16, 16), wavelet="db4", down_sample=(2, 3), down_sample_method="mean"),
    dict(sigma=(16, 16), wavelet="db4", gaussian_filter_2d=True, rotate=90),
    dict(sigma=(0, 0), gaussian_filter_2d=True, flip_upside_down=True),
])
def test_process_img_variants(kw):
    from pystripe import core
    img = synth.plane(15, (130, 171))
    got = core.process_img(img.copy(), **kw)
    ref = orc.process_img(img.copy(), **kw)
    _cmp_int(f"pi/{kw}", got, ref)


def test_process_img_flat_field_float_path():
    from pystripe import core
    img = synth.plane(16, (128, 160))
    flat = orc.normalize_flat(synth.flat_field((128, 160)))
    kw = dict(sigma=(32, 32), wavelet="db10", dark=100, padding_mode="reflect")
    got = core.process_img(img.copy(), flat=flat, **kw)
    ref = orc.process_img(img.copy(), flat=flat, **kw)
    _cmp_int("pi/flat", got, ref)


def test_process_img_config3_flat_gaussian_downsample_8bit():
    """BASELINE configs[2] at a small size: flat + 5x5 Gaussian on the float image + (2,2) block max + destripe + dark +
    8-bit shift.  cv2's float Gaussian is not bit-pinned (pointwise.cu), so uint8 is asserted within 1 LSB / 99.9 %."""
    from pystripe import core
    img = synth.plane(17, (256, 320))
    flat = orc.normalize_flat(synth.flat_field((256, 320)))
    for shift in (8, 4):
        kw = dict(sigma=(32, 32), wavelet="db10", dark=100, padding_mode="reflect", gaussian_filter_2d=True,
                  down_sample=(2, 2), convert_to_8bit=True, bit_shift_to_right=shift)
        got = core.process_img(img.copy(), flat=flat, **kw)
        ref = orc.process_img(img.copy(), flat=flat, **kw)
        _cmp_int(f"pi/config3/shift{shift}", got, ref, min_exact=0.999)


def test_process_img_median_downsample():
    from pystripe import core
    img = synth.plane(18, (130, 171))
    kw = dict(sigma=(16, 16), wavelet="db4", down_sample=(2, 3), down_sample_method="median")
    _cmp_int("pi/median", core.process_img(img.copy(), **kw), orc.process_img(img.copy(), **kw))


@pytest.mark.parametrize("shape,kw", [
    ((300, 340), dict(sigma=(0, 0), lightsheet=True)),                                   # defaults: 150 / 200 / 0.25 / 2.0
    ((256, 512), dict(sigma=(0, 0), lightsheet=True, dark=110, artifact_length=64, background_window_size=100,
                      percentile=0.4, lightsheet_vs_background=3.0, rotate=90, flip_upside_down=True)),
    ((200, 260), dict(sigma=(24, 24), wavelet="db6", lightsheet=True, dark=100, convert_to_8bit=True,
                      bit_shift_to_right=2, padding_mode="reflect")),
    ((131, 177), dict(sigma=(0, 0), lightsheet=True, artifact_length=50, background_window_size=75, percentile=1.0)),
])
def test_lightsheet_clean_bit_exact(shape, kw):
    """lightsheet_correct (numba percentile + scipy zoom + wrap-around uint16 arithmetic) is integer/float64 work:
    bit-exact, including scipy's zero column when the last zoom coordinate rounds above the grid."""
    from pystripe import core
    rng = np.random.default_rng(21)
    img = synth.plane(19, shape)
    img[:: 7] += rng.integers(0, 300, img[::7].shape).astype(np.uint16)     # row artefacts
    got = core.process_img(img.copy(), **kw)
    ref = orc.process_img(img.copy(), **kw)
    _cmp_int(f"lightsheet/{shape}", got, ref, min_exact=1.0 if tuple(kw["sigma"]) == (0, 0) else MIN_EXACT)


def test_lightsheet_on_flat_fielded_float_image():
    from pystripe import core
    img = synth.plane(20, (160, 200))
    flat = orc.normalize_flat(synth.flat_field((160, 200)))
    kw = dict(sigma=(0, 0), lightsheet=True, dark=90, artifact_length=40, background_window_size=60)
    got = core.process_img(img.copy(), flat=flat, **kw)
    ref = orc.process_img(img.copy(), flat=flat, **kw)
    _cmp_int("lightsheet/flat", got, ref, min_exact=1.0)


def test_flat_field_without_destripe():
    from pystripe import core
    img = synth.plane(22, (96, 130))
    flat = orc.normalize_flat(synth.flat_field((96, 130)))
    kw = dict(sigma=(0, 0), dark=50, convert_to_8bit=True, bit_shift_to_right=3)
    _cmp_int("pi/flat_nodestripe", core.process_img(img.copy(), flat=flat, **kw), orc.process_img(img.copy(), flat=flat, **kw),
             min_exact=1.0)


def test_uniform_plane_shortcut():
    from pystripe import core
    stack = synth.stack(3, (64, 80))
    stack[1] = 9
    out = core.process_img(stack, sigma=(8, 8), wavelet="db2", convert_to_8bit=True, rotate=90)
    assert out.shape == (3, 80, 64) and out.dtype == np.uint8
    assert (out[1] == 0).all() and out[0].any()
    for z in (0, 2):
        assert np.array_equal(out[z], orc.process_img(stack[z], sigma=(8, 8), wavelet="db2", convert_to_8bit=True, rotate=90)) \
            or np.abs(out[z].astype(int) - orc.process_img(stack[z], sigma=(8, 8), wavelet="db2", convert_to_8bit=True, rotate=90)).max() <= 1


def test_full_size_plane_against_oracle():
    """BASELINE config 1 at full size: one 2048x2048 plane, sigma=(256,256), db10, defaults."""
    from pystripe import core
    img = synth.plane(0, (2048, 2048))
    got = core.filter_streaks(img, sigma=(256, 256), wavelet="db10")
    ref = orc.filter_streaks(img, sigma=(256, 256), wavelet="db10")
    _cmp_int("config1/2048", got, ref)


def test_large_non_square_plane_against_oracle():
    """beyond the benchmark size: 3000 x 4096, the production wavelet, both axes filtered (sub-band sides 1655 x 2203,
    i.e. a Bluestein factor 331 and the prime 2203), batch of two planes."""
    from pystripe import core
    stack = np.stack([synth.plane(40, (3000, 4096)), synth.plane(41, (3000, 4096))])
    kw = dict(sigma=(250, 250), wavelet="db9", padding_mode="reflect", bidirectional=True)
    got = core.filter_streaks(stack, **kw)
    ref = orc.filter_streaks(stack[1], **kw)
    _cmp_int("large/3000x4096", got[1], ref)


def test_full_size_property_columns_only_image_is_a_fixed_point():
    """size-independent property: an image that is constant along y has cH == 0 at every level, so the destripe must
    return it unchanged (the float32 wavelet round trip is far below half an LSB)."""
    from pystripe import core
    rng = np.random.default_rng(5)
    row = rng.integers(100, 4000, 2048).astype(np.uint16)
    img = np.ascontiguousarray(np.broadcast_to(row, (2048, 2048)))
    out = core.filter_streaks(img, sigma=(256, 256), wavelet="db10", padding_mode="reflect")
    assert np.array_equal(out, img)


def test_batch_filter_step3_call_on_tiff_files(tmp_path):
    """the call process_images.py:420 makes (db9, reflect, bidirectional, uint16 out) on a small tile directory, plus
    dark and flat: files in, files out, every plane identical to the oracle's process_img."""
    from pystripe import core
    src, dst = tmp_path / "in", tmp_path / "out"
    (src / "ch0").mkdir(parents=True)
    flat = synth.flat_field((96, 120))
    planes = {}
    for z in range(11):
        img = synth.plane(30 + z, (96, 120))
        if z == 4:
            img[:] = 123                                    # uniform plane -> zeros
        core.imsave_tif(src / "ch0" / f"img_{z:05d}.tif", img, compression=None)
        planes[f"ch0/img_{z:05d}.tif"] = img
    kw = dict(sigma=(16, 16), level=0, wavelet="db9", padding_mode="reflect", bidirectional=True, dark=105)
    rc = core.batch_filter(src, dst, workers=4, threads_per_gpu=4, flat=flat, d_type="uint16", compression=None, **kw)
    assert rc == 0
    nflat = orc.normalize_flat(flat)
    for name, img in planes.items():
        got = core.imread_tif_raw_png(dst / name)
        ref = orc.process_img(img.copy(), flat=nflat, d_type=np.dtype("uint16"), **kw)
        assert got.dtype == np.uint16 and np.array_equal(got, ref), name
    # continue_process skips what exists
    assert core.batch_filter(src, dst, workers=2, flat=flat, d_type="uint16", continue_process=True, **kw) == 0


def test_zz_write_parity_report():
    out = ROOT / "gpurun_out"
    out.mkdir(exist_ok=True)
    (out / "parity_report.json").write_text(json.dumps(REPORT, indent=1, default=str))


@pytest.mark.parametrize("shape,kw", [
    ((200, 260), dict(sigma=(24, 24), wavelet="db6", dark=100, new_size=(173, 211), convert_to_8bit=True, bit_shift_to_right=4)),
    ((200, 260), dict(sigma=(24, 24), wavelet="db6", new_size=(333, 401))),
    ((256, 256), dict(sigma=(32, 32), wavelet="db10", padding_mode="reflect", dark=100, down_sample=(2, 2), new_size=(108, 108),
                      convert_to_8bit=True, bit_shift_to_right=8, rotate=270)),          # config 3 with its optional resize
    ((200, 260), dict(sigma=(0, 0), new_size=(77, 300 - 41), flip_upside_down=True, convert_to_16bit=True)),
    ((200, 260), dict(sigma=(24, 24), wavelet="db6", lightsheet=True, artifact_length=40, background_window_size=50,
                      dark=100, new_size=(150, 195), rotate=180)),
])
def test_process_img_new_size(shape, kw):
    """new_size: order-1 resize (skimage.transform.resize restated over the real scipy.ndimage.zoom in the oracle)
    between the lightsheet stage and the final conversion; float64 arithmetic in scipy's order => bit-exact."""
    from pystripe import core
    stack = np.stack([synth.plane(50, shape), synth.plane(51, shape), np.full(shape, 9, np.uint16)])
    got = core.process_img(stack, **kw)
    for z in range(len(stack)):
        ref = orc.process_img(stack[z].copy(), **kw)
        assert got[z].shape == ref.shape and got[z].dtype == ref.dtype
        _cmp_int(f"new_size/{kw}/{z}", got[z], ref)


def test_process_img_new_size_after_flat_is_float32_zoom():
    from pystripe import core
    img = synth.plane(52, (160, 200))
    flat = core.normalize_flat(synth.flat_field((160, 200)))
    kw = dict(sigma=(16, 16), wavelet="db6", dark=100, padding_mode="reflect", new_size=(131, 163), d_type="uint16")
    got = core.process_img(img, flat=flat, **kw)
    ref = orc.process_img(img.copy(), flat=orc.normalize_flat(synth.flat_field((160, 200))), **kw)
    _cmp_int("new_size/flat", got, ref)


def test_bleach_correction_goldens_bit_exact():
    """correct_bleaching (core.py:501-559: clip, scipy.signal.sosfiltfilt of a first-order Butterworth section in float64,
    img / filter * max) inside filter_streaks: every bleach golden written by the reference run verbatim, bit for bit —
    float32 outputs included."""
    gold = np.load(ROOT / "tests" / "golden" / "pystripe_golden.npz")
    n = 0
    for name, kind, img, kw in cases.all_cases():
        if "bleach" not in name and "clipmin" not in name:
            continue
        got = np.asarray(_gpu_case(kind, img, kw))
        ref = gold[name]
        assert got.dtype == ref.dtype and got.shape == ref.shape, name
        same = got.view(np.uint32) == ref.view(np.uint32) if ref.dtype == np.float32 else got == ref
        REPORT["bleach/" + name] = {"exact_fraction": float(same.mean())}
        assert same.all(), (name, float(same.mean()), float(np.abs(got.astype(np.float64) - ref).max()))
        n += 1
    assert n == 11


@pytest.mark.parametrize("shape,freq", [((700, 900), 1 / 700.0), ((257, 2051), 1 / 2048.0)])
def test_bleach_correction_against_oracle(shape, freq):
    from pystripe import core
    img = synth.plane(11, shape)
    kw = dict(sigma=(64, 64), wavelet="db9", padding_mode="reflect", bleach_correction_frequency=freq,
              bleach_correction_clip_min=4.7, bleach_correction_clip_med=5.5, bleach_correction_clip_max=7.9)
    got = core.filter_streaks(np.stack([img, img[::-1].copy()]), **kw)          # batch of two through one plan
    for z, src in enumerate((img, img[::-1].copy())):
        ref = orc.filter_streaks(src, **kw)
        _cmp_int(f"bleach/{shape}/{z}", got[z], ref, min_exact=1.0)


@pytest.mark.parametrize("kw", [dict(), dict(bleach_correction_clip_min=4.9), dict(padding_mode="constant"),
                                dict(bleach_correction_max_method=True, bleach_correction_clip_max=7.8)])
def test_bleach_clip_levels_from_multiotsu_per_plane(kw):
    """clip levels left to threshold_multiotsu (core.py:1066-1077): per-plane levels from exact GPU histograms, every plane of
    a batch with its own levels (and its own constant-padding value)."""
    from pystripe import core
    stack = np.stack([synth.plane(21, (300, 420)), (synth.plane(22, (300, 420)).astype(np.uint32) * 3).clip(0, 65535).astype(np.uint16),
                      synth.plane(23, (300, 420))[::-1].copy()])
    base = dict(sigma=(32, 32), wavelet="db9", padding_mode="reflect", bleach_correction_frequency=1 / 300.0)
    base.update(kw)
    got = core.filter_streaks(stack, **base)
    for z in range(3):
        ref = orc.filter_streaks(stack[z], **base)
        _cmp_int(f"bleach_otsu/{kw}/{z}", got[z], ref, min_exact=1.0)
    # process_img: a uniform plane in the batch is zeros and does not disturb the others
    stack[1] = 9
    out = core.process_img(stack, dark=100, **base)
    assert not out[1].any()
    assert np.array_equal(out[0], orc.process_img(stack[0].copy(), dark=100, **base))


@pytest.mark.parametrize("int_path,dark,work", [(1, 0.0, 1), (0, 100.0, 1), (1, 100.0, 1), (1, 7.0, 0), (0, 0.0, 1)])
def test_epilogue_threshold_table_is_exhaustively_exact(int_path, dark, work):
    """the fast epilogue's log value -> integer map (expm1f, [rint + clip], dark, clip) comes from 65 535 thresholds + an
    approximate exponential (pointwise.cu): identical to the direct evaluation of the fdlibm mirror for EVERY float32 bit
    pattern (2^32 of them, NaNs and infinities included), which also proves the mirror monotone where it matters."""
    import ctypes as C
    from pystripe import _native
    ctx = _native.context(0)
    bad = C.c_uint64(123)
    for first in range(0, 1 << 32, 1 << 30):
        ctx.check(_native.lib().b2s_debug_expm1_table_check(ctx._h, int_path, dark, work, first, 1 << 30, C.byref(bad)))
        assert bad.value == 0, (hex(first), bad.value)


def test_get_img_mask_matches_the_oracle():
    """get_img_mask (core.py:475-489) on the GPU: threshold, box close / open (even and odd kernels, kernels larger than the
    plane), corner flood fills — every mask identical to the oracle's (itself pinned against cv2)."""
    import torch
    from scipy import ndimage
    from pystripe import core
    rng = np.random.default_rng(5)
    for t in range(12):
        h, w = (int(v) for v in rng.integers(40, 400, 2))
        f = ndimage.gaussian_filter(rng.random((3, h, w)), (0, 4, 4)).astype(np.float32)
        thr = float(np.quantile(f, rng.uniform(0.3, 0.7)))
        c, o = int(rng.integers(1, 14)), int(rng.integers(1, 40))
        got = core.get_img_mask(f, thr, c, o)
        assert got.dtype == bool and got.shape == f.shape
        for z in range(3):
            assert np.array_equal(got[z], orc.get_img_mask(f[z], np.float32(thr), c, o)), (t, z, h, w, c, o)
    ring = cases.mask_plane()
    assert np.array_equal(core.get_img_mask(ring, 900, 5, 9), orc.get_img_mask(ring, 900, 5, 9))
    assert np.array_equal(core.get_img_mask(ring, 900), orc.get_img_mask(ring, 900))                # 50 / 500 on a small plane
    big = np.zeros((1200, 1500), np.uint16)
    big[100:1100, 200:1300] = 1000
    big[400:700, 500:900] = 0                                                                        # a hole
    big[0:30, 0:30] = 1000
    spiral = core.get_img_mask(torch.from_numpy(big).cuda(), 500)                                    # the reference's defaults
    assert spiral.is_cuda and np.array_equal(spiral.cpu().numpy(), orc.get_img_mask(big, 500))
    with pytest.raises(NotImplementedError):
        core.get_img_mask(ring, 900, flood_fill_flag=8)


@pytest.mark.parametrize("kw", [
    dict(sigma=(16, 16), wavelet="db6", close_steps=4, open_steps=7, bleach_correction_frequency=1 / 64.0),     # Otsu: levels + threshold
    dict(sigma=(12, 12), wavelet="db3", padding_mode="symmetric", close_steps=3, open_steps=5),                # Otsu threshold only
    dict(sigma=(16, 16), wavelet="db4", padding_mode="maximum", close_steps=5, open_steps=9, bleach_correction_clip_med=6.8),
    dict(sigma=(16, 16), wavelet="db4", padding_mode="linear_ramp", close_steps=5, open_steps=9, bleach_correction_clip_med=np.float64(6.8)),
    dict(sigma=(8, 8), wavelet="db2", padding_mode="wrap", close_steps=None, open_steps=9, bleach_correction_clip_med=6.8),   # no mask
])
def test_enable_masking_per_plane(kw):
    """filter_streaks(enable_masking=True): img *= get_img_mask(img, clip_med) on the log image ahead of numpy.pad
    (core.py:1079-1080); clip_med left to multi-Otsu gives every plane of a batch its own threshold."""
    from pystripe import core
    a = cases.mask_plane()
    b = cases.mask_plane(seed=22)[::-1].copy()
    b[b > 2000] += 700
    stack = np.stack([a, b])
    got = core.filter_streaks(stack, enable_masking=True, **kw)
    for z in range(2):
        ref = orc.filter_streaks(stack[z], enable_masking=True, **kw)
        _cmp_int(f"mask/{kw}/{z}", got[z], ref)
        assert kw.get("close_steps") is None or (ref == 0).mean() > 0.3


@pytest.mark.parametrize("kw", [dict(sigma=(24, 24), wavelet="db3"), dict(sigma=(16, 48), wavelet="db6", padding_mode="reflect"),
                                dict(sigma=(20, 20), wavelet="db9", bidirectional=True, padding_mode="symmetric")])
def test_integer_pixels_without_log1p_run_in_float64(kw):
    """log1p_normalization_needed=False on uint16 / uint8 pixels: pywt and scipy promote to float64 and every pass ends with
    `.astype(d_type)` (truncation, core.py:939).  The GPU evaluates the same chain in float64 without mirroring the operation
    order: identical integers except within ~1e-9 of an integer boundary."""
    from pystripe import core
    for dtype, shape in ((np.uint16, (150, 211)), (np.uint8, (96, 128))):
        img = synth.plane(31, shape)
        img = (img >> 4).astype(np.uint8) if dtype == np.uint8 else img
        stack = np.stack([img, img[::-1].copy()])
        got = core.filter_streaks(stack, log1p_normalization_needed=False, **kw)
        for z in range(2):
            ref = orc.filter_streaks(stack[z], log1p_normalization_needed=False, **kw)
            _cmp_int(f"nolog_int/{kw}/{dtype.__name__}/{z}", got[z], ref)
    out = core.process_img(stack, log1p_normalization_needed=False, dark=3, convert_to_8bit=False, **kw)
    assert np.array_equal(out[0], orc.process_img(stack[0].copy(), log1p_normalization_needed=False, dark=3, **kw))


def test_bleach_argument_errors():
    from pystripe import core
    img = synth.plane(0, (64, 64))
    with pytest.raises(NotImplementedError):          # multi-Otsu levels need the integer image entering filter_streaks
        core.process_img(img, sigma=(8, 8), bleach_correction_frequency=0.01, gaussian_filter_2d=True)
    with pytest.raises(AssertionError):               # core.py:524-527
        core.filter_streaks(img, sigma=(8, 8), bleach_correction_frequency=0.01, bleach_correction_clip_min=5.0,
                            bleach_correction_clip_med=4.0, bleach_correction_clip_max=6.0)
    with pytest.raises(NotImplementedError):          # the threshold would come from skimage's exact integer histogram path
        core.filter_streaks(img, sigma=(8, 8), enable_masking=True, log1p_normalization_needed=False)
    with pytest.raises(NotImplementedError):          # per-image log1p(multi-Otsu clip_min) as the constant-padding value
        core.filter_streaks(img, sigma=(8, 8), enable_masking=True, padding_mode="constant")
