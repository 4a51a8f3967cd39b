"""CPU: the oracle against known answers, the golden vectors made from the reference run verbatim, and (where
/root/reference exists) the reference itself."""
import json
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
from scipy.fftpack import rfft

from oracle import pystripe_oracle as orc
from oracle import pywt_shim as pw
from oracle import ref_runner
from tests.golden import cases

ROOT = Path(__file__).resolve().parents[1]

# known-answer vectors
REC_LO = {
    "db2": [0.48296291314469025, 0.836516303737469, 0.22414386804185735, -0.12940952255092145],
    "db3": [0.3326705529509569, 0.8068915093133388, 0.4598775021193313, -0.13501102001039084, -0.08544127388224149,
            0.035226291882100656],
    "db4": [0.23037781330885523, 0.7148465705525415, 0.6308807679295904, -0.02798376941698385, -0.18703481171888114,
            0.030841381835986965, 0.032883011666982945, -0.010597401784997278],
}
# published PyWavelets tables (dec_lo): sym4 from SURVEY.md section 8c, coif1-3 as printed by pywt.Wavelet(name).dec_lo
DEC_LO_KAT = {
    "sym7": [0.002681814568257878, -0.0010473848886829163, -0.01263630340325193, 0.03051551316596357, 0.0678926935013727,
             -0.049552834937127255, 0.017441255086855827, 0.5361019170917628, 0.767764317003164, 0.2886296317515146,
             -0.14004724044296152, -0.10780823770381774, 0.004010244871533663, 0.010268176708511255],
    "sym4": [-0.07576571478927333, -0.02963552764599851, 0.49761866763201545, 0.8037387518059161, 0.29785779560527736,
             -0.09921954357684722, -0.012603967262037833, 0.0322231006040427],
    "coif1": [-0.01565572813546454, -0.0727326195128539, 0.38486484686420286, 0.8525720202122554, 0.3378976624578092,
              -0.0727326195128539],
    "coif2": [-0.0007205494453645122, -0.0018232088707029932, 0.0056114348193944995, 0.023680171946334084,
              -0.0594344186464569, -0.0764885990783064, 0.41700518442169254, 0.8127236354455423, 0.3861100668211622,
              -0.06737255472196302, -0.04146493678175915, 0.016387336463522112],
    "coif3": [-3.459977283621256e-05, -7.098330313814125e-05, 0.0004662169601128863, 0.0011175187708906016,
              -0.0025745176887502236, -0.00900797613666158, 0.015880544863615904, 0.03455502757306163,
              -0.08230192710688598, -0.07179982161931202, 0.42848347637761874, 0.7937772226256206, 0.4051769024096169,
              -0.06112339000267287, -0.0657719112818555, 0.023452696141836267, 0.007782596427325418,
              -0.003793512864491014],
}
DB10_DEC_LO = [-1.3264202894521245e-5, 9.3588670320069591e-5, -1.1646685512928545e-4, -6.8585669495971163e-4,
               1.9924052951850561e-3, 1.3953517470529012e-3, -1.0733175483330575e-2, 3.6065535669561697e-3,
               3.3212674059341002e-2, -2.9457536821875813e-2, -7.1394147166397087e-2, 9.3057364603572351e-2,
               1.2736934033579326e-1, -1.9594627437737704e-1, -2.4984642432731538e-1, 2.8117234366057746e-1,
               6.8845903945360357e-1, 5.2720118893172559e-1, 1.8817680007769149e-1, 2.6670057900555554e-2]


@pytest.mark.parametrize("name", sorted(REC_LO))
def test_wavelet_known_answers(name):
    w = pw.Wavelet(name)
    assert np.allclose(w.rec_lo, REC_LO[name], atol=5e-12, rtol=0)
    assert np.allclose(w.dec_lo, REC_LO[name][::-1], atol=5e-12, rtol=0)


@pytest.mark.parametrize("name", sorted(DEC_LO_KAT))
def test_sym_coif_known_answers(name):
    assert np.allclose(pw.Wavelet(name).dec_lo, DEC_LO_KAT[name], atol=1e-11, rtol=0)
    assert np.array_equal(pw.Wavelet("sym2").dec_lo, pw.Wavelet("db2").dec_lo)
    assert np.array_equal(pw.Wavelet("sym3").dec_lo, pw.Wavelet("db3").dec_lo)


def test_db10_table():
    assert np.allclose(pw.Wavelet("db10").dec_lo, DB10_DEC_LO, atol=1e-15, rtol=1e-13)


@pytest.mark.parametrize("name", ["db1", "db2", "db5", "db9", "db10", "db16", "db20", "db38", "sym5", "sym8", "sym13", "sym20",
                                  "coif1", "coif5", "coif8", "coif12", "coif15", "coif17"])
def test_wavelet_orthonormal_and_moments(name):
    w = pw.Wavelet(name)
    h = w.dec_lo
    F = h.size
    if name.startswith("coif"):      # 6N taps: 2N vanishing wavelet moments, scaling moments 1..2N-1 vanish about tap 2N
        N = F // 6
        assert abs(h.sum() - np.sqrt(2)) < 1e-12
        for m in range(F // 2):
            assert abs(float(np.dot(h[: F - 2 * m], h[2 * m:])) - (1.0 if m == 0 else 0.0)) < 1e-12
        kc = (np.arange(F) - 2 * N) / (2.0 * N)
        r = w.rec_lo
        for p in range(2 * N):
            assert abs(np.dot(((-1.0) ** np.arange(F)) * kc ** p, r)) < 1e-9 * 2.0 ** p, (name, p)
        for p in range(1, 2 * N):
            assert abs(np.dot(kc ** p, r)) < 1e-9 * 2.0 ** p, (name, p)
        return
    assert abs(h.sum() - np.sqrt(2)) < 1e-13
    for m in range(F // 2):
        s = float(np.dot(h[: F - 2 * m], h[2 * m:]))
        assert abs(s - (1.0 if m == 0 else 0.0)) < 1e-12
    # N = F/2 vanishing moments of the high-pass
    g = w.dec_hi
    k = np.arange(F, dtype=np.float64)
    for p in range(F // 2):
        assert abs(np.dot(g, k ** p)) < 1e-6 * max(1.0, (F ** p))
    assert np.array_equal(w.rec_hi, [(-1) ** i * w.rec_lo[F - 1 - i] for i in range(F)])


@pytest.mark.parametrize("shape,wav", [((64, 64), "db2"), ((101, 60), "db4"), ((39, 183), "db3"), ((200, 266), "db10")])
@pytest.mark.parametrize("dt,tol", [(np.float64, 1e-12), (np.float32, 2e-5)])
def test_perfect_reconstruction(shape, wav, dt, tol):
    x = np.random.default_rng(0).standard_normal(shape).astype(dt)
    c = pw.wavedec2(x, wav)
    r = pw.waverec2(c, wav)
    assert r.dtype == dt
    assert np.abs(r[: shape[0], : shape[1]] - x).max() < tol


def test_level_rule_and_shapes():
    c = pw.wavedec2(np.zeros((2648, 2648), np.float32), "db10")
    assert [d[0].shape[0] for d in c[1:]][::-1] == [1333, 676, 347, 183, 101, 60, 39]
    assert pw.dwt_max_level(3252, 90) == 5 and pw.dwt_max_level(2588, 18) == 7 and pw.dwt_max_level(16, 18) == 0


def test_symmetric_extension_definition():
    """dwt output == direct evaluation on the half-sample symmetric extension (float64, order-insensitive)."""
    rng = np.random.default_rng(3)
    for n, name in [(37, "db3"), (20, "db10"), (64, "db4")]:
        x = rng.standard_normal(n)
        w = pw.Wavelet(name)
        F = w.dec_len
        ext = np.concatenate([x[::-1], x, x[::-1]])
        a, d = pw.dwt_axis(x[None, :], w, -1)
        for o in range((n + F - 1) // 2):
            ref = sum(w.dec_lo[j] * ext[n + 2 * o + 1 - j] for j in range(F))
            assert abs(a[0, o] - ref) < 1e-12


def test_scalar_known_answers():
    assert orc.calculate_pad_size((2048, 2048), 256) == 300
    assert orc.calculate_pad_size((2048, 2048), 250) == 294
    assert orc.calculate_pad_size((2048, 2048), 512) == 602
    assert orc.calculate_pad_size((2048, 2048), 100) == 118
    g = orc.np_notch(8, 2.0)
    assert np.allclose(g, [0, .11750311, .39346933, .67534757, .86466473, .9560631, .988891, .9978125], atol=2e-7)
    assert np.allclose(rfft(np.arange(8, dtype=np.float32)), [28, -4, 9.656855, -4, 4, -4, 1.6568542, -4], atol=1e-5)


def _run_case(kind, img, kw):
    if kind == "filter_streaks":
        return orc.filter_streaks(img.copy(), **kw)
    kw = dict(kw)
    flat = kw.pop("_flat", None)
    if flat is not None:
        return orc.process_img(img.copy(), flat=flat, d_type="uint16", **kw)
    quirks = not kw.get("gaussian_filter_2d", False)
    return orc.process_img(img.copy(), quirks=quirks, **kw)


def test_oracle_matches_golden_vectors():
    """goldens were produced by the reference source executed verbatim (tests/golden/make_golden.py)."""
    gold = np.load(ROOT / "tests" / "golden" / "pystripe_golden.npz")
    meta = json.loads((ROOT / "tests" / "golden" / "pystripe_golden.json").read_text())
    n = 0
    for name, kind, img, kw in cases.all_cases():
        got = _run_case(kind, img, kw)
        ref = gold[name]
        assert str(got.dtype) == meta[name]["dtype"] and list(got.shape) == meta[name]["shape"], name
        assert np.array_equal(got, ref), f"{name}: {(got != ref).mean():.4%} pixels differ"
        n += 1
    assert n == len(meta)


def test_get_img_mask_restatement_matches_opencv():
    """oracle.get_img_mask (window counts + scipy.ndimage.label) against the cv2 calls of core.py:479-487, including even
    kernels (asymmetric window), kernels larger than the image and holes that no corner reaches."""
    cv2 = pytest.importorskip("cv2")
    from scipy import ndimage

    def with_cv2(img, thr, c, o):
        mask = (img > thr).astype(np.uint8)
        mask = cv2.morphologyEx(mask, cv2.MORPH_CLOSE, np.ones((c, c), np.uint8))
        mask = cv2.morphologyEx(mask, cv2.MORPH_OPEN, np.ones((o, o), np.uint8)).astype(bool)
        inv = np.logical_not(mask).astype(np.uint8)
        h, w = img.shape
        for pt in ((0, 0), (0, h - 1), (w - 1, 0), (w - 1, h - 1)):
            cv2.floodFill(inv, None, pt, 0, flags=4)
        return mask | inv.astype(bool)
    rng = np.random.default_rng(3)
    holes = 0
    for t in range(60):
        h, w = (int(v) for v in rng.integers(20, 140, 2))
        img = ndimage.gaussian_filter(rng.random((h, w)), rng.uniform(1, 6))
        thr = np.quantile(img, rng.uniform(0.2, 0.8))
        c, o = int(rng.integers(1, 12)), int(rng.integers(1, 25))
        a, b = orc.get_img_mask(img, thr, c, o), with_cv2(img, thr, c, o)
        assert np.array_equal(a, b), (t, h, w, c, o)
        holes += int((b & ~(img > thr)).any())
    assert holes > 5
    ring = cases.mask_plane()
    assert np.array_equal(orc.get_img_mask(ring, 900, 5, 9), with_cv2(ring, 900, 5, 9))
    assert orc.get_img_mask(ring, 900, 5, 9)[70, 80] and not orc.get_img_mask(ring, 900, 5, 9)[0, 0]
    assert np.array_equal(orc.get_img_mask(ring, 900, 50, 500), with_cv2(ring, 900, 50, 500))


@pytest.mark.skipif(not ref_runner.available(), reason="/root/reference only exists in the build container")
def test_oracle_matches_reference_verbatim():
    r = subprocess.run([sys.executable, "-m", "oracle.ref_check"], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    rep = json.loads(r.stdout.strip().splitlines()[-1])
    assert rep["bad"] == 0 and len(rep["cases"]) >= 20


def test_libm_pins():
    x = np.linspace(0, 70000, 200001, dtype=np.float32)
    y = orc.log1p_f32(x)
    assert np.abs(y - np.log1p(x.astype(np.float64))).max() < 2e-6
    z = orc.expm1_f32(y)
    assert (np.abs(z - x) <= 1e-6 * np.maximum(x, 1.0)).all()


def _fft_lib():
    import ctypes
    L = pw.lib()
    L.orc_fftpack_r2r_f32.restype = ctypes.c_int
    return L


@pytest.mark.parametrize("lengths", [list(range(1, 140)), [143, 149, 169, 183, 256, 274, 289, 347, 411, 625, 676, 821, 1000],
                                     [1001, 1029, 1327, 1333, 1337, 1341, 2047, 2187, 3251]])
def test_fft_restatement_is_bit_identical_to_scipy_fftpack(lengths):
    """oracle/pocketfft_c.c (the pass structure the GPU mirrors) against the scipy.fftpack the reference calls
    (core.py:751,753): every bit, forward and inverse, rows vectorised and not."""
    import ctypes
    from scipy.fftpack import irfft, rfft
    L = _fft_lib()
    rng = np.random.default_rng(5)
    for n in lengths:
        x = rng.standard_normal((6, n)).astype(np.float32) * 100
        for fwd, ref in ((1, rfft(x, axis=-1)), (0, irfft(x, axis=-1))):
            y = x.copy()
            rc = L.orc_fftpack_r2r_f32(ctypes.c_void_p(y.ctypes.data), ctypes.c_size_t(6), ctypes.c_size_t(n), ctypes.c_int(fwd))
            assert rc == 0 and np.array_equal(ref.view(np.uint32), y.view(np.uint32)), (n, fwd)


def test_fft_restatement_even_lengths_above_1000():
    """ducc0 runs even lengths > 1000 as a half-length complex transform; the restatement (and the GPU mirror of it)
    covers every half length with a prime factor >= 7 (generic radices, Bluestein pass for factors >= 110).  Pinned on the 4-wide SIMD rows scipy processes (8 rows),
    which is what a sub-band with > 1000 rows consists of; the production db9 lengths are checked on leftover rows too."""
    import ctypes
    from scipy.fftpack import irfft, rfft
    L = _fft_lib()
    L.orc_fft_class.restype = ctypes.c_int
    L.orc_fft_class.argtypes = [ctypes.c_size_t]
    rng = np.random.default_rng(6)
    covered = 0
    for n in list(range(1002, 1500, 2)) + [1670, 2048, 2636, 2648, 3252]:
        cls = L.orc_fft_class(n)
        x = rng.standard_normal((8, n)).astype(np.float32) * 50
        y = x.copy()
        rc = L.orc_fftpack_r2r_f32(ctypes.c_void_p(y.ctypes.data), ctypes.c_size_t(8), ctypes.c_size_t(n), ctypes.c_int(1))
        if cls < 0:
            assert rc == 1
            continue
        covered += 1
        z = x.copy()
        L.orc_fftpack_r2r_f32(ctypes.c_void_p(z.ctypes.data), ctypes.c_size_t(8), ctypes.c_size_t(n), ctypes.c_int(0))
        assert np.array_equal(rfft(x, axis=-1).view(np.uint32), y.view(np.uint32)), n
        assert np.array_equal(irfft(x, axis=-1).view(np.uint32), z.view(np.uint32)), n
    assert covered > 100
    for n in (1326, 1150, 1332, 1302):           # db9 on 2048^2 tiles, sigma 250 / 100 / 256, and 2000^2: any row count
        for rows in (1, 6):
            x = rng.standard_normal((rows, n)).astype(np.float32)
            y = x.copy()
            L.orc_fftpack_r2r_f32(ctypes.c_void_p(y.ctypes.data), ctypes.c_size_t(rows), ctypes.c_size_t(n), ctypes.c_int(1))
            assert np.array_equal(rfft(x, axis=-1).view(np.uint32), y.view(np.uint32)), (n, rows)


@pytest.mark.parametrize("n,freq", [(7, 0.25), (64, 1 / 64.0), (301, 0.01), (2048, 1 / 2048.0)])
def test_sosfiltfilt_restatement(n, freq):
    """the operation sequence csrc/bleach.cu executes for butter_lowpass_filter (core.py:493-499) against the real
    scipy.signal.sosfiltfilt: every bit of the float64 result."""
    from scipy.signal import butter, sosfilt_zi, sosfiltfilt
    rng = np.random.default_rng(n)
    x = rng.uniform(0.7, 9.0, (5, n)).astype(np.float32)
    sos = butter(1, freq, output='sos')
    zi = sosfilt_zi(sos)
    assert sos.shape == (1, 6) and sos[0, 2] == 0 and sos[0, 5] == 0 and sos[0, 3] == 1 and zi[0, 1] == 0
    got = orc.sosfiltfilt_order1_restated(x, sos, float(zi[0, 0]))
    ref = sosfiltfilt(sos, x)
    assert ref.dtype == np.float64 and np.array_equal(got.view(np.uint64), ref.view(np.uint64))


def test_oracle_reproduces_the_reference_on_its_real_tile():
    """tests/golden/real_tile_golden.*: the reference source run verbatim on LsDeconvolveMultiGPU/supplements/test.png
    (Step-3 call, process_images.py:420-447); the restated oracle must give the same array (CRC32 of the whole output)."""
    import json
    import zlib
    from tests.golden import make_golden_real_tile as rt
    meta = json.loads((ROOT / "tests" / "golden" / "real_tile_golden.json").read_text())
    img = rt.load_input()
    assert zlib.crc32(img.tobytes()) == meta["input_crc32"]
    name = "real_tile_step3_db9_bidir_u16"
    res = orc.process_img(img.copy(), tile_size=img.shape, **rt.CASES[name])
    assert str(res.dtype) == meta[name]["dtype"]
    assert zlib.crc32(np.ascontiguousarray(res).tobytes()) == meta[name]["crc32"]
