"""Seeded cases shared by make_golden.py (reference, verbatim) and the tests (oracle / GPU)."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from tools import synth  # noqa: E402


def normalize_flat(flat):
    f = flat.astype(np.float32)
    return f / f.max()


MASK_LOG_THRESHOLD = 6.8          # log1p(900) ~ 6.80


def mask_plane(shape=(150, 180), seed=21):
    """a plane with a dim background (< 900), a bright ring (hole inside) and a bright block touching one border."""
    img = (synth.plane(seed, shape) // 8).astype(np.uint16)
    img = np.minimum(img, 700)
    y, x = np.mgrid[0:shape[0], 0:shape[1]]
    r = np.hypot(y - 70, x - 80)
    ring = (r > 22) & (r < 45)
    img[ring] += 2500
    img[100:150, 0:40] += 1800
    img[10:14, 150:170] += 3000          # a thin streak the opening removes
    return img


def all_cases():
    """yields (name, kind, img, kwargs).  kind in {"filter_streaks", "process_img"}; '_flat' = flat-field array."""
    rng = np.random.default_rng(7)
    a = synth.plane(0, (96, 128))
    b = rng.integers(0, 65536, size=(70, 91)).astype(np.uint16)
    c = synth.plane(1, (30, 30))
    d = synth.plane(2, (160, 200))
    fs = "filter_streaks"
    yield "fs_db10_wrap", fs, a, dict(sigma=(24, 24), wavelet="db10")
    yield "fs_db9_reflect_bidir", fs, a, dict(sigma=(16, 16), wavelet="db9", padding_mode="reflect", bidirectional=True)
    yield "fs_db4_dual_sigma", fs, a, dict(sigma=(8, 32), wavelet="db4", padding_mode="symmetric")
    yield "fs_fullrange_odd", fs, b, dict(sigma=(10, 10), wavelet="db5", padding_mode="edge")
    yield "fs_tiny_min34", fs, c, dict(sigma=(2, 2), wavelet="db2", padding_mode="constant")
    yield "fs_level2", fs, a, dict(sigma=(24, 24), wavelet="db3", level=2)
    yield "fs_nolog", fs, a, dict(sigma=(24, 24), wavelet="db3", log1p_normalization_needed=False)
    yield "fs_nolog_f32", fs, a.astype(np.float32), dict(sigma=(24, 24), wavelet="db3", log1p_normalization_needed=False)
    yield "fs_f32", fs, a.astype(np.float32), dict(sigma=(24, 24), wavelet="db8", padding_mode="reflect")
    yield "fs_db10_big", fs, d, dict(sigma=(32, 32), wavelet="db10", padding_mode="reflect")
    yield "fs_u8", fs, (a >> 4).astype(np.uint8), dict(sigma=(16, 16), wavelet="db4")
    # bleach correction (core.py:501-559): explicit clip levels (log domain), non-max method; real scipy.signal in the reference
    bl = dict(bleach_correction_frequency=1 / 64.0, bleach_correction_clip_min=4.8, bleach_correction_clip_med=5.6,
              bleach_correction_clip_max=7.4)
    yield "fs_bleach_db9_reflect", fs, d, dict(sigma=(32, 32), wavelet="db9", padding_mode="reflect", **bl)
    yield "fs_bleach_bidir_wrap_lowclip", fs, a, dict(sigma=(16, 16), wavelet="db6", bidirectional=True,
                                                     bleach_correction_frequency=0.01, bleach_correction_clip_min=0.5,
                                                     bleach_correction_clip_med=5.0, bleach_correction_clip_max=6.5)
    yield "fs_bleach_constant_pad_f32", fs, a.astype(np.float32), dict(sigma=(24, 24), wavelet="db4", padding_mode="constant", **bl)
    yield "fs_clipmin_constant_pad_nofreq", fs, a, dict(sigma=(24, 24), wavelet="db4", padding_mode="constant",
                                                       bleach_correction_clip_min=4.9, bleach_correction_clip_med=6.0,
                                                       bleach_correction_clip_max=8.0)
    zeros_in = a.copy()
    zeros_in[20:40, 30:90] = 0
    yield "fs_bleach_zero_patch_odd", fs, zeros_in[:95, :127], dict(sigma=(8, 8), wavelet="db2", padding_mode="symmetric", **bl)
    yield "fs_bleach_max_method", fs, d, dict(sigma=(32, 32), wavelet="db9", padding_mode="reflect",
                                             bleach_correction_max_method=True, **bl)
    yield "fs_bleach_max_method_zero_rows_odd", fs, zeros_in[:95, :127], dict(sigma=(0, 0), bleach_correction_max_method=True, **bl)
    yield "fs_bleach_only_sigma0_odd", fs, zeros_in[:95, :127], dict(sigma=(0, 0), **bl)
    # masking (core.py:475-489, 1079-1080): the threshold is compared with the LOG image; a bright ring leaves a hole that
    # no corner reaches, a dim corner patch keeps background connected to a corner
    yield "fs_mask_explicit_reflect", fs, mask_plane(), dict(sigma=(16, 16), wavelet="db4", padding_mode="reflect", enable_masking=True,
                                                              bleach_correction_clip_med=MASK_LOG_THRESHOLD, close_steps=5, open_steps=9)
    yield "fs_mask_even_kernels_constant", fs, mask_plane(), dict(sigma=(8, 24), wavelet="db2", padding_mode="constant", enable_masking=True,
                                                                   bleach_correction_clip_min=3.0, bleach_correction_clip_med=MASK_LOG_THRESHOLD,
                                                                   bleach_correction_clip_max=9.0, close_steps=6, open_steps=12)
    yield "fs_mask_nolog_int", fs, mask_plane(), dict(sigma=(16, 16), wavelet="db4", padding_mode="reflect", enable_masking=True,
                                                       log1p_normalization_needed=False, bleach_correction_clip_med=900, close_steps=5,
                                                       open_steps=9)
    yield "fs_mask_sigma0_bleach", fs, mask_plane(), dict(sigma=(0, 0), enable_masking=True, close_steps=5, open_steps=9, **dict(bl, bleach_correction_clip_med=MASK_LOG_THRESHOLD, bleach_correction_clip_max=9.5))
    yield "fs_mask_default_steps_small_tile", fs, mask_plane(), dict(sigma=(16, 16), wavelet="db4", enable_masking=True,
                                                                      bleach_correction_clip_med=MASK_LOG_THRESHOLD)
    pi = "process_img"
    img = synth.plane(3, (96, 128))
    yield "pi_bleach_only_flat_16bit", pi, img, dict(sigma=(0, 0), dark=100, convert_to_16bit=True,
                                                    _flat=normalize_flat(synth.flat_field((96, 128))), **bl)
    yield "pi_bleach_dark_8bit", pi, img, dict(sigma=(16, 16), wavelet="db6", dark=100, convert_to_8bit=True,
                                              bit_shift_to_right=3, padding_mode="reflect", **bl)
    yield "pi_dark_8bit", pi, img, dict(sigma=(16, 16), wavelet="db6", dark=100, convert_to_8bit=True, bit_shift_to_right=3)
    yield "pi_ds_max", pi, img, dict(sigma=(16, 16), wavelet="db6", down_sample=(2, 2), dark=90, padding_mode="reflect")
    yield "pi_ds_min_rot", pi, img, dict(sigma=(0, 0), down_sample=(3, 2), down_sample_method="min", rotate=90,
                                        flip_upside_down=True)
    yield "pi_16bit_rot270", pi, img, dict(sigma=(16, 16), wavelet="db6", convert_to_16bit=True, rotate=270)
    yield "pi_rot180_flip", pi, img, dict(sigma=(12, 12), wavelet="db2", rotate=180, flip_upside_down=True, dark=120)
    yield "pi_lightsheet", pi, img, dict(sigma=(0, 0), lightsheet=True, artifact_length=30, background_window_size=40,
                                        dark=100)
    yield "pi_flat_dark", pi, img, dict(sigma=(16, 16), wavelet="db6", dark=100, padding_mode="reflect",
                                       _flat=normalize_flat(synth.flat_field((96, 128))))
    yield "pi_flat_8bit", pi, img, dict(sigma=(16, 16), wavelet="db10", dark=100, convert_to_8bit=True,
                                       bit_shift_to_right=4, _flat=normalize_flat(synth.flat_field((96, 128))))
    # new_size (core.py:1356-1359): skimage.transform.resize restated over the real scipy.ndimage.zoom
    yield "pi_resize_down_8bit", pi, img, dict(sigma=(16, 16), wavelet="db6", dark=100, new_size=(81, 108),
                                              convert_to_8bit=True, bit_shift_to_right=3)
    yield "pi_resize_up_rot", pi, img, dict(sigma=(12, 12), wavelet="db4", new_size=(130, 171), rotate=90,
                                           flip_upside_down=True)
    yield "pi_resize_down_nodestripe", pi, img, dict(sigma=(0, 0), new_size=(50, 127), dark=105)
    yield "pi_resize_lightsheet", pi, img, dict(sigma=(0, 0), lightsheet=True, artifact_length=30,
                                               background_window_size=40, dark=100, new_size=(120, 160))
    yield "pi_resize_flat", pi, img, dict(sigma=(16, 16), wavelet="db6", dark=100, padding_mode="reflect", new_size=(77, 99),
                                         _flat=normalize_flat(synth.flat_field((96, 128))))
    # new_size larger along one axis, smaller along the other: anti_aliasing=True with a Gaussian along the shrinking axis
    # (float64 for an integer image, float32 after a flat-field)
    yield "pi_resize_mixed_aa_x", pi, img, dict(sigma=(16, 16), wavelet="db6", dark=100, new_size=(120, 50))
    yield "pi_resize_mixed_aa_x_8bit_rot", pi, img, dict(sigma=(0, 0), new_size=(97, 31), convert_to_8bit=True,
                                                        bit_shift_to_right=3, rotate=270)
    yield "pi_resize_mixed_aa_flat", pi, img, dict(sigma=(16, 16), wavelet="db6", dark=100, padding_mode="reflect",
                                                  new_size=(150, 77), _flat=normalize_flat(synth.flat_field((96, 128))))
    yield "pi_resize_mixed_lightsheet", pi, img, dict(sigma=(0, 0), lightsheet=True, artifact_length=30,
                                                     background_window_size=40, dark=100, new_size=(100, 90))
    yield "pi_uniform", pi, np.full((64, 80), 7, np.uint16), dict(sigma=(8, 8), wavelet="db2", down_sample=(2, 2),
                                                                  rotate=90, convert_to_8bit=True)
