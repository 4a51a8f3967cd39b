"""ORACLE — TEST INFRASTRUCTURE ONLY.

Pins the restated oracle (oracle/pystripe_oracle.py) against the reference's own source executed verbatim
(oracle/ref_runner.py).  Only meaningful where /root/reference exists (the build container).  Run as
    python -m oracle.ref_check            (re-execs itself with numpy's AVX512 dispatch disabled)
Exit code 0 = every case bit-identical.
"""
import json
import sys

from . import ref_runner

if __name__ == "__main__":
    ref_runner.ensure_pinned_env()

import numpy as np  # noqa: E402

from . import pystripe_oracle as orc  # noqa: E402


def cases():
    sys.path.insert(0, str(ref_runner.Path(__file__).resolve().parent.parent))
    from tools import synth
    rng = np.random.default_rng(7)
    a = synth.plane(0, (96, 128))
    b = rng.integers(0, 65536, size=(70, 91)).astype(np.uint16)
    c = synth.plane(1, (30, 30))
    yield "fs_db10_wrap", a, dict(sigma=(24, 24), wavelet="db10")
    yield "fs_db9_reflect_bidir", a, dict(sigma=(16, 16), wavelet="db9", padding_mode="reflect", bidirectional=True)
    yield "fs_db4_dual_sigma", a, dict(sigma=(8, 32), wavelet="db4", padding_mode="symmetric")
    yield "fs_fullrange_odd", b, dict(sigma=(10, 10), wavelet="db5", padding_mode="edge")
    yield "fs_tiny_min34", c, dict(sigma=(2, 2), wavelet="db2", padding_mode="constant")
    yield "fs_level2", a, dict(sigma=(24, 24), wavelet="db3", level=2)
    yield "fs_nolog", a, dict(sigma=(24, 24), wavelet="db3", log1p_normalization_needed=False)


def main():
    core, ls = ref_runner.load()
    bad = 0
    report = {}
    for name, img, kw in cases():
        ref = core.filter_streaks(img.copy(), **kw)
        got = orc.filter_streaks(img.copy(), **kw)
        ok = ref.dtype == got.dtype and ref.shape == got.shape and np.array_equal(ref, got)
        report[name] = bool(ok)
        bad += not ok
    # scalar helpers
    for shape, s in [((2048, 2048), 256), ((2048, 2048), 250), ((2000, 2000), 100), ((1600, 2000), 512), ((64, 64), 1)]:
        ok = core.calculate_pad_size(shape, s) == orc.calculate_pad_size(shape, s)
        report[f"pad_{shape}_{s}"] = bool(ok)
        bad += not ok
    for n, s in [(8, 2.0), (1333, 128.87), (347, 33.5)]:
        ok = np.array_equal(core.np_notch(n, s), orc.np_notch(n, s))
        report[f"notch_{n}"] = bool(ok)
        bad += not ok
    # process_img: integer path (no flat: the as-written flat path raises), 8-bit, dark, down-sample, flips
    from tools import synth
    img = synth.plane(3, (96, 128))
    for name, kw in [
        ("pi_dark_8bit", dict(sigma=(16, 16), wavelet="db6", dark=100, convert_to_8bit=True, bit_shift_to_right=3)),
        ("pi_ds_max", dict(sigma=(16, 16), wavelet="db6", down_sample=(2, 2), dark=90, padding_mode="reflect")),
        ("pi_ds_min_rot", dict(sigma=(0, 0), down_sample=(3, 2), down_sample_method="min", rotate=90, flip_upside_down=True)),
        ("pi_16bit", dict(sigma=(16, 16), wavelet="db6", convert_to_16bit=True, rotate=270)),
        ("pi_lightsheet", dict(sigma=(0, 0), lightsheet=True, artifact_length=30, background_window_size=40, dark=100)),
        ("pi_resize_down", dict(sigma=(16, 16), wavelet="db6", dark=100, new_size=(81, 108), convert_to_8bit=True, bit_shift_to_right=3)),
        ("pi_resize_up", dict(sigma=(12, 12), wavelet="db4", new_size=(130, 171), rotate=90, flip_upside_down=True)),
        ("pi_resize_ls", dict(sigma=(0, 0), lightsheet=True, artifact_length=30, background_window_size=40, dark=100, new_size=(120, 160))),
    ]:
        ref = core.process_img(img.copy(), **kw)
        got = orc.process_img(img.copy(), quirks=True, **kw)
        ok = ref.dtype == got.dtype and ref.shape == got.shape and np.array_equal(ref, got)
        report[name] = bool(ok)
        bad += not ok
    # float32 input through process_img with flat (reference works when the caller passes float32)
    flat = orc.normalize_flat(synth.flat_field((96, 128)))
    ref = core.process_img(img.astype(np.float32), flat=flat, sigma=(16, 16), wavelet="db6", dark=100, d_type="uint16")
    got = orc.process_img(img.copy(), flat=flat, sigma=(16, 16), wavelet="db6", dark=100, d_type="uint16")
    ok = ref.dtype == got.dtype and np.array_equal(ref, got)
    report["pi_flat_float"] = bool(ok)
    bad += not ok
    uni = np.full((64, 80), 7, np.uint16)
    ref = core.process_img(uni.copy(), sigma=(8, 8), wavelet="db2", down_sample=(2, 2), rotate=90, convert_to_8bit=True)
    got = orc.process_img(uni.copy(), sigma=(8, 8), wavelet="db2", down_sample=(2, 2), rotate=90, convert_to_8bit=True)
    ok = ref.dtype == got.dtype and ref.shape == got.shape and np.array_equal(ref, got)
    report["pi_uniform"] = bool(ok)
    bad += not ok
    print(json.dumps({"bad": int(bad), "cases": report}))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
