"""Full-size (2048 x 2048) parity of every BASELINE.json configuration against the oracle, the reference's real tile
against goldens written by the reference source run verbatim, and the public host paths (pageable / pinned arrays,
batch_filter over files) the bench times.  All through the public pystripe API -> C ABI -> CUDA."""
import json
import zlib
from pathlib import Path

import numpy as np
import pytest

from oracle import pystripe_oracle as orc
from tests.golden import make_golden_real_tile as rt
from tools import synth

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]
MIN_EXACT = 0.9999
H = W = 2048
REPORT = {}

# BASELINE.json configs[1..4] as bench.py --config 2..5 runs them, plus Step 3 as shipped (process_images.py:420-447)
FULL = {
    "config2_flat_dark_reflect": (dict(sigma=(256, 256), wavelet="db10", level=0, padding_mode="reflect", dark=100), True, 1.0),
    "config3_flat_gauss_ds2_8bit": (dict(sigma=(256, 256), wavelet="db10", level=0, padding_mode="reflect", dark=100,
                                         gaussian_filter_2d=True, down_sample=(2, 2), convert_to_8bit=True,
                                         bit_shift_to_right=8), True, MIN_EXACT),   # cv2's float Gaussian is not bit-pinned
    "config4_lightsheet_destripe": (dict(sigma=(256, 256), wavelet="db10", level=0, padding_mode="wrap", lightsheet=True), False, 1.0),
    "config5_coif15_dual_sigma": (dict(sigma=(128, 512), wavelet="coif15", level=0, padding_mode="reflect"), False, 1.0),
    "step3_db9_sigma250_bidirectional": (dict(sigma=(250, 250), wavelet="db9", level=0, padding_mode="reflect",
                                              bidirectional=True, d_type="uint16"), False, 1.0),
}


def _cmp(name, got, ref, min_exact):
    assert got.dtype == ref.dtype and got.shape == ref.shape, (name, got.dtype, ref.dtype, got.shape, ref.shape)
    d = np.abs(got.astype(np.int64) - ref.astype(np.int64))
    exact = float((d == 0).mean())
    REPORT[name] = {"max_abs_diff": int(d.max()), "exact_fraction": exact, "pixels": int(d.size)}
    assert d.max() <= 1, f"{name}: max |diff| = {d.max()}"
    assert exact >= min_exact, f"{name}: only {exact:.6%} bit-exact"


@pytest.mark.parametrize("name", list(FULL))
def test_full_size_config_against_oracle(name):
    from pystripe import core
    kw, use_flat, min_exact = FULL[name]
    stack = np.stack([synth.plane(3, (H, W)), synth.plane(4, (H, W))])      # a batch of two through one plan
    flat = core.normalize_flat(synth.flat_field((H, W))) if use_flat else None
    got = core.process_img(stack, flat=flat, **kw)
    ref = orc.process_img(stack[1].copy(), flat=None if flat is None else orc.normalize_flat(synth.flat_field((H, W))), **kw)
    _cmp(name, np.asarray(got[1]), ref, min_exact)
    core.clear_plan_cache()


@pytest.mark.parametrize("name", list(rt.CASES))
def test_reference_real_tile_matches_golden(name):
    """LsDeconvolveMultiGPU/supplements/test.png through the Step-3 call: goldens written by the reference source itself."""
    from pystripe import core
    meta = json.loads((ROOT / "tests" / "golden" / "real_tile_golden.json").read_text())
    gold = np.load(ROOT / "tests" / "golden" / "real_tile_golden.npz")
    img = rt.load_input()
    assert zlib.crc32(img.tobytes()) == meta["input_crc32"]
    got = np.asarray(core.process_img(img, tile_size=img.shape, **rt.CASES[name]))
    m = meta[name]
    assert list(got.shape) == m["shape"] and str(got.dtype) == m["dtype"]
    n_diff = n_px = worst = 0
    for key, arr in rt.digest(got).items():
        ref = gold[f"{name}/{key}"]
        d = np.abs(arr.astype(np.int64) - ref.astype(np.int64))
        worst, n_diff, n_px = max(worst, int(d.max())), n_diff + int((d != 0).sum()), n_px + d.size
    crc_equal = zlib.crc32(np.ascontiguousarray(got).tobytes()) == m["crc32"]
    REPORT["real_tile/" + name] = {"max_abs_diff": worst, "digest_pixels": n_px, "differing": n_diff, "crc32_equal": crc_equal}
    assert worst <= 1 and 1 - n_diff / n_px >= MIN_EXACT, (name, worst, n_diff)


# ------------------------------------------------------------------------------------------------ host paths
@pytest.mark.parametrize("n_planes,max_batch,ramp", [(5, 3, None), (8, 3, None), (11, 3, None), (7, 2, "5"), (13, 5, None)])
def test_host_batches_never_exceed_the_plan_batch(n_planes, max_batch, ramp, monkeypatch):
    """with B = 3 the ramp-down issued batches of 4 whenever 5 planes were left (5, 8, 11 ... planes),
    writing one plane past every slot buffer.  Every plane of every such stack must equal the single-plane result."""
    from pystripe import core
    if ramp:
        monkeypatch.setenv("B2S_HOST_RAMP", ramp)      # read once per process: covers the clamp when it is the first use
    stack = synth.stack(n_planes, (96, 130), seed=77)
    kw = dict(sigma=(12, 12), wavelet="db4", dark=50, padding_mode="reflect")
    got = core.process_img(stack, _max_batch=max_batch, **kw)
    for z in range(n_planes):
        assert np.array_equal(got[z], orc.process_img(stack[z].copy(), **kw)), z


def test_pageable_pinned_and_device_inputs_agree():
    """the three ways a stack reaches the GPU — pageable numpy (copy threads), page-locked numpy (no staging), CUDA tensor
    (no copy) — and the pooled page-locked result arrays."""
    import torch
    from pystripe import core
    stack = synth.stack(21, (300, 420), seed=5)
    kw = dict(sigma=(32, 32), wavelet="db10", dark=100, padding_mode="reflect")
    a = core.process_img(stack, **kw)
    pinned = core.pinned_empty(stack.shape, stack.dtype)
    pinned[:] = stack
    b = core.process_img(pinned, **kw)
    c = core.process_img(torch.from_numpy(stack).cuda(), **kw).cpu().numpy()
    assert np.array_equal(a, b) and np.array_equal(a, c)
    assert np.array_equal(a[20], orc.process_img(stack[20].copy(), **kw))
    # results live in recycled page-locked blocks: dropping one and asking again must not corrupt the one still held
    keep = a.copy()
    del b
    d = core.process_img(stack[::-1].copy(), **kw)
    assert np.array_equal(a, keep) and np.array_equal(d[0], a[20])


def test_two_threads_share_one_plan_safely():
    """two threads calling with the same arguments get the same plan; runs are serialised per plan."""
    import threading
    from pystripe import core
    stack = synth.stack(12, (200, 260), seed=9)
    kw = dict(sigma=(16, 16), wavelet="db6", padding_mode="reflect")
    ref = core.filter_streaks(stack, **kw)
    res = [None] * 4

    def work(i):
        for _ in range(3):
            res[i] = core.filter_streaks(stack, **kw)
    ts = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    for r in res:
        assert np.array_equal(r, ref)


def test_batch_filter_native_codec_and_odd_tiles(tmp_path):
    """batch_filter end to end on files: deflate and LZW inputs through the native codec, a big-endian .raw tile, one tile
    of another shape that is resized to tile_size (core.py:1540-1549) and one unreadable file."""
    from PIL import Image
    from pystripe import _io, core, raw
    src, dst = tmp_path / "in", tmp_path / "out"
    src.mkdir()
    shape = (128, 160)
    planes = {}
    for z in range(9):
        img = synth.plane(60 + z, shape)
        name = f"img_{z:04d}.tif"
        if z % 3 == 0:
            _io.write_tiff(src / name, img, ("ADOBE_DEFLATE", 1))
        elif z % 3 == 1:
            Image.fromarray(img).save(src / name, format="TIFF", compression="tiff_lzw")
        else:
            _io.write_tiff(src / name, img, None)
        planes[name] = img
    with open(src / "be_0100.raw", "wb") as f:
        np.array([shape[1], shape[0]], dtype=">u4").tofile(f)
        synth.plane(70, shape).astype(">u2").tofile(f)
    planes["be_0100.tif"] = synth.plane(70, shape)
    odd = synth.plane(71, (150, 200))
    _io.write_tiff(src / "odd_0200.tif", odd, None)
    (src / "bad_0300.tif").write_bytes(b"II*\0garbage")
    kw = dict(sigma=(16, 16), level=0, wavelet="db9", padding_mode="reflect", bidirectional=True, dark=105)
    core.NUM_RETRIES, saved = 2, core.NUM_RETRIES
    try:
        rc = core.batch_filter(src, dst, workers=6, threads_per_gpu=4, d_type="uint16", tile_size=shape, **kw)
    finally:
        core.NUM_RETRIES = saved
    assert rc == 0
    for name, img in planes.items():
        got = core.imread_tif_raw_png(dst / name)
        ref = orc.process_img(img.copy(), d_type=np.dtype("uint16"), **kw)
        assert got.dtype == np.uint16 and np.array_equal(got, ref), name
    # the odd tile: anti-aliased resize to tile_size first (float64 in the reference), then the same pipeline
    got = core.imread_tif_raw_png(dst / "odd_0200.tif")
    resized = orc.skimage_resize(odd, shape, preserve_range=True, anti_aliasing=True)
    ref = orc.process_img(resized, d_type=np.dtype("uint16"), **kw)
    assert got.shape == shape
    _cmp("batch_filter/odd_tile_resized", got, ref.astype(np.uint16), MIN_EXACT)
    # the unreadable file becomes a zeros tile because tile_size and d_type are known (core.py:1521-1531)
    assert not core.imread_tif_raw_png(dst / "bad_0300.tif").any()


def test_resize_to_tile_matches_skimage_restatement():
    from pystripe import core
    for dtype, shape, target in ((np.uint16, (150, 200), (128, 160)), (np.uint8, (97, 131), (64, 64)),
                                 (np.float32, (120, 90), (100, 90)), (np.uint16, (100, 100), (140, 120))):
        img = synth.plane(80, shape)
        img = (img >> 4).astype(np.uint8) if dtype == np.uint8 else img.astype(dtype)
        got = core.resize_to_tile(img, target)
        ref = orc.skimage_resize(img, target, preserve_range=True, anti_aliasing=True)
        assert got.shape == target and got.dtype == np.float32
        assert np.array_equal(got, ref.astype(np.float32)), (dtype, shape, target, float(np.abs(got - ref).max()))


def test_zz_write_config_parity_report():
    out = ROOT / "gpurun_out"
    out.mkdir(exist_ok=True)
    (out / "parity_report_configs.json").write_text(json.dumps(REPORT, indent=1, default=str))
