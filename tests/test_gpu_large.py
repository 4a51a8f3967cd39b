"""GPU parity on planes larger than a camera tile (the post-stitch caller of process_img,
process_images.py:702-740 through parallel_image_processor.py:371-435) against golden digests written by the reference
source run verbatim (tests/golden/make_golden_large.py): strided sample, corner / centre crops, CRC32 of the whole output.
Also covers the workspace fit (kSlots x max_batch planes of a 13464 x 15512 working image do not fit one B200)."""
import json
import zlib
from pathlib import Path

import numpy as np
import pytest

from tests.golden import make_golden_large as gl

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"
MIN_EXACT = 0.9999


@pytest.mark.parametrize("name", list(gl.CASES))
def test_stitched_slice_matches_reference_digest(name):
    from pystripe import core
    meta = json.loads((GOLD / "large_plane_golden.json").read_text())[name]
    gold = np.load(GOLD / "large_plane_golden.npz")
    img, kw = gl.plane_for(name)
    kw["tile_size"] = img.shape
    got = core.process_img(img, **kw)
    core.clear_plan_cache()                 # gigabytes of workspace: give them back before the next case
    assert list(got.shape) == meta["shape"] and str(got.dtype) == meta["dtype"]
    worst, n_diff, n_px = 0, 0, 0
    for key, arr in gl.digest(got).items():
        ref = gold[f"{name}/{key}"]
        d = np.abs(arr.astype(np.int64) - ref.astype(np.int64))
        worst = max(worst, int(d.max()))
        n_diff += int((d != 0).sum())
        n_px += d.size
    crc_equal = zlib.crc32(np.ascontiguousarray(got).tobytes()) == meta["crc32"]
    print(f"{name}: digest pixels {n_px}, differing {n_diff}, max |diff| {worst}, whole-plane crc32 equal: {crc_equal}")
    assert worst <= 1, f"{name}: max |diff| = {worst}"
    assert 1 - n_diff / n_px >= MIN_EXACT, f"{name}: {n_diff} of {n_px} digest pixels differ"


def test_workspace_is_fitted_to_device_memory():
    """a plan whose max_batch x 5 slots would need far more than 180 GB is created with a smaller batch / fewer slots."""
    from pystripe import _native, core
    plan = core._get_plan(0, (12000, 16000), _native.U16, process=1, sigma=(512, 512), level=0, wavelet="db9",
                          threshold=None, padding_mode="reflect", bidirectional=False, log1p=True, max_batch=32)
    import torch
    free, total = torch.cuda.mem_get_info(0)
    assert 0 < plan.info.workspace_bytes < 0.75 * total
    core.clear_plan_cache()
