"""GPU half of the stack statistics: b2s_histogram against numpy.bincount, and the whole
estimate_img_related_params chain against the pixel-level oracle."""
import numpy as np
import pytest

from oracle import pystripe_oracle as orc
from tools import synth

pytestmark = pytest.mark.gpu


def test_histogram_is_exact_for_host_and_device_planes():
    import torch
    from pystripe import stack_stats as ss
    rng = np.random.default_rng(1)
    stack = np.stack([synth.plane(z, (300, 420)) for z in range(19)])      # more planes than one host slice (16)
    stack[3] = 65535
    stack[4, :100] = 0
    stack[5] = rng.integers(0, 65536, stack[5].shape).astype(np.uint16)   # every bin
    ref = np.stack([np.bincount(p.ravel(), minlength=65536) for p in stack]).astype(np.int64)
    assert np.array_equal(ss.histogram(stack), ref.sum(0))
    assert np.array_equal(ss.histogram(stack, per_plane=True), ref)
    assert np.array_equal(ss.histogram(stack[7]), ref[7])
    t = torch.from_numpy(stack).cuda()
    assert np.array_equal(ss.histogram(t).cpu().numpy(), ref.sum(0))
    acc = ss.histogram(t[:10])
    acc = ss.histogram(t[10:], out=acc)                                    # histograms add: shards of a stack
    assert np.array_equal(acc.cpu().numpy(), ref.sum(0))
    u8 = (stack[:3] >> 8).astype(np.uint8)
    assert np.array_equal(ss.histogram(u8)[:256], np.bincount(u8.ravel(), minlength=256))
    big = synth.plane(1, (2048, 2048))
    assert np.array_equal(ss.histogram(big), np.bincount(big.ravel(), minlength=65536))
    with pytest.raises(TypeError):
        ss.histogram(stack.astype(np.float32))


def test_estimate_img_related_params_matches_the_pixel_level_oracle():
    from pystripe import stack_stats as ss
    planes = [synth.plane(30 + z, (512, 512)) for z in range(12)]
    planes[6][:] = 3                                                        # uniform sample plane: skipped
    bg, shift, sigma, cmin, cmed, cmax, freq = ss.estimate_img_related_params(lambda z: planes[z], 12, need_bleach_correction=True,
                                                                             tile_size=(512, 512))
    logs = [np.log1p(planes[z], dtype=np.float32) for z in (3, 7, 9)]
    ref = [orc.threshold_multiotsu(l) for l in logs]
    assert (cmin, cmed, cmax) == ref[2]
    assert shift == max(orc.estimate_bit_shift(l, r[2], 99.99) for l, r in zip(logs, ref))
    assert bg == int(np.round(np.expm1(ref[2][0]))) and sigma == (1024, 1024)
    whole = ss.whole_stack_params(np.stack(planes))
    assert whole["pixels"] == 12 * 512 * 512 and 0 <= whole["bit_shift"] <= 8
