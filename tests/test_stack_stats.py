"""Stack statistics: the host half — numpy arithmetic on intensity histograms — against the pixel-level
oracle, the oracle's multi-Otsu against a float64 brute force, and the histogram all-reduce under a world-size-2 gloo group.
The GPU half (b2s_histogram) is covered by tests/test_gpu_stats.py."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

from oracle import pystripe_oracle as orc
from pystripe import stack_stats as ss
from tools import synth

ROOT = Path(__file__).resolve().parents[1]


def _hist(img):
    return np.bincount(img.reshape(-1).astype(np.int64), minlength=65536).astype(np.int64)


def _planes():
    a = synth.plane(3, (256, 320))
    b = synth.plane(4, (200, 200))
    b[:50] = 0                                         # zeros: log1p(0) = 0 is the image minimum
    c = (synth.plane(5, (128, 128)) >> 3).astype(np.uint16)
    return [a, b, c]


def test_histogram_binning_equals_numpy_on_the_pixels():
    """np.histogram over the occupied integer levels weighted by their counts == np.histogram over the pixels."""
    for img in _planes():
        log = np.log1p(img, dtype=np.float32)
        ref, edges = np.histogram(log.reshape(-1), bins=256)
        v, w = ss._log_values(_hist(img))
        got, edges2 = np.histogram(v, bins=256, weights=w)
        assert np.array_equal(ref, got) and np.array_equal(edges, edges2)


@pytest.mark.parametrize("nbins", [24, 256])
def test_multiotsu_vectorised_search_equals_the_scalar_restatement(nbins):
    for img in _planes()[: 3 if nbins == 24 else 1]:
        log = np.log1p(img, dtype=np.float32)
        ref = orc.threshold_multiotsu(log, classes=4, nbins=nbins)
        got = ss.threshold_multiotsu_from_histogram(_hist(img), classes=4, nbins=nbins)
        assert all(isinstance(t, np.float32) for t in got)              # process_images.py:628-630 asserts float32
        assert got == ref, (got, ref)


def test_multiotsu_restatement_against_float64_brute_force():
    """the definition: the three thresholds that maximise the between-class variance of the 4 classes (float64, no LUT)."""
    img = _planes()[0]
    log = np.log1p(img, dtype=np.float32)
    nbins = 20
    hist, edges = np.histogram(log.reshape(-1), bins=nbins)
    p = hist / hist.sum()
    centers = (edges[:-1] + edges[1:]) / 2
    lev = np.arange(nbins, dtype=np.float64)
    lev[0] = 1.0                       # the Cython code seeds first_moment[0] with prob[0] (weight 1 for bin 0)
    best, arg = -1.0, None
    for a in range(nbins - 3):
        for b in range(a + 1, nbins - 2):
            for c in range(b + 1, nbins - 1):
                s = 0.0
                for lo, hi in ((0, a), (a + 1, b), (b + 1, c), (c + 1, nbins - 1)):
                    w = p[lo:hi + 1].sum()
                    if w > 0:
                        if lo == 0 and hi == 0:
                            continue                                   # var_btwcls[0] is never written
                        s += (p[lo:hi + 1] * lev[lo:hi + 1]).sum() ** 2 / w
                if s > best * (1 + 1e-12):
                    best, arg = s, (a, b, c)
    ref = tuple(np.float32(centers[i]) for i in arg)
    assert orc.threshold_multiotsu(log, classes=4, nbins=nbins) == ref


def test_bit_shift_and_percentile_from_histogram_equal_the_pixel_level_oracle():
    for img in _planes():
        log = np.log1p(img, dtype=np.float32)
        h = _hist(img)
        lb, mb, ub = ss.threshold_multiotsu_from_histogram(h)
        for thr in (lb, mb, ub):
            for q in (50.0, 99.9, 99.99, 100):
                sel = log[log > thr]
                assert ss.percentile_above_from_histogram(h, thr, q) == orc._percentile_numba(sel, q), (thr, q)
            assert ss.estimate_bit_shift_from_histogram(h, thr, 99.99) == orc.estimate_bit_shift(log, thr, 99.99)
        assert ss.estimate_bit_shift_from_histogram(h, np.float32(1e9), 99.99) == orc.estimate_bit_shift(log, np.float32(1e9), 99.99)


def test_estimate_img_related_params_follows_the_reference_sampling(monkeypatch):
    """three samples at 25 / 50 / 75 %, uniform planes skipped, max of the bit shifts, clip levels of the last sample."""
    planes = [synth.plane(10 + z, (96, 96)) for z in range(8)]
    planes[4][:] = 7                                                     # the 50 % sample is uniform: the next plane is used
    planes[6] = (planes[6].astype(np.uint32) * 9).clip(0, 65535).astype(np.uint16)   # brighter: larger bit shift
    monkeypatch.setattr(ss, "histogram", lambda img, **kw: _hist(np.asarray(img)))
    seen = []

    def read(z):
        seen.append(z)
        return planes[z]
    bg, shift, sigma, cmin, cmed, cmax, freq = ss.estimate_img_related_params(read, 8, need_bleach_correction=True,
                                                                             tile_size=(96, 96))
    assert seen == [2, 4, 5, 6]
    logs = [np.log1p(planes[z], dtype=np.float32) for z in (2, 5, 6)]
    ref_shifts = [orc.estimate_bit_shift(l, orc.threshold_multiotsu(l)[2], 99.99) for l in logs]
    assert shift == max(ref_shifts) and (cmin, cmed, cmax) == orc.threshold_multiotsu(logs[2])
    assert bg == int(np.round(np.expm1(cmin))) and sigma == (192, 192) and freq is None


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE=str(world), RANK=str(rank), LOCAL_RANK=str(rank))
    sys.path[:0] = [str(ROOT), str(ROOT / "image-preprocessing-pipeline_b200")]
    import torch.distributed as dist
    from pystripe import core, stack_stats
    dist.init_process_group("gloo", rank=rank, world_size=world)
    stack = np.stack([synth.plane(20 + z, (64, 64)) for z in range(7)])
    lo, hi = core.z_shard(len(stack), world, rank)                        # this rank's Z-shard
    stack_stats.histogram = lambda img, out=None, **kw: _hist(np.asarray(img)) + (0 if out is None else out)
    res = stack_stats.whole_stack_params(stack[lo:hi], need_bleach_correction=True)
    if rank == 0:
        np.save(out, np.array([res["bit_shift"], res["pixels"], res["background"], float(res["clip_max"])]))
    dist.barrier()
    dist.destroy_process_group()


def test_whole_stack_statistics_all_reduce_two_ranks_gloo(tmp_path):
    """each rank counts its shard; one all-reduce of 65 536 counters; every rank ends with the whole-stack statistics."""
    import torch.multiprocessing as mp
    out = tmp_path / "res.npy"
    mp.spawn(_worker, args=(2, 29700 + os.getpid() % 2000, str(out)), nprocs=2, join=True)
    shift, pixels, bg, cmax = np.load(out)
    stack = np.stack([synth.plane(20 + z, (64, 64)) for z in range(7)])
    h = _hist(stack)
    lb, mb, ub = ss.threshold_multiotsu_from_histogram(h)
    assert pixels == stack.size and cmax == float(ub)
    assert shift == ss.estimate_bit_shift_from_histogram(h, ub, 99.99) and bg == int(np.round(np.expm1(lb)))
