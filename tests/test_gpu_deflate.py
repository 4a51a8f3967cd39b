"""b2s_deflate_strips: every strip of every result plane as a zlib stream produced on the GPU (csrc/deflate.cu) — checked by
inflating the streams with zlib itself (round trip, byte for byte), by reading the TIFF files laid out from them with Pillow
and the native codec, and through batch_filter's default compression=('ADOBE_DEFLATE', 1)."""
import zlib

import numpy as np
import pytest

from tools import synth

pytestmark = pytest.mark.gpu


def _round_trip(planes):
    import torch
    from pystripe import core
    d = core.gpu_deflate(torch.from_numpy(planes).cuda())
    n = planes.shape[0] if planes.ndim == 3 else 1
    assert d.sizes.shape == d.offsets.shape and d.sizes.shape[0] == n
    assert int(d.offsets.ravel()[-1] + d.sizes.ravel()[-1]) == d.data.size       # packed back to back
    flat = d.offsets.ravel().astype(np.int64)
    assert np.array_equal(flat[1:], flat[:-1] + d.sizes.ravel()[:-1])
    src = planes if planes.ndim == 3 else planes[None]
    for i in range(n):
        for k, (o, sz) in enumerate(zip(d.offsets[i], d.sizes[i])):
            stream = d.data[int(o):int(o) + int(sz)].tobytes()
            assert stream[:2] == b"\x78\x01"
            want = src[i, k * d.rows_per_strip:(k + 1) * d.rows_per_strip].tobytes()
            assert zlib.decompress(stream) == want, (i, k)
        assert np.array_equal(d.inflate(i), src[i])
    return d


@pytest.mark.parametrize("shape,dtype", [((3, 300, 411), np.uint16), ((2, 97, 53), np.uint8), ((1, 64, 2048), np.uint16),
                                         ((2, 5, 7), np.uint16), ((1, 33, 40001), np.uint8)])
def test_streams_inflate_to_the_planes(shape, dtype):
    rng = np.random.default_rng(11)
    y, x = np.mgrid[0:shape[1], 0:shape[2]]
    planes = np.stack([((np.sin(x / 37.0) + np.cos(y / 23.0)) * 500 + 2000 + rng.normal(0, 30, shape[1:])) for _ in range(shape[0])])
    planes = (planes / (16 if dtype == np.uint8 else 1)).astype(dtype)
    d = _round_trip(planes)
    assert d.data.size < planes.nbytes or planes.nbytes < 4096                     # it does compress camera-like data


def test_degenerate_and_hostile_distributions():
    rng = np.random.default_rng(5)
    const = np.full((2, 40, 64), 1234, np.uint16)                                  # two byte values + end of block
    zeros = np.zeros((1, 40, 64), np.uint8)                                        # one byte value: both codes one bit
    noise = rng.integers(0, 65536, (2, 64, 512)).astype(np.uint16)                 # incompressible: stays within the bound
    # Fibonacci-like byte frequencies force code lengths beyond 15 bits before the limiter
    fib = np.concatenate([np.full(int(1.62 ** k) + 1, k, np.uint8) for k in range(2, 27)])
    fib = np.resize(fib, (1, 8, fib.size // 8))
    for planes in (const, zeros, noise, fib):
        _round_trip(np.ascontiguousarray(planes))


def test_full_size_planes_compress_like_zlib_level_1():
    planes = synth.stack(4, (2048, 2048))
    d = _round_trip(planes)
    ours = d.data.size
    ref = sum(len(zlib.compress(planes[i].tobytes(), 1)) for i in range(4))
    assert ours < 1.15 * ref, (ours, ref)


def test_batch_filter_default_compression_goes_through_the_gpu_encoder(tmp_path, monkeypatch):
    from PIL import Image
    from pystripe import _io, core
    stack = synth.stack(6, (256, 320))
    (tmp_path / "in").mkdir()
    for z in range(6):
        _io.write_tiff(tmp_path / "in" / f"t_{z:03d}.tif", stack[z], None)
    calls = []
    real = core.gpu_deflate
    monkeypatch.setattr(core, "gpu_deflate", lambda res: (calls.append(tuple(res.shape)), real(res))[1])
    kw = dict(sigma=(32, 32), level=0, wavelet="db5", padding_mode="reflect", bidirectional=False, dark=50)
    assert core.batch_filter(tmp_path / "in", tmp_path / "out", workers=4, compression=("ADOBE_DEFLATE", 1), **kw) == 0
    assert sum(c[0] for c in calls) == 6
    want = core.process_img(stack, **kw)
    for z in range(6):
        f = tmp_path / "out" / f"t_{z:03d}.tif"
        assert _io.probe(f)[2].compression == 8
        assert np.array_equal(_io.read(f), want[z])
        with Image.open(f) as im:
            assert np.array_equal(np.array(im), want[z])
        assert f.stat().st_size < want[z].nbytes
    monkeypatch.setenv("B200STRIPE_GPU_DEFLATE", "0")                               # the host encoder stays available
    calls.clear()
    assert core.batch_filter(tmp_path / "in", tmp_path / "out_host", workers=4, compression=("ADOBE_DEFLATE", 1), **kw) == 0
    assert not calls and np.array_equal(_io.read(tmp_path / "out_host" / "t_002.tif"), want[2])
