"""Isotropic down-sampling of the post-stitch path (parallel_image_processor.py:156-187, 371-433).
CPU part: the host logic against the reference's own method source (extracted from /root/reference with ast and executed
verbatim on a dummy object; skipped where the reference is absent) and the oracle restatement's invariants.
GPU part: the CUDA path against the oracle (real scipy.ndimage underneath), bit for bit."""
import ast
import textwrap
from math import ceil
from pathlib import Path

import numpy as np
import pytest

from oracle import pystripe_oracle as orc
from tools import synth

REF = Path("/root/reference/parallel_image_processor.py")
CASES = [((301, 407), (1.0, 0.8, 0.8), 10.0), ((512, 768), (2.0, 1.3, 1.3), 5.0), ((257, 1001), (1.0, 2.0, 0.7), 9.0),
         ((96, 128), (1.0, 1.0, 1.0), 2.0)]


@pytest.mark.skipif(not REF.exists(), reason="the reference tree is only present in the build container")
@pytest.mark.parametrize("shape,voxel,target", CASES)
@pytest.mark.parametrize("rotated", [False, True])
@pytest.mark.parametrize("alternating", [True, False])
def test_down_sampling_target_matches_reference_method(shape, voxel, target, rotated, alternating):
    from pystripe import isotropic as iso
    tree = ast.parse(REF.read_text())
    fn = next(n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name == "calculate_down_sampling_target")
    src = textwrap.dedent(ast.get_source_segment(REF.read_text(), fn))
    ns = {"array": np.array, "np_floor": np.floor, "np_sqrt": np.sqrt, "np_max": np.max, "np_mean": np.mean,
          "Tuple": tuple, "print": lambda *a, **k: None}
    exec(src, ns)

    class Dummy:
        pass
    d = Dummy()
    d.source_voxel, d.shape, d.target_voxel = voxel, shape, target
    new_shape = (shape[1], shape[0]) if rotated else shape
    ns["calculate_down_sampling_target"](d, new_shape, rotated, alternating)
    got_shape, got_methods = iso.calculate_down_sampling_target(shape, new_shape, voxel, target, rotated, alternating)
    assert got_shape == tuple(int(v) for v in d.target_shape)
    names = {np.max: "max", np.mean: "mean", None: None}
    assert got_methods == [(names[a], names[b]) for a, b in d.down_sampling_methods]


def test_reduced_shape_and_kernels():
    from pystripe import isotropic as iso
    from scipy.ndimage import _filters
    t, m = iso.calculate_down_sampling_target((4096, 6144), (4096, 6144), (1.0, 0.8, 0.8), 10.0)
    assert t == (328, 492) and m == [("max", "mean"), ("mean", "max"), ("max", "mean")]
    pre = iso.reduced_shape((4096, 6144), t, m)
    assert pre == (512, 768)
    aa = iso.anti_aliasing_kernels(pre, t)
    for ax in range(2):
        sd = (pre[ax] / t[ax] - 1) / 2
        r = int(4.0 * sd + 0.5)
        assert aa[ax][0] == r and np.array_equal(aa[ax][1], _filters._gaussian_kernel1d(sd, 0, r)[::-1])


def test_oracle_down_sample_invariants():
    img = synth.plane(4, (301, 407))
    out = orc.down_sample_xy(img, (24, 33), [("max", "mean"), ("mean", "max"), ("max", "mean")])
    assert out.shape == (24, 33) and out.dtype == np.float32 and img.min() <= out.min() and out.max() <= img.max()
    assert not orc.down_sample_xy(np.full((64, 64), 9, np.uint16), (8, 8), [("max", "mean")]).any()
    z = np.stack([out, out * 2, out * 3])
    r = orc.down_sample_z(z, ["max", "mean", "max"], "uint16")
    m1 = np.maximum(z[0], z[1])
    m2 = np.maximum(z[2], 0)
    assert np.array_equal(r, np.clip((m1 + m2) / np.float32(2), 0, 65535).astype(np.uint16))


def _gpu_vs_oracle_xy(img, target, methods):
    from pystripe import isotropic as iso
    got = iso.down_sample_xy(img, target, methods)
    planes = img if img.ndim == 3 else img[None]
    got3 = got if img.ndim == 3 else got[None]
    for z in range(planes.shape[0]):
        ref = orc.down_sample_xy(planes[z], target, methods)
        assert got3[z].dtype == np.float32 and got3[z].shape == ref.shape
        same = got3[z].view(np.uint32) == ref.view(np.uint32)
        assert same.all(), (planes.shape, target, float(same.mean()), float(np.abs(got3[z] - ref).max()))


@pytest.mark.gpu
@pytest.mark.parametrize("shape,voxel,target", CASES)
def test_gpu_down_sample_xy_bit_exact(shape, voxel, target):
    from pystripe import isotropic as iso
    t, m = iso.calculate_down_sampling_target(shape, shape, voxel, target)
    _gpu_vs_oracle_xy(synth.plane(6, shape), t, m)
    _gpu_vs_oracle_xy((synth.plane(7, shape) >> 4).astype(np.uint8), t, m)
    _gpu_vs_oracle_xy(synth.plane(8, shape).astype(np.float32) * np.float32(0.37), t, [("mean", "mean")] * len(m))


@pytest.mark.gpu
def test_gpu_down_sample_xy_batch_uniform_and_large():
    from pystripe import isotropic as iso
    shape = (2000, 3000)
    t, m = iso.calculate_down_sampling_target(shape, shape, (1.0, 0.8, 0.8), 10.0)
    stack = np.stack([synth.plane(1, shape, n_blobs=20), np.full(shape, 300, np.uint16), synth.plane(2, shape, n_blobs=20)])
    _gpu_vs_oracle_xy(stack, t, m)
    got = iso.down_sample_xy(stack, t, m)
    assert not got[1].any()                     # uniform plane -> zeros (parallel_image_processor.py:373-374)


@pytest.mark.gpu
@pytest.mark.parametrize("nz,dtype,post", [(5, "float32", None), (8, "uint16", None), (3, "uint8", None), (4, "uint8", "uint8"), (1, "uint16", None)])
def test_gpu_down_sample_z_bit_exact(nz, dtype, post):
    from pystripe import isotropic as iso
    rng = np.random.default_rng(nz)
    z = (rng.uniform(0, 70000 if dtype != "uint8" or post is None else 255, (nz, 41, 57))).astype(np.float32)
    methods = ["max" if i % 2 == 0 else "mean" for i in range(4)]
    got = iso.down_sample_z(z, methods, dtype, post)
    ref = orc.down_sample_z(z, methods, dtype, post)
    assert got.dtype == ref.dtype and np.array_equal(got, ref)
    assert not iso.down_sample_z(np.full((4, 8, 8), 3.0, np.float32), methods).any()


@pytest.mark.gpu
def test_gpu_isotropic_accepts_cuda_tensors_zero_copy():
    import torch
    from pystripe import isotropic as iso
    shape = (301, 407)
    t, m = iso.calculate_down_sampling_target(shape, shape, (1.0, 0.8, 0.8), 10.0)
    img = synth.plane(9, shape)
    d = torch.from_numpy(img).cuda()
    got = iso.down_sample_xy(d, t, m)
    assert got.is_cuda and got.dtype == torch.float32 and tuple(got.shape) == t
    ref = orc.down_sample_xy(img, t, m)
    assert np.array_equal(got.cpu().numpy().view(np.uint32), ref.view(np.uint32))
    stack = torch.stack([got, got * 2, got * 0.5])
    z = iso.down_sample_z(stack, ["max", "mean"], "uint16")
    assert z.is_cuda and z.dtype == torch.uint16
    assert np.array_equal(z.cpu().numpy(), orc.down_sample_z(stack.cpu().numpy(), ["max", "mean"], "uint16"))
    assert iso.is_uniform(torch.full((64, 64), 5, dtype=torch.uint16, device="cuda")) and not iso.is_uniform(d)
