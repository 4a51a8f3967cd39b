"""Multi-GPU host logic on CPU: Z partition and the one-process-per-GPU batch_filter path under a world-size-2 gloo
group.  The GPU kernels are not involved: process_img is replaced by a marker so that only scheduling, sharding, file
naming and I/O are exercised (planes are independent, no collective on the data path)."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def test_z_shard_is_a_partition():
    from pystripe import core
    for n in (0, 1, 7, 250, 2000, 10001):
        for world in (1, 2, 3, 4, 8):
            spans = [core.z_shard(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    with pytest.raises(ValueError):
        core.z_shard(10, 2, 2)


def _worker(rank, world, port, src, dst):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE=str(world), RANK=str(rank),
                      LOCAL_RANK=str(rank))
    sys.path[:0] = [str(ROOT), str(ROOT / "image-preprocessing-pipeline_b200")]
    import torch
    import torch.distributed as dist
    from pystripe import core
    dist.init_process_group("gloo", rank=rank, world_size=world)
    seen = []

    def fake_process_img(stack, **kw):          # stands in for the GPU call: marks every plane with the rank
        seen.append(int(stack.shape[0]))
        return (stack // 2 + rank).astype(stack.dtype)

    core.process_img = fake_process_img
    core.use_device = lambda d: __import__("contextlib").nullcontext()
    core._batch_buffer = lambda device, shape, dtype: np.empty(shape, dtype)   # pinned memory needs the GPU
    core._to_device = lambda buf, device: buf
    core._own_stream = lambda device, index=0: __import__("contextlib").nullcontext()
    rc = core.batch_filter(Path(src), Path(dst), workers=2, threads_per_gpu=3, sigma=(8, 8), wavelet="db2")
    t = torch.tensor([sum(seen), rc], dtype=torch.int64)
    dist.all_reduce(t)                            # bookkeeping only (tests): total planes processed, sum of return codes
    if rank == 0:
        Path(dst, "total.txt").write_text(f"{int(t[0])} {int(t[1])}")
    dist.barrier()
    dist.destroy_process_group()


def test_batch_filter_two_ranks_gloo(tmp_path):
    import torch.multiprocessing as mp
    sys.path[:0] = [str(ROOT / "image-preprocessing-pipeline_b200")]
    from pystripe import core
    src, dst = tmp_path / "in", tmp_path / "out"
    (src / "a").mkdir(parents=True)
    rng = np.random.default_rng(3)
    planes = {}
    for z in range(7):
        img = rng.integers(0, 4000, (24, 40)).astype(np.uint16)
        name = f"a/img_{z:04d}.tif"
        core.imsave_tif(src / name, img, compression=None)
        planes[name] = img
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(src), str(dst)), nprocs=2, join=True)
    total, rc = (int(v) for v in (dst / "total.txt").read_text().split())
    assert total == 7 and rc == 0
    lo0, hi0 = core.z_shard(7, 2, 0)
    names = sorted(planes)
    for i, name in enumerate(names):              # every plane written exactly once, by the rank that owns it
        out = core.imread_tif_raw_png(dst / name)
        owner = 0 if lo0 <= i < hi0 else 1
        assert np.array_equal(out, planes[name] // 2 + owner), name
