"""batch_filter under torchrun, one process per GPU: every rank filters its contiguous Z shard of the same
directory; after a barrier rank 0 checks every output file against the oracle.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_batch_torchrun.py"""
import os
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "image-preprocessing-pipeline_b200")]
import numpy as np
import torch
import torch.distributed as dist
from pystripe import core
from tools import synth

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
root = Path("/tmp/b2s_batch_check")
src, dst = root / "in", root / "out"
n, shape = 37, (200, 260)
flat = synth.flat_field(shape)
if rank == 0:
    (src / "ch0").mkdir(parents=True, exist_ok=True)
    for z in range(n):
        core.imsave_tif(src / "ch0" / f"img_{z:05d}.tif", synth.plane(70 + z, shape), compression=None)
dist.barrier()
kw = dict(sigma=(32, 32), level=0, wavelet="db9", padding_mode="reflect", bidirectional=True, dark=105)
rc = core.batch_filter(src, dst, workers=4, threads_per_gpu=4, flat=flat, d_type="uint16", compression=None, **kw)
assert rc == 0
dist.barrier()
if rank == 0:
    from oracle import pystripe_oracle as orc
    nflat = orc.normalize_flat(flat)
    bad = 0
    for z in range(n):
        got = core.imread_tif_raw_png(dst / "ch0" / f"img_{z:05d}.tif")
        ref = orc.process_img(synth.plane(70 + z, shape), flat=nflat, d_type=np.dtype("uint16"), **kw)
        bad += not (got is not None and np.array_equal(got, ref))
    print(f"batch_filter under torchrun x{world}: {n} files, {bad} differ from the oracle")
    assert bad == 0
dist.destroy_process_group()
