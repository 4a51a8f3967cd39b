"""Where the file-based path spends its time (GPU box): codec read / write rates into / out of page-locked batches on tmpfs,
process_img on pinned batches, and the whole batch_filter, per stage.  python tools/io_probe.py [n_files]"""
import os
import shutil
import sys
import tempfile
import time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "image-preprocessing-pipeline_b200")]
import numpy as np
from pystripe import _io, core
from tools import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 96
H = W = 2048
work = Path(tempfile.mkdtemp(prefix="b2s_io_", dir="/dev/shm"))
src, dst = work / "in", work / "out"
src.mkdir(); dst.mkdir()
base = synth.stack(8, (H, W))
stack = np.concatenate([base] * (n // 8))
paths = [src / f"img_{z:05d}.tif" for z in range(n)]
outs = [dst / f"img_{z:05d}.tif" for z in range(n)]
_io.write_tiff_batch(paths, stack, None)
gb = stack.nbytes / 1e9
pinned = core.pinned_empty(stack.shape, stack.dtype)
pageable = np.empty_like(stack); pageable[:] = 0
print("cores", os.cpu_count())
for th in (1, 2, 4, 8, 16):
    for name, buf in (("pinned", pinned), ("pageable", pageable)):
        t = time.perf_counter(); st = _io.read_batch(paths, buf, threads=th); dt = time.perf_counter() - t
        assert not any(st)
        print(f"read_batch  {name:8s} threads={th:2d}: {gb / dt:6.2f} GB/s")
    t = time.perf_counter(); st = _io.write_tiff_batch(outs, pinned, None, threads=th); dt = time.perf_counter() - t
    print(f"write_batch stored   threads={th:2d}: {gb / dt:6.2f} GB/s")
for th in (8, 16):
    t = time.perf_counter(); st = _io.write_tiff_batch(outs[:16], pinned[:16], ("ADOBE_DEFLATE", 1), threads=th); dt = time.perf_counter() - t
    print(f"write_batch deflate1 threads={th:2d}: {gb * 16 / n / dt:6.2f} GB/s")
    t = time.perf_counter(); st = _io.read_batch(outs[:16], pageable[:16], threads=th); dt = time.perf_counter() - t
    print(f"read_batch  deflate1 threads={th:2d}: {gb * 16 / n / dt:6.2f} GB/s")
# batches of 8 like batch_filter issues them
t = time.perf_counter()
for i in range(0, n, 8):
    _io.read_batch(paths[i:i + 8], pinned[i:i + 8], threads=8)
print(f"read in batches of 8, 8 threads: {gb / (time.perf_counter() - t):6.2f} GB/s")
kw = dict(sigma=(256, 256), wavelet="db10", padding_mode="reflect", dark=100)
for _ in range(2):
    r = core.process_img(pinned[:8], _max_batch=8, **kw)
t = time.perf_counter()
for i in range(0, n, 8):
    r = core.process_img(pinned[i:i + 8], _max_batch=8, **kw)
dt = time.perf_counter() - t
print(f"process_img on pinned batches of 8: {n * H * W / dt / 1e6:8.0f} Mpx/s")
for _ in range(2):
    r = core.process_img(pinned, _max_batch=32, **kw)
t = time.perf_counter(); r = core.process_img(pinned, _max_batch=32, **kw); dt = time.perf_counter() - t
print(f"process_img on the whole pinned stack: {n * H * W / dt / 1e6:8.0f} Mpx/s")
for hosts in ("4", "8", "12", "16"):
    pass
for workers in (8, 16, 32):
    for rep in range(2):
        shutil.rmtree(dst, ignore_errors=True)
        t = time.perf_counter()
        so = sys.stdout; sys.stdout = open(os.devnull, "w")
        try:
            rc = core.batch_filter(src, dst, workers=workers, threads_per_gpu=8, compression=None, **kw)
        finally:
            sys.stdout = so
        dt = time.perf_counter() - t
    print(f"batch_filter workers={workers}: {n * H * W / dt / 1e6:8.0f} Mpx/s rc={rc}", flush=True)
shutil.rmtree(work, ignore_errors=True)
