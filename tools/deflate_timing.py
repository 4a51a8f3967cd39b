"""time core.gpu_deflate on 32 planes of 2048^2 (device part and the whole call incl. D2H of the packed streams)."""
import sys
import time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "image-preprocessing-pipeline_b200")]
import numpy as np
import torch
from pystripe import core
from tools import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
base = synth.stack(4, (2048, 2048))
t = torch.from_numpy(np.concatenate([base] * (n // 4))).cuda()
for _ in range(2):
    d = core.gpu_deflate(t)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(reps):
    d = core.gpu_deflate(t)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / reps
print(f"gpu_deflate: {n} planes, {dt * 1e3:.2f} ms per call, {n * 2048 * 2048 / dt / 1e6:.0f} Mpixel/s, ratio {t.numel() * 2 / d.data.size:.3f}")
