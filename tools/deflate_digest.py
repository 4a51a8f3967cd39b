"""md5 of the packed streams core.gpu_deflate produces for a fixed set of planes (to compare code paths: B2S_DEFLATE_BUILD=0/1)."""
import hashlib
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "image-preprocessing-pipeline_b200")]
import numpy as np
import torch
from pystripe import core
from tools import synth

rng = np.random.default_rng(1)
planes = np.concatenate([synth.stack(4, (2048, 2048)), rng.integers(0, 65536, (1, 2048, 2048)).astype(np.uint16),
                         np.full((1, 2048, 2048), 77, np.uint16)])
d = core.gpu_deflate(torch.from_numpy(planes).cuda())
print(hashlib.md5(d.data.tobytes()).hexdigest(), d.data.size, hashlib.md5(d.sizes.tobytes()).hexdigest())
