"""filter_streaks with bleach correction (core.py:501-559) on a batch of 2048^2 planes, for ncu and for timing
(`python tools/prof_bleach.py 8 time` prints us/plane of the bleach stage from the library's per-class timers)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "image-preprocessing-pipeline_b200")]
import numpy as np
import torch
from pystripe import core, _native
from tools import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
base = synth.stack(4, (2048, 2048))
stack = torch.from_numpy(np.concatenate([base] * (n // 4))).cuda()
kw = dict(sigma=(256, 256), wavelet="db10", padding_mode="reflect", bleach_correction_frequency=1 / 2048.0,
          bleach_correction_clip_min=4.7, bleach_correction_clip_med=5.5, bleach_correction_clip_max=8.0)
out = core.filter_streaks(stack, **kw)
torch.cuda.synchronize()
if "time" in sys.argv:
    ctx = _native.context(0)
    ctx.timing_enable(True); ctx.timing_read(reset=True)
    for _ in range(3):
        out = core.filter_streaks(stack, **kw)
    tm = ctx.timing_read(reset=True)
    ctx.timing_enable(False)
    print({k: round(v[0] / (3 * n) * 1e3, 1) for k, v in tm.items() if v[1]}, "us/plane")
print("ok", int(out[0, ::256, ::256].sum()))
